/* step kernel instantiations: logging off, feature set 2 = general + container family (see qk_step_inst.cuh) */
#include "qk_step_inst.cuh"

StepKernel qk_pick_step_l0f2(uint32_t nt) { return qk_pick_step<false, 2>(nt); }
