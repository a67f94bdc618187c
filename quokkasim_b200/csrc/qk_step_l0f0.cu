/* step kernel instantiations: logging off, feature set 0 (see qk_step_inst.cuh) */
#include "qk_step_inst.cuh"

StepKernel qk_pick_step_l0f0(uint32_t nt) { return qk_pick_step<false, 0>(nt); }
