/* step kernel instantiations: logging on, feature set 2 = general + container family (see qk_step_inst.cuh) */
#include "qk_step_inst.cuh"

StepKernel qk_pick_step_l1f2(uint32_t nt) { return qk_pick_step<true, 2>(nt); }
