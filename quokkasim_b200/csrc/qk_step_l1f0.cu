/* step kernel instantiations: logging on, feature set 0 (see qk_step_inst.cuh) */
#include "qk_step_inst.cuh"

StepKernel qk_pick_step_l1f0(uint32_t nt) { return qk_pick_step<true, 0>(nt); }
