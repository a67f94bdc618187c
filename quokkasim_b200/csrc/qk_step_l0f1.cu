/* step kernel instantiations: logging off, feature set 1 (see qk_step_inst.cuh) */
#include "qk_step_inst.cuh"

StepKernel qk_pick_step_l0f1(uint32_t nt) { return qk_pick_step<false, 1>(nt); }
