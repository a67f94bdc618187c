/* step kernel instantiations: logging off, feature set 3 = general + user-defined vector resources of up to
 * QK_VEC_MAX_DIM fields (see qk_step_inst.cuh) */
#include "qk_step_inst.cuh"

StepKernel qk_pick_step_l0f3(uint32_t nt) { return qk_pick_step<false, 3>(nt); }
