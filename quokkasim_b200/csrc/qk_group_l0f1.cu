/* lane-group kernel instantiations: logging off, feature set 1 (see qk_group_inst.cuh) */
#include "qk_group_inst.cuh"

StepKernel qk_pick_group_l0f1(uint32_t g) { return qk_pick_group<false, 1>(g); }
StepKernel qk_pick_phased_l0f1(uint32_t g) { return qk_pick_phased<false, 1>(g); }
