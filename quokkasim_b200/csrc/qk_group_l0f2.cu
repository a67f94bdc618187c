/* lane-group kernel instantiations: logging off, feature set 2 (see qk_group_inst.cuh) */
#include "qk_group_inst.cuh"

StepKernel qk_pick_group_l0f2(uint32_t g) { return qk_pick_group<false, 2>(g); }
StepKernel qk_pick_phased_l0f2(uint32_t g) { return qk_pick_phased<false, 2>(g); }
