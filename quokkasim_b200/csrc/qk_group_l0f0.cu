/* lane-group kernel instantiations: logging off, feature set 0 (see qk_group_inst.cuh) */
#include "qk_group_inst.cuh"

StepKernel qk_pick_group_l0f0(uint32_t g) { return qk_pick_group<false, 0>(g); }
StepKernel qk_pick_phased_l0f0(uint32_t g) { return qk_pick_phased<false, 0>(g); }
