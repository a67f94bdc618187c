/* lane-group kernel instantiations: logging on, feature set 2 (see qk_group_inst.cuh) */
#include "qk_group_inst.cuh"

StepKernel qk_pick_group_l1f2(uint32_t g) { return qk_pick_group<true, 2>(g); }
StepKernel qk_pick_phased_l1f2(uint32_t g) { return qk_pick_phased<true, 2>(g); }
