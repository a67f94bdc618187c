/* lane-group kernel instantiations: logging on, feature set 1 (see qk_group_inst.cuh) */
#include "qk_group_inst.cuh"

StepKernel qk_pick_group_l1f1(uint32_t g) { return qk_pick_group<true, 1>(g); }
StepKernel qk_pick_phased_l1f1(uint32_t g) { return qk_pick_phased<true, 1>(g); }
