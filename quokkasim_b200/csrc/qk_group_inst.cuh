/*
 * qk_group_inst.cuh -- one translation unit per (logging, feature set) instantiates the lane-group kernel for every
 * group width; the units compile in parallel (nvcc --threads), reproducibly.
 */
#ifndef QK_GROUP_INST_CUH
#define QK_GROUP_INST_CUH
#include "qk_group.cuh"

typedef void (*StepKernel)(const KParams);

template <bool LOG, int FT>
static StepKernel qk_pick_group(uint32_t g) {
  switch (g) {
    case 4: return qk_group_kernel<LOG, 4, FT>;
    case 8: return qk_group_kernel<LOG, 8, FT>;
    case 16: return qk_group_kernel<LOG, 16, FT>;
    default: return nullptr;
  }
}
template <bool LOG, int FT>
static StepKernel qk_pick_phased(uint32_t g) {
  switch (g) {
    case 4: return qk_phased_kernel<LOG, 4, FT>;
    case 8: return qk_phased_kernel<LOG, 8, FT>;
    case 16: return qk_phased_kernel<LOG, 16, FT>;
    default: return nullptr;
  }
}
#endif
