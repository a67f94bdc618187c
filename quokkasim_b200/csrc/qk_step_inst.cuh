/*
 * qk_step_inst.cuh -- one translation unit per (logging, feature set) instantiates the step kernel for every block
 * size; the units compile in parallel (nvcc --threads) and, unlike --split-compile, reproducibly.
 */
#ifndef QK_STEP_INST_CUH
#define QK_STEP_INST_CUH
#include "qk_kernels.cuh"

typedef void (*StepKernel)(const KParams);

/* one instantiation per (logging, feature set, state placement, block size): the block size is a template constant
 * so that state-slot offsets become LDS/STS immediates.  nt == 0: HBM-resident state ; other sizes not in the list:
 * the run-time-block-size variant */
template <bool LOG, int FT>
static StepKernel qk_pick_step(uint32_t nt) {
  switch (nt) {
    case 0: return qk_step_kernel<LOG, 0, FT>;
    case 32: return qk_step_kernel<LOG, 32, FT>;
    case 64: return qk_step_kernel<LOG, 64, FT>;
    case 96: return qk_step_kernel<LOG, 96, FT>;
    case 128: return qk_step_kernel<LOG, 128, FT>;
    case 160: return qk_step_kernel<LOG, 160, FT>;
    case 192: return qk_step_kernel<LOG, 192, FT>;
    case 224: return qk_step_kernel<LOG, 224, FT>;
    case 256: return qk_step_kernel<LOG, 256, FT>;
    case 0xFFFFFFFEu: return qk_step_kernel<LOG, QK_SM_SPLIT, FT>; /* split placement, run-time block size */
    case 0xFFFFFFFDu: return qk_step_kernel<LOG, QK_SM_QSPLIT, FT>; /* ... with the two-tier scheduler queue */
    default: return qk_step_kernel<LOG, -1, FT>;
  }
}
#endif
