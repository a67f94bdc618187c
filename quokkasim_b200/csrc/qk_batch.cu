/*
 * qk_batch.cu -- batch lifecycle behind the C ABI: device memory, kernel launches, result read-back and the
 * per-GPU summary reduction.  One qk_batch == the replications [rep_offset, rep_offset + n_reps) on ONE GPU;
 * multi-GPU runs are one process (and one qk_batch) per GPU, the only cross-GPU step being an all-reduce of
 * the small arrays filled by qk_batch_reduce_stats_device (done by the caller with NCCL).
 *
 * Replaces SimInit::init + Simulation::step_until for the batched path
 * (/root/reference/quokkasim_examples/src/bin/discrete_queue.rs:109-111).
 */
#include <cuda.h> /* CUtensorMap and the enums of cuTensorMapEncodeTiled (types only: the entry point is asked of the runtime) */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "qk_step_inst.cuh"
#include "qk_model.h"

#define CUDA_TRY(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      qk_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return QK_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)


#define QK_SLICE_PARTS 8

struct qk_batch {
  StepKernel kernel;
  qk_batch_config cfg;
  QkLowered low;
  bool log;
  bool hbm; /* state operated on in place in HBM instead of staged in shared memory */
  uint32_t nt, grid, smem_bytes, table_bytes_padded;
  bool use_tma;       /* hot planes staged by TMA tensor copies (make_tensor_maps) */
  QkTmap tm64, tm32;
  uint64_t rep_pad;
  uint64_t* g64;
  uint32_t* g32;
  uint64_t* gs64; /* statistics planes */
  uint32_t* gs32;
  uint64_t* gc64; /* cold state planes (item rings, source ids of pending events) */
  uint32_t* gc32;
  uint8_t* d_tables;
  uint4* d_log;
  std::vector<double*> h_tapes; /* device pointers, host copy */
  std::vector<uint64_t> h_tape_len;
  const double** d_tapes;
  uint64_t* d_tape_len;
  bool any_tape;
  cudaStream_t stream;
  bool own_stream;
  cudaEvent_t ev0, ev1;
  bool initialised;
  float last_ms;
  uint32_t last_launches;
  uint64_t total_launches;
  /* reduction scratch */
  uint64_t* d_part; /* packed partials [n_comp][QK_PARTIAL_WORDS] */
  uint64_t* d_hist;
  uint32_t red_blocks;
  /* small persistent buffers for the result calls (no cudaMalloc / cudaFree -- a device-wide synchronisation each --
   * on the read-back path) */
  uint8_t* d_scratch;
  size_t scratch_bytes;
  uint8_t* h_pinned;
  size_t pinned_bytes;
  uint8_t* d_arena; /* ONE device allocation behind d_tables, d_tapes, d_tape_len, d_part, d_hist and d_scratch */
  size_t arena_bytes;
  uint32_t grp_reps, grp_stride64, grp_stride32, grp_bar_off, grp_x64; /* lane-group kernel geometry (qk_group.cuh) */
  bool phased;
  bool split; /* QK_BATCH_STATE_SPLIT: component records live in the cold planes (DevHeader.split) */
  /* sliced stepping: the tiles are split into QK_SLICE_PARTS parts, each advanced chunk by chunk on its own stream */
  cudaStream_t side[QK_SLICE_PARTS - 1];
  cudaEvent_t ev_fork, ev_join[QK_SLICE_PARTS - 1];
  bool have_side;
  uint32_t lanes;       /* lanes per replication: 1, or the group width of the lane-group kernel */
  uint32_t last_chunks; /* state residencies of the last step call */
  uint32_t n_sm, wave_blocks; /* SMs of the device ; blocks of the step kernel resident on the whole GPU at once */
};

/* the planes the component records live in: the hot planes, or (split placement) the cold ones */
static inline uint64_t* warm64(const qk_batch* b) { return b->split ? b->gc64 : b->g64; }
static inline uint32_t* warm32(const qk_batch* b) { return b->split ? b->gc32 : b->g32; }
/* the cold planes are [slot][replication], except under the split placements: [replication][slot] (qk_kernels.cuh) */
static inline const uint64_t* cold64_at(const qk_batch* b, uint32_t slot, uint64_t rep) {
  return b->split ? b->gc64 + (size_t)rep * b->low.hdr.n64c + slot : b->gc64 + (size_t)slot * b->rep_pad + rep;
}
static inline const uint32_t* cold32_at(const qk_batch* b, uint32_t slot, uint64_t rep) {
  return b->split ? b->gc32 + (size_t)rep * b->low.hdr.n32c + slot : b->gc32 + (size_t)slot * b->rep_pad + rep;
}
static inline const uint64_t* warm64_at(const qk_batch* b, uint32_t slot, uint64_t rep) {
  return b->split ? cold64_at(b, slot, rep) : b->g64 + (size_t)slot * b->rep_pad + rep;
}
static inline const uint32_t* warm32_at(const qk_batch* b, uint32_t slot, uint64_t rep) {
  return b->split ? cold32_at(b, slot, rep) : b->g32 + (size_t)slot * b->rep_pad + rep;
}
static inline uint32_t aos64(const qk_batch* b) { return b->split ? b->low.hdr.n64c : 0u; }
static inline uint32_t aos32(const qk_batch* b) { return b->split ? b->low.hdr.n32c : 0u; }

/* the step kernel instantiations live in qk_step_l{0,1}f{0,1}.cu */
StepKernel qk_pick_step_l0f0(uint32_t nt);
StepKernel qk_pick_step_l0f1(uint32_t nt);
StepKernel qk_pick_step_l1f0(uint32_t nt);
StepKernel qk_pick_step_l1f1(uint32_t nt);
StepKernel qk_pick_step_l0f2(uint32_t nt);
StepKernel qk_pick_step_l1f2(uint32_t nt);
StepKernel qk_pick_step_l0f3(uint32_t nt);
StepKernel qk_pick_step_l1f3(uint32_t nt);
/* ... the lane-group kernels in qk_group_l{0,1}f{0,1,2}.cu */
StepKernel qk_pick_group_l0f0(uint32_t g);
StepKernel qk_pick_group_l0f1(uint32_t g);
StepKernel qk_pick_group_l0f2(uint32_t g);
StepKernel qk_pick_group_l1f0(uint32_t g);
StepKernel qk_pick_group_l1f1(uint32_t g);
StepKernel qk_pick_group_l1f2(uint32_t g);
StepKernel qk_pick_phased_l0f0(uint32_t g);
StepKernel qk_pick_phased_l0f1(uint32_t g);
StepKernel qk_pick_phased_l0f2(uint32_t g);
StepKernel qk_pick_phased_l1f0(uint32_t g);
StepKernel qk_pick_phased_l1f1(uint32_t g);
StepKernel qk_pick_phased_l1f2(uint32_t g);
static StepKernel pick_phased_kernel(bool log, uint32_t g, uint32_t features) {
  if (features >= 2) return log ? qk_pick_phased_l1f2(g) : qk_pick_phased_l0f2(g);
  if (log) return features ? qk_pick_phased_l1f1(g) : qk_pick_phased_l1f0(g);
  return features ? qk_pick_phased_l0f1(g) : qk_pick_phased_l0f0(g);
}
static StepKernel pick_group_kernel(bool log, uint32_t g, uint32_t features) {
  if (features >= 2) return log ? qk_pick_group_l1f2(g) : qk_pick_group_l0f2(g);
  if (log) return features ? qk_pick_group_l1f1(g) : qk_pick_group_l1f0(g);
  return features ? qk_pick_group_l0f1(g) : qk_pick_group_l0f0(g);
}
static uint32_t group_max_threads(uint32_t g) { return g <= 4 ? 256u : (g == 8 ? 512u : 768u); } /* qk_group_max_threads */
static StepKernel pick_kernel(bool log, bool hbm, uint32_t nt, uint32_t features, uint32_t split = 0) {
  const uint32_t key = split == 2 ? 0xFFFFFFFDu : (split ? 0xFFFFFFFEu : (hbm ? 0u : nt));
  if (features == 3) return log ? qk_pick_step_l1f3(key) : qk_pick_step_l0f3(key);
  if (features >= 2) return log ? qk_pick_step_l1f2(key) : qk_pick_step_l0f2(key);
  if (log) return features ? qk_pick_step_l1f1(key) : qk_pick_step_l1f0(key);
  return features ? qk_pick_step_l0f1(key) : qk_pick_step_l0f0(key);
}

/* max_thr: resident threads per SM the register allocation of the kernel variant allows (its launch bounds) */
static uint32_t choose_nt(uint32_t per_rep_bytes, uint32_t table_bytes, uint32_t max_thr, uint32_t* threads_per_sm,
                          uint32_t min_nt = 64) {
  uint32_t best_nt = 0, best_thr = 0;
  for (uint32_t nt = 256; nt >= 32; nt -= 32) {
    const uint64_t bytes = (uint64_t)table_bytes + (uint64_t)per_rep_bytes * nt;
    if (bytes > 227u * 1024u - 256u) continue; /* 256 B: the kernels' static shared memory (an mbarrier) and the 128-byte alignment of the dynamic part */
    uint32_t bps = (uint32_t)((227u * 1024u) / (bytes + 1024u));
    if (bps > 32) bps = 32;
    uint32_t thr = bps * nt;
    if (thr > max_thr) thr = max_thr / nt * nt;
    if (thr > best_thr) {
      best_thr = thr;
      best_nt = nt;
    }
  }
  /* among block sizes within 12% of the best residency prefer the smallest (>= 64): a block holds its shared memory
   * until its slowest warp finishes, so small blocks recycle capacity sooner (measured on discrete_queue: 8 x 64
   * threads per SM beats 3 x 192 although the latter holds two more warps) */
  for (uint32_t nt = min_nt; nt < best_nt; nt += 32) {
    const uint64_t bytes = (uint64_t)table_bytes + (uint64_t)per_rep_bytes * nt;
    uint32_t bps = (uint32_t)((227u * 1024u) / (bytes + 1024u));
    if (bps > 32) bps = 32;
    uint32_t thr = bps * nt;
    if (thr > max_thr) thr = max_thr / nt * nt;
    if ((uint64_t)thr * 100 >= (uint64_t)best_thr * 88) {
      best_nt = nt;
      best_thr = thr;
      break;
    }
  }
  if (threads_per_sm) *threads_per_sm = best_thr;
  return best_nt;
}

struct QkSlice {
  uint32_t tile0, n_tiles; /* n_tiles == 0: every tile of the batch */
  cudaStream_t stream;
};

/* Tensor maps of the hot planes for the step kernel's staging: plane = 2-D array [n_slots][rep_pad] of 8- / 4-byte
 * elements, box = n_slots x nt (one block's replications of every slot), dense in shared memory.  Returns false where
 * the tile form does not apply (more than 256 slots or threads, rows not 16-byte multiples, no driver entry point):
 * the kernel then moves the rows one bulk copy each, as before. */
static bool make_tensor_maps(qk_batch* b) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static_assert(sizeof(QkTmap) == sizeof(CUtensorMap), "QkTmap mirrors CUtensorMap");
  const DevHeader& h = b->low.hdr;
  if (getenv("QK_NO_TMA")) return false; /* experiments: A/B against the per-row bulk copies */
  if (h.n64 == 0 || h.n64 > 256 || h.n32 > 256 || b->nt > 256 || (b->nt & 31u) || b->rep_pad >= (1ull << 31) ||
      (b->table_bytes_padded & 127u))
    return false;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
      q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  const EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  const cuuint32_t estr[2] = {1, 1};
  for (int k = 0; k < 2; ++k) {
    const uint32_t rows = k == 0 ? h.n64 : h.n32, esz = k == 0 ? 8u : 4u;
    if (rows == 0) continue;
    const cuuint64_t gdim[2] = {b->rep_pad, rows};
    const cuuint64_t gstr[1] = {b->rep_pad * esz};
    const cuuint32_t box[2] = {b->nt, rows};
    CUtensorMap tm;
    const CUresult rc = encode(&tm, k == 0 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2,
                               k == 0 ? (void*)b->g64 : (void*)b->g32, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return false;
    memcpy(k == 0 ? &b->tm64 : &b->tm32, &tm, sizeof(tm));
  }
  return true;
}

static int launch_step(qk_batch* b, uint32_t do_init, uint64_t t0, uint64_t t_until, uint64_t budget,
                       bool record_begin = true, bool record_end = true, const QkSlice* slice = nullptr) {
  KParams P;
  memset(&P, 0, sizeof(P));
  P.g64 = b->g64;
  P.g32 = b->g32;
  P.gs64 = b->gs64;
  P.gs32 = b->gs32;
  P.gc64 = b->gc64;
  P.gc32 = b->gc32;
  P.n64c = b->low.hdr.n64c;
  P.n32c = b->low.hdr.n32c;
  P.tables = b->d_tables;
  P.table_bytes = b->table_bytes_padded;
  P.n64 = b->low.hdr.n64;
  P.n32 = b->low.hdr.n32;
  P.n_reps = b->cfg.n_reps;
  P.rep_pad = b->rep_pad;
  P.rep_offset = b->cfg.rep_offset;
  P.seed = b->cfg.base_seed;
  P.log = b->d_log;
  P.log_cap = b->log ? b->cfg.log_capacity : 0;
  P.tapes = b->any_tape ? b->d_tapes : nullptr;
  P.tape_len = b->d_tape_len;
  P.t0 = t0;
  P.t_until = t_until;
  P.event_budget = budget;
  P.do_init = do_init;
  uint32_t grid = b->grid;
  P.grp_reps = b->grp_reps;
  P.grp_stride64 = b->grp_stride64;
  P.grp_stride32 = b->grp_stride32;
  P.grp_bar_off = b->grp_bar_off;
  P.grp_x64 = b->grp_x64;
  P.use_tma = b->use_tma ? 1u : 0u;
  if (b->use_tma) {
    P.tm64 = b->tm64;
    P.tm32 = b->tm32;
  }
  cudaStream_t stream = b->stream;
  if (slice && slice->n_tiles) {
    P.tile0 = slice->tile0;
    grid = slice->n_tiles;
    stream = slice->stream;
  }
  if (record_begin) {
    CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    b->last_launches = 0;
  }
  b->kernel<<<grid, b->nt, b->smem_bytes, stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  if (record_end) CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
  b->last_launches += 1;
  b->total_launches += 1;
  return QK_OK;
}

/* ================================= per-GPU summary reduction ================================= */

__global__ void qk_comp_stats_kernel(const uint8_t* tables, const uint64_t* g64, const uint32_t* g32,
                                     const uint64_t* w64, const uint32_t* w32, uint32_t wa64, uint32_t wa32, const uint64_t* gs64, const uint32_t* gs32, uint64_t rep_pad, uint64_t rep0, uint64_t n, uint32_t comp, qk_comp_stats* out) {
  const DevHeader* h = reinterpret_cast<const DevHeader*>(tables);
  const DevComp* c = reinterpret_cast<const DevComp*>(tables + h->off_comps) + comp;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const DevVec* vecs = reinterpret_cast<const DevVec*>(tables + h->off_vecs);
  if (i < n) qk_comp_stats_dev(c, vecs, comp, g64, g32, w64, w32, wa64, wa32, gs64, gs32, rep_pad, rep0 + i, &out[i]);
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t w = __shfl_xor_sync(0xFFFFFFFFu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

/* grid = (red_blocks, n_comp): the packed partials of one component (quokka_b200.h, QK_PARTIAL_WORDS / QP_*) over the
 * replications of this batch.  Everything is an exact integer -- second moments too: time_ns = a * 2^32 + b, so
 * time_ns^2 = a^2 * 2^64 + 2ab * 2^32 + b^2, and each of a^2, ab, b^2 (< 2^64) is summed as its low and high 32-bit
 * halves -- so the result does not depend on how the replications are partitioned over blocks or GPUs. */
__global__ void __launch_bounds__(256) qk_reduce_kernel(const uint8_t* tables, const uint64_t* g64,
                                                        const uint32_t* g32, const uint64_t* w64, const uint32_t* w32,
                                                        uint32_t wa64, uint32_t wa32, const uint64_t* gs64,
                                                        const uint32_t* gs32, uint64_t rep_pad, uint64_t n_reps,
                                                        uint64_t* part) {
  const DevHeader* h = reinterpret_cast<const DevHeader*>(tables);
  const uint32_t comp = blockIdx.y;
  const DevComp* c = reinterpret_cast<const DevComp*>(tables + h->off_comps) + comp;
  uint64_t a[QK_PARTIAL_WORDS];
  for (int k = 0; k < QK_PARTIAL_WORDS; ++k) a[k] = (k >= QP_MIN_COUNT && k <= QP_NMAX_TIME) ? ~0ull : 0ull;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reps; r += (uint64_t)gridDim.x * blockDim.x) {
    qk_comp_stats s;
    qk_comp_stats_dev(c, reinterpret_cast<const DevVec*>(tables + h->off_vecs), comp, g64, g32, w64, w32, wa64, wa32, gs64, gs32, rep_pad, r, &s);
    const uint64_t lo = s.time_ns & 0xFFFFFFFFull, hi = s.time_ns >> 32;
    a[QP_COUNT] += s.count;
    a[QP_TIME_LO] += lo;
    a[QP_TIME_HI] += hi;
    a[QP_LEVEL] += s.level;
    if (comp == 0) {
      a[QP_NEVENTS] += g64[(size_t)G64_NEVENTS * rep_pad + r];
      a[QP_NFAULTED] += g32[(size_t)G32_STATUS * rep_pad + r] ? 1 : 0;
      a[QP_NREPS] += 1;
    }
    a[QP_MIN_COUNT] = s.count < a[QP_MIN_COUNT] ? s.count : a[QP_MIN_COUNT];
    a[QP_MIN_TIME] = s.time_ns < a[QP_MIN_TIME] ? s.time_ns : a[QP_MIN_TIME];
    a[QP_NMAX_COUNT] = ~s.count < a[QP_NMAX_COUNT] ? ~s.count : a[QP_NMAX_COUNT];
    a[QP_NMAX_TIME] = ~s.time_ns < a[QP_NMAX_TIME] ? ~s.time_ns : a[QP_NMAX_TIME];
    const uint64_t aa = hi * hi, ab = hi * lo, bb = lo * lo;
    a[QP_AA_LO] += aa & 0xFFFFFFFFull;
    a[QP_AA_HI] += aa >> 32;
    a[QP_AB_LO] += ab & 0xFFFFFFFFull;
    a[QP_AB_HI] += ab >> 32;
    a[QP_BB_LO] += bb & 0xFFFFFFFFull;
    a[QP_BB_HI] += bb >> 32;
    const uint64_t cn = s.count & 0xFFFFFFFFull; /* counts are 32-bit accumulators (ST32_COUNT) */
    const uint64_t cc = cn * cn;
    a[QP_CC_LO] += cc & 0xFFFFFFFFull;
    a[QP_CC_HI] += cc >> 32;
    a[QP_LEVEL_SQ] += s.level * s.level;
  }
  __shared__ uint64_t sh[8][QK_PARTIAL_WORDS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < QK_PARTIAL_WORDS; ++k) {
    const bool is_min = k >= QP_MIN_COUNT && k <= QP_NMAX_TIME;
    const uint64_t v = is_min ? warp_min_u64(a[k]) : warp_sum_u64(a[k]);
    if (lane == 0) sh[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < QK_PARTIAL_WORDS) {
    const int k = threadIdx.x, nw = blockDim.x >> 5;
    const bool is_min = k >= QP_MIN_COUNT && k <= QP_NMAX_TIME;
    uint64_t v = is_min ? ~0ull : 0ull;
    for (int w = 0; w < nw; ++w) v = is_min ? (sh[w][k] < v ? sh[w][k] : v) : v + sh[w][k];
    unsigned long long* dst = (unsigned long long*)&part[(size_t)comp * QK_PARTIAL_WORDS + k];
    if (is_min)
      atomicMin(dst, (unsigned long long)v);
    else if (v)
      atomicAdd(dst, (unsigned long long)v);
  }
}

/* the neutral element of the combination: 0 for the sums, ~0 for the four extrema words */
__global__ void qk_partials_clear_kernel(uint64_t* part, uint32_t n_words) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_words) {
    const uint32_t k = i % QK_PARTIAL_WORDS;
    part[i] = (k >= QP_MIN_COUNT && k <= QP_NMAX_TIME) ? ~0ull : 0ull;
  }
}

/* out = combination of n_parts packed partial arrays laid end to end (what an all-gather over the GPUs delivers):
 * unsigned min on the extrema words, wrapping sum on everything else */
__global__ void qk_partials_combine_kernel(const uint64_t* parts, uint32_t n_parts, uint32_t n_words, uint64_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  const uint32_t k = i % QK_PARTIAL_WORDS;
  const bool is_min = k >= QP_MIN_COUNT && k <= QP_NMAX_TIME;
  uint64_t v = is_min ? ~0ull : 0ull;
  for (uint32_t p = 0; p < n_parts; ++p) {
    const uint64_t w = parts[(size_t)p * n_words + i];
    v = is_min ? (w < v ? w : v) : v + w;
  }
  out[i] = v;
}

/* hist[c][0][bin] over time_ns, hist[c][1][bin] over count; ranges from the extrema words of (combined) partials */
__global__ void __launch_bounds__(256) qk_hist_kernel(const uint8_t* tables, const uint64_t* g64, const uint32_t* g32,
                                                      const uint64_t* w64, const uint32_t* w32, uint32_t wa64, uint32_t wa32, const uint64_t* gs64, const uint32_t* gs32, uint64_t rep_pad, uint64_t n_reps, const uint64_t* part,
                                                      uint64_t* hist) {
  const DevHeader* h = reinterpret_cast<const DevHeader*>(tables);
  const uint32_t comp = blockIdx.y;
  const DevComp* c = reinterpret_cast<const DevComp*>(tables + h->off_comps) + comp;
  __shared__ unsigned long long sh[2][QK_HIST_BINS];
  if (threadIdx.x < 2 * QK_HIST_BINS) sh[threadIdx.x / QK_HIST_BINS][threadIdx.x % QK_HIST_BINS] = 0;
  __syncthreads();
  const uint64_t* ext = part + (size_t)comp * QK_PARTIAL_WORDS;
  const uint64_t lo_c = ext[QP_MIN_COUNT], lo_t = ext[QP_MIN_TIME];
  const uint64_t hi_c = ~ext[QP_NMAX_COUNT], hi_t = ~ext[QP_NMAX_TIME];
  const uint64_t w_c = (hi_c - lo_c) / QK_HIST_BINS + 1, w_t = (hi_t - lo_t) / QK_HIST_BINS + 1;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reps; r += (uint64_t)gridDim.x * blockDim.x) {
    qk_comp_stats s;
    qk_comp_stats_dev(c, reinterpret_cast<const DevVec*>(tables + h->off_vecs), comp, g64, g32, w64, w32, wa64, wa32, gs64, gs32, rep_pad, r, &s);
    if (s.time_ns >= lo_t && s.time_ns <= hi_t) atomicAdd(&sh[0][(s.time_ns - lo_t) / w_t], 1ull);
    if (s.count >= lo_c && s.count <= hi_c) atomicAdd(&sh[1][(s.count - lo_c) / w_c], 1ull);
  }
  __syncthreads();
  if (threadIdx.x < 2 * QK_HIST_BINS) {
    const unsigned long long v = sh[threadIdx.x / QK_HIST_BINS][threadIdx.x % QK_HIST_BINS];
    if (v) atomicAdd((unsigned long long*)&hist[(size_t)comp * 2 * QK_HIST_BINS + threadIdx.x], v);
  }
}

/* the retained records of replications [rep0, rep0 + n) in canonical order (oldest first), one block per replication:
 * out[(r - rep0) * log_cap + i], kept[r - rep0] = records retained, total[r - rep0] = records ever written */
__global__ void qk_gather_logs_kernel(const uint4* log, const uint32_t* g32, uint64_t rep_pad, uint64_t rep0,
                                      uint32_t log_cap, uint4* out, uint32_t* kept_out, uint64_t* total_out) {
  const uint64_t r = rep0 + blockIdx.x;
  const uint32_t cnt = g32[(size_t)G32_LOGCNT * rep_pad + r];
  const uint32_t kept = cnt < log_cap ? cnt : log_cap;
  const uint32_t first = cnt - kept;
  const uint4* ring = log + 2 * (size_t)r * log_cap;
  uint4* dst = out + 2 * (size_t)blockIdx.x * log_cap;
  for (uint32_t i = threadIdx.x; i < kept; i += blockDim.x) {
    const uint32_t pos = (first + i) & (log_cap - 1);
    uint4 a = ring[2 * pos], b = ring[2 * pos + 1];
    if (((a.z >> 16) & 0xFFu) != QK_LOG_VEC_DATA) b.x &= 0xFFFFu; /* aux: device-internal, cleared on read */
    dst[2 * i] = a;
    dst[2 * i + 1] = b;
  }
  if (threadIdx.x == 0) {
    kept_out[blockIdx.x] = kept;
    total_out[blockIdx.x] = cnt;
  }
}

__global__ void qk_batch_summary_kernel(const uint64_t* g64, const uint32_t* g32, uint64_t rep_pad, uint64_t n_reps,
                                        uint64_t* out /* n_events, n_faulted, n_log, min_now, ~max_now */) {
  uint64_t ev = 0, fa = 0, lg = 0, mn = ~0ull, mx = ~0ull;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reps; r += (uint64_t)gridDim.x * blockDim.x) {
    ev += g64[(size_t)G64_NEVENTS * rep_pad + r];
    fa += g32[(size_t)G32_STATUS * rep_pad + r] ? 1 : 0;
    lg += g32[(size_t)G32_LOGCNT * rep_pad + r];
    const uint64_t now = g64[(size_t)G64_NOW * rep_pad + r];
    mn = now < mn ? now : mn;
    mx = ~now < mx ? ~now : mx;
  }
  ev = warp_sum_u64(ev);
  fa = warp_sum_u64(fa);
  lg = warp_sum_u64(lg);
  mn = warp_min_u64(mn);
  mx = warp_min_u64(mx);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd((unsigned long long*)&out[0], (unsigned long long)ev);
    atomicAdd((unsigned long long*)&out[1], (unsigned long long)fa);
    atomicAdd((unsigned long long*)&out[2], (unsigned long long)lg);
    atomicMin((unsigned long long*)&out[3], (unsigned long long)mn);
    atomicMin((unsigned long long*)&out[4], (unsigned long long)mx);
  }
}

/* one element of a [slot][replication] plane */
static cudaError_t fetch(const uint32_t* plane, uint64_t rep_pad, uint32_t slot, uint64_t rep, uint32_t* out) {
  return cudaMemcpy(out, plane + (size_t)slot * rep_pad + rep, 4, cudaMemcpyDeviceToHost);
}
static cudaError_t fetch(const uint64_t* plane, uint64_t rep_pad, uint32_t slot, uint64_t rep, uint64_t* out) {
  return cudaMemcpy(out, plane + (size_t)slot * rep_pad + rep, 8, cudaMemcpyDeviceToHost);
}

/* ============================================ C ABI ============================================ */
extern "C" {

/* The small things a batch needs next to its planes -- model tables, tape table, reduction partials, read-back scratch,
 * and a page-locked host buffer -- are ONE device arena and one pinned buffer, and a destroyed batch leaves them in a
 * small per-process cache instead of giving them back to the driver: cudaMalloc, cudaFree and above all cudaHostAlloc
 * are device-wide (and, with one process per GPU, machine-wide) synchronisations of 0.1-1 ms each, which a user who
 * builds one batch after another paid on every batch (end to end at 8 GPUs: 8 ms of `create` on a 66 ms step). */
struct QkSmall {
  int device;
  uint8_t* d_arena;
  size_t arena_bytes;
  uint8_t* h_pinned;
  size_t pinned_bytes;
};
static std::mutex g_small_mu;
static std::vector<QkSmall> g_small;
enum { QK_SMALL_CACHE_MAX = 16 };

static bool small_take(int device, size_t arena_bytes, size_t pinned_bytes, QkSmall* out) {
  std::lock_guard<std::mutex> lk(g_small_mu);
  for (size_t i = 0; i < g_small.size(); ++i)
    if (g_small[i].device == device && g_small[i].arena_bytes >= arena_bytes && g_small[i].pinned_bytes >= pinned_bytes) {
      *out = g_small[i];
      g_small.erase(g_small.begin() + (long)i);
      return true;
    }
  return false;
}
static void small_give(const QkSmall& e) {
  {
    std::lock_guard<std::mutex> lk(g_small_mu);
    if (g_small.size() < QK_SMALL_CACHE_MAX) {
      g_small.push_back(e);
      return;
    }
  }
  cudaFree(e.d_arena);
  cudaFreeHost(e.h_pinned);
}
static size_t align256(size_t x) { return (x + 255u) & ~(size_t)255u; }

/* inside qk_batch_create, once `b` exists: a failed CUDA call (typically cudaMalloc on a batch that does not fit the
 * GPU) releases what was allocated so far */
#define CREATE_TRY(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      qk_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      cudaGetLastError();                                                                        \
      qk_batch_destroy(b);                                                                       \
      return QK_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

int qk_batch_create(const qk_model* m, const qk_batch_config* cfg, qk_batch** out) {
  if (!m || !cfg || !out) {
    qk_set_error("qk_batch_create: NULL argument");
    return QK_ERR_INVALID_ARG;
  }
  if (cfg->n_reps > 0xFFFF0000ull) {
    qk_set_error("qk_batch_create: at most 2^32 - 65536 replications per batch (one GPU); shard across batches");
    return QK_ERR_INVALID_ARG;
  }
  if (cfg->n_reps == 0) {
    qk_set_error("qk_batch_create: n_reps must be > 0");
    return QK_ERR_INVALID_ARG;
  }
  if (cfg->log_capacity && (cfg->log_capacity & (cfg->log_capacity - 1))) {
    qk_set_error("qk_batch_create: log_capacity must be a power of two");
    return QK_ERR_INVALID_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    qk_set_error("qk_batch_create: no CUDA device (this library has no CPU fallback)");
    return QK_ERR_NO_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    qk_set_error("qk_batch_create: device %d out of range (%d devices)", cfg->device, ndev);
    return QK_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(cfg->device));
  qk_batch* b = new qk_batch();
  memset(&b->cfg, 0, sizeof(b->cfg));
  b->cfg = *cfg;
  b->log = cfg->log_capacity != 0;
  b->g64 = nullptr;
  b->g32 = nullptr;
  b->gs64 = nullptr;
  b->gs32 = nullptr;
  b->gc64 = nullptr;
  b->gc32 = nullptr;
  b->d_tables = nullptr;
  b->d_log = nullptr;
  b->d_tapes = nullptr;
  b->d_tape_len = nullptr;
  b->any_tape = false;
  b->stream = nullptr;
  b->own_stream = false;
  b->ev0 = b->ev1 = nullptr;
  b->initialised = false;
  b->last_ms = 0;
  b->last_launches = 0;
  b->total_launches = 0;
  b->d_part = b->d_hist = nullptr;
  b->d_scratch = nullptr;
  b->h_pinned = nullptr;
  b->scratch_bytes = b->pinned_bytes = 0;
  b->d_arena = nullptr;
  b->arena_bytes = 0;
  b->n_sm = 148;
  b->wave_blocks = 0;
  b->lanes = 1;
  b->last_chunks = 1;
  b->grp_reps = b->grp_stride64 = b->grp_stride32 = b->grp_bar_off = b->grp_x64 = 0;
  b->phased = false;
  b->split = false;
  b->use_tma = false;
  b->have_side = false;
  int rc = qk_lower(m, b->log, 0, &b->low);
  if (rc) {
    delete b;
    return rc;
  }
  const DevHeader& h = b->low.hdr; /* (a reference: a second lowering below is seen through it) */
  b->table_bytes_padded = (h.total_bytes + 127u) & ~127u;
  uint32_t per_rep = h.n64 * 8 + h.n32 * 4;
  uint32_t nt = cfg->threads_per_block;
  const uint32_t max_thr = (uint32_t)qk_lb_resident((int)h.features, b->log); /* the kernel variant's launch bounds */
  b->hbm = (cfg->flags & QK_BATCH_STATE_IN_HBM) != 0;
  bool group = (cfg->flags & QK_BATCH_STATE_LANE_GROUP) != 0;
  bool split = (cfg->flags & QK_BATCH_STATE_SPLIT) != 0;
  bool qsplit = split && (cfg->flags & QK_BATCH_QUEUE_TWO_TIER) != 0;
  if (group && h.features == 3) {
    qk_set_error("qk_batch_create: the lane-group kernels do not carry user-defined vector resources (feature set 3)");
    delete b;
    return QK_ERR_INVALID_ARG;
  }
  if (split && (b->hbm || group || (cfg->flags & QK_BATCH_STATE_ON_CHIP))) {
    qk_set_error("qk_batch_create: QK_BATCH_STATE_SPLIT excludes the other placement flags");
    delete b;
    return QK_ERR_INVALID_ARG;
  }
  uint32_t G = cfg->lanes_per_replication;
  if (const char* e = getenv("QK_GROUP_LANES")) /* experiments: overrides the group width wherever groups are used */
    if (atoi(e) > 0) G = (uint32_t)atoi(e);
  if (!b->hbm && !group && !split && nt == 0 && !(cfg->flags & QK_BATCH_STATE_ON_CHIP)) {
    /* placement heuristic (measured, profiles/README.md): one thread per replication with the state staged in shared
     * memory wins while at least 12 warps fit per SM (discrete_queue: 8 x 64 threads); below that a lane group per
     * replication keeps the state on chip with fewer replications and more warps */
    uint32_t thr = 0;
    choose_nt(per_rep, b->table_bytes_padded, max_thr, &thr);
    if (thr < 384) {
      /* ... when the scheduler queue is long enough for the cooperative pop to matter (measured: the 96- and
       * 67-component networks gain 10-45 %, the 10- and 13-component vector / container models lose a factor of two
       * against operating on the L1/L2-cached HBM planes, profiles/README.md) */
      /* (round 2, profiles/README.md) ... and both lose to the split placement -- one thread per replication, with only
       * the scheduler queue and the stocks' words on chip and the component records read in place through L2: 2-3 times the
       * HBM planes on the 96- and 67-component networks, 1.1-1.3 times on the 10- and 13-component vector / container
       * models.  Lean models (feature set 0) with a long queue also move their keyed events off the chip (two-tier queue). */
      const char* e = getenv("QK_AUTO_PLACEMENT"); /* "hbm" / "group" / "split" / "qsplit": force the choice */
      if (e && e[0] == 'h') b->hbm = true;
      else if (e && e[0] == 'g') group = true;
      else if (e && e[0] == 'q') split = qsplit = true;
      else if (e && e[0] == 's') split = true;
      else {
        split = true;
        qsplit = h.features == 0 && h.n_sched >= 32;
      }
    }
  }
  if (split) {
    /* lower again with the component records in the cold planes; what is staged is the scheduler queue and the stocks */
    rc = qk_lower(m, b->log, qsplit ? 2 : 1, &b->low);
    if (rc) {
      delete b;
      return rc;
    }
    b->split = true;
    b->table_bytes_padded = (h.total_bytes + 127u) & ~127u;
    per_rep = h.n64 * 8 + h.n32 * 4;
  }
  bool phased = group && (cfg->flags & QK_BATCH_GROUP_PHASED) != 0;
  if (const char* e = getenv("QK_GROUP_PHASED")) /* experiments: "1" / "0" overrides the variant wherever groups are used */
    phased = group && e[0] == '1';
  uint32_t phased_blocks = 2; /* blocks per SM the phased geometry aims at: they overlap each other's phases */
  if (const char* e = getenv("QK_PHASED_BLOCKS"))
    if (atoi(e) >= 1 && atoi(e) <= 8) phased_blocks = (uint32_t)atoi(e);
  if (group) {
    if (G == 0) G = 8;
    QkGroupGeom geo;
    /* per-block budget: the SM's 228 KB over the blocks, less the 1 KB the system reserves per block and the static part */
    const uint32_t budget = phased ? (228u * 1024u) / phased_blocks - 1024u - 256u : 227u * 1024u - 256u;
    const bool ok = (G == 4 || G == 8 || G == 16) &&
                    qk_group_geometry(G, 32u, h.n64, h.n32, phased ? 2u : 0u, phased ? 6u : 0u, b->table_bytes_padded,
                                      group_max_threads(G), budget < 227u * 1024u - 256u ? budget : 227u * 1024u - 256u,
                                      1024u, 4u, &geo) != 0 &&
                    geo.reps >= 8;
    if (!ok) {
      if (cfg->flags & QK_BATCH_STATE_LANE_GROUP) {
        qk_set_error("qk_batch_create: no lane-group geometry for %u lanes per replication (%u B of state per "
                     "replication, %u B of tables)", G, per_rep, b->table_bytes_padded);
        delete b;
        return QK_ERR_CAPACITY;
      }
      group = false;
      b->hbm = true; /* too large even for a lane group per replication */
    } else {
      /* prefer the largest block that still leaves the SM full: several small blocks recycle capacity sooner, but
       * every block carries its own copy of the tables */
      b->lanes = G;
      b->grp_reps = geo.reps;
      b->grp_stride64 = geo.stride64;
      b->grp_stride32 = geo.stride32;
      b->grp_bar_off = geo.bar_off;
      b->grp_x64 = phased ? 2u : 0u;
      b->phased = phased;
      nt = geo.threads;
    }
  }
  if (group) {
    /* geometry chosen above */
  } else if (!b->hbm) {
    uint32_t thr_cap = max_thr;
    if (split) {
      /* the split kernels take their block size at run time and are compiled without a residency target: what the
       * register file allows is read off the kernel itself */
      cudaFuncAttributes fa;
      CREATE_TRY(cudaFuncGetAttributes(&fa, (const void*)pick_kernel(b->log, false, 0, h.features, h.split)));
      const uint32_t regs = ((uint32_t)(fa.numRegs > 0 ? fa.numRegs : 128) + 7u) & ~7u;
      thr_cap = (65536u / regs) & ~31u;
      /* two-tier queue: its working set (4 KB per resident replication) lives in L2 -- past 12 warps per SM the extra
       * replications cost more in L2 misses than their warps hide (measured: 3 x 128 threads beat 3 x 160 and 2 x 256) */
      uint32_t qcap = 384u;
      if (const char* e = getenv("QK_QSPLIT_THREADS")) /* experiments */
        if (atoi(e) >= 32) qcap = (uint32_t)atoi(e) & ~31u;
      if (qsplit && thr_cap > qcap) thr_cap = qcap;
    }
    /* (two-tier queue, measured on the 96-component network: 3 x 128 and 2 x 192 threads 4.2e9 events/s, 4 x 96 3.8e9,
     * 5 x 64 3.7e9 -- every block carries its own 22 KB copy of the tables) */
    if (nt == 0) nt = choose_nt(per_rep, b->table_bytes_padded, thr_cap, nullptr, qsplit ? 128u : 64u);
    const bool fits = nt != 0 && nt <= QK_MAX_NT && !(nt & 31) &&
                      (uint64_t)b->table_bytes_padded + (uint64_t)per_rep * nt <= 227u * 1024u - 256u;
    if (!fits) {
      if ((cfg->flags & QK_BATCH_STATE_ON_CHIP) || split) {
        qk_set_error("qk_batch_create: model needs %u B of on-chip state per replication (+%u B tables); "
                     "no block size fits the 227 KB of shared memory", per_rep, b->table_bytes_padded);
        delete b;
        return QK_ERR_CAPACITY;
      }
      b->hbm = true; /* too large for one warp per block on chip: operate on the HBM planes in place */
      nt = cfg->threads_per_block;
    }
  }
  if (b->hbm) {
    if (nt == 0) nt = 128;
    if (nt > QK_MAX_NT || (nt & 31) || b->table_bytes_padded > 200u * 1024u) {
      qk_set_error("qk_batch_create: invalid threads_per_block %u or tables too large (%u B)", nt,
                   b->table_bytes_padded);
      delete b;
      return QK_ERR_CAPACITY;
    }
  }
  b->nt = nt;
  const uint32_t reps_per_block = group ? b->grp_reps : nt;
  b->smem_bytes = group ? b->grp_bar_off + 16u : b->table_bytes_padded + (b->hbm ? 0u : per_rep * nt);
  if (const char* pad = getenv("QK_DEBUG_EXTRA_SMEM")) /* occupancy experiments only: lowers resident blocks per SM */
    b->smem_bytes += (uint32_t)atoi(pad);
  b->grid = (uint32_t)((cfg->n_reps + reps_per_block - 1) / reps_per_block);
  b->rep_pad = (uint64_t)b->grid * reps_per_block;
  if (cfg->stream) {
    b->stream = (cudaStream_t)cfg->stream;
  } else {
    CREATE_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    b->own_stream = true;
  }
  CREATE_TRY(cudaEventCreate(&b->ev0));
  CREATE_TRY(cudaEventCreate(&b->ev1));
  /* The planes and the log rings (gigabytes) come from the device's stream-ordered memory pool, which keeps what a
   * destroyed batch hands back: cudaMalloc / cudaFree of such buffers cost tens to hundreds of milliseconds each
   * (page mapping, device-wide synchronisation), which is what a user who builds one batch after another would wait
   * for -- and what made the end-to-end leg of the bench noisy. */
  {
    cudaMemPool_t pool;
    CREATE_TRY(cudaDeviceGetDefaultMemPool(&pool, cfg->device));
    uint64_t keep = ~0ull;
    CREATE_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  CREATE_TRY(cudaMallocAsync(&b->g64, (size_t)h.n64 * b->rep_pad * 8, b->stream));
  CREATE_TRY(cudaMallocAsync(&b->g32, (size_t)h.n32 * b->rep_pad * 4, b->stream));
  CREATE_TRY(cudaMallocAsync(&b->gs64, (size_t)h.n_comp * ST_PER_COMP * b->rep_pad * 8, b->stream));
  CREATE_TRY(cudaMallocAsync(&b->gs32, (size_t)h.n_comp * ST_PER_COMP * b->rep_pad * 4, b->stream));
  CREATE_TRY(cudaMallocAsync(&b->gc64, (size_t)(h.n64c ? h.n64c : 1) * b->rep_pad * 8, b->stream));
  CREATE_TRY(cudaMallocAsync(&b->gc32, (size_t)(h.n32c ? h.n32c : 1) * b->rep_pad * 4, b->stream));
  if (b->log) CREATE_TRY(cudaMallocAsync(&b->d_log, (size_t)b->rep_pad * cfg->log_capacity * 32, b->stream));
  CREATE_TRY(cudaStreamSynchronize(b->stream)); /* allocation failures of the pool surface here at the latest */
  b->use_tma = !b->hbm && !group && make_tensor_maps(b);
  {
    /* the arena: tables | tape pointers | tape lengths | partials | histograms | scratch, 256-byte aligned pieces */
    const size_t n_dist = h.n_dist ? h.n_dist : 1;
    b->scratch_bytes = std::max<size_t>(1u << 20, (size_t)h.n_comp * (QK_PARTIAL_WORDS + 2 * QK_HIST_BINS) * 8);
    b->pinned_bytes = b->scratch_bytes;
    const size_t o_tables = 0, o_tapes = o_tables + align256(b->table_bytes_padded);
    const size_t o_len = o_tapes + align256(sizeof(double*) * n_dist), o_part = o_len + align256(sizeof(uint64_t) * n_dist);
    const size_t o_hist = o_part + align256(sizeof(uint64_t) * QK_PARTIAL_WORDS * h.n_comp);
    const size_t o_scratch = o_hist + align256(sizeof(uint64_t) * 2 * QK_HIST_BINS * h.n_comp);
    const size_t need = o_scratch + align256(b->scratch_bytes);
    QkSmall e;
    if (small_take(cfg->device, need, b->pinned_bytes, &e)) {
      b->d_arena = e.d_arena;
      b->arena_bytes = e.arena_bytes;
      b->h_pinned = e.h_pinned;
      b->pinned_bytes = e.pinned_bytes;
    } else {
      CREATE_TRY(cudaMalloc(&b->d_arena, need));
      b->arena_bytes = need;
      CREATE_TRY(cudaHostAlloc(&b->h_pinned, b->pinned_bytes, cudaHostAllocDefault));
    }
    b->d_tables = b->d_arena + o_tables;
    b->d_tapes = reinterpret_cast<const double**>(b->d_arena + o_tapes);
    b->d_tape_len = reinterpret_cast<uint64_t*>(b->d_arena + o_len);
    b->d_part = reinterpret_cast<uint64_t*>(b->d_arena + o_part);
    b->d_hist = reinterpret_cast<uint64_t*>(b->d_arena + o_hist);
    b->d_scratch = b->d_arena + o_scratch;
    /* tables (zero-padded to the 128-byte multiple the kernels copy) and an empty tape table, in stream order: the
     * launches that read them follow on the same stream.  (The source is pageable: the copy is staged before the call
     * returns.) */
    CREATE_TRY(cudaMemsetAsync(b->d_arena, 0, o_part, b->stream));
    CREATE_TRY(cudaMemcpyAsync(b->d_tables, b->low.blob.data(), b->low.blob.size(), cudaMemcpyHostToDevice, b->stream));
  }
  b->h_tapes.assign(h.n_dist, nullptr);
  b->h_tape_len.assign(h.n_dist, 0);
  b->kernel = group ? (b->phased ? pick_phased_kernel(b->log, b->lanes, h.features) : pick_group_kernel(b->log, b->lanes, h.features))
                    : pick_kernel(b->log, b->hbm, b->nt, h.features, h.split);
  {
    /* the opt-in limit covers static + dynamic shared memory: the kernels keep an mbarrier (and little else) in static */
    cudaFuncAttributes fa;
    CREATE_TRY(cudaFuncGetAttributes(&fa, (const void*)b->kernel));
    CREATE_TRY(cudaFuncSetAttribute((const void*)b->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024 - (int)fa.sharedSizeBytes));
    if (b->smem_bytes + fa.sharedSizeBytes > 227u * 1024u) {
      qk_set_error("qk_batch_create: %u B of dynamic + %zu B of static shared memory exceed the 227 KB of an SM",
                   b->smem_bytes, fa.sharedSizeBytes);
      qk_batch_destroy(b);
      return QK_ERR_CAPACITY;
    }
  }
  {
    int sms = 0, nb = 0;
    CREATE_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
    CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)b->kernel, (int)b->nt, b->smem_bytes));
    b->n_sm = (uint32_t)sms;
    b->wave_blocks = (uint32_t)(sms * nb);
  }
  b->red_blocks = b->n_sm * 4;
  *out = b;
  return QK_OK;
}

int qk_batch_info_get(qk_batch* b, qk_batch_info* out) {
  if (!b || !out) return QK_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  memset(out, 0, sizeof(*out));
  out->threads_per_block = b->nt;
  out->grid_blocks = b->grid;
  out->shared_bytes_per_block = b->smem_bytes;
  out->table_bytes = b->table_bytes_padded;
  out->slots64 = b->low.hdr.n64;
  out->slots32 = b->low.hdr.n32;
  out->state_bytes_per_rep = b->low.hdr.n64 * 8 + b->low.hdr.n32 * 4;
  out->cold_bytes_per_rep = b->low.hdr.n64c * 8 + b->low.hdr.n32c * 4;
  int nb = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)b->kernel, (int)b->nt, b->smem_bytes));
  out->state_in_hbm = b->hbm ? 1u : 0u;
  out->state_split = b->split ? 1u : 0u;
  out->resident_blocks_per_sm = (uint32_t)nb;
  out->replications_padded = b->rep_pad;
  out->state_bytes_total = (uint64_t)(out->state_bytes_per_rep + out->cold_bytes_per_rep) * b->rep_pad;
  out->log_bytes_total = b->log ? (uint64_t)b->rep_pad * b->cfg.log_capacity * 32 : 0;
  out->lanes_per_replication = b->lanes;
  out->wave_blocks = b->wave_blocks;
  out->last_step_chunks = b->last_chunks;
  return QK_OK;
}

void qk_batch_destroy(qk_batch* b) {
  if (!b) return;
  cudaSetDevice(b->cfg.device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  if (b->stream) { /* back to the pool, in stream order */
    void* pooled[] = {b->g64, b->g32, b->gs64, b->gs32, b->gc64, b->gc32, b->d_log};
    for (void* p : pooled)
      if (p) cudaFreeAsync(p, b->stream);
    cudaStreamSynchronize(b->stream);
  }
  for (size_t i = 0; i < b->h_tapes.size(); ++i) cudaFree(b->h_tapes[i]);
  if (b->have_side) {
    for (int i = 0; i < QK_SLICE_PARTS - 1; ++i) {
      cudaStreamDestroy(b->side[i]);
      cudaEventDestroy(b->ev_join[i]);
    }
    cudaEventDestroy(b->ev_fork);
  }
  if (b->d_arena && b->h_pinned) { /* (the stream was synchronised above: nothing in flight touches them) */
    QkSmall e = {b->cfg.device, b->d_arena, b->arena_bytes, b->h_pinned, b->pinned_bytes};
    small_give(e);
  } else {
    if (b->d_arena) cudaFree(b->d_arena);
    if (b->h_pinned) cudaFreeHost(b->h_pinned);
  }
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->own_stream) cudaStreamDestroy(b->stream);
  delete b;
}

int qk_batch_set_tape(qk_batch* b, uint32_t dist_id, const double* values, uint64_t per_rep_len) {
  if (!b || dist_id >= b->low.hdr.n_dist || (!values && per_rep_len)) {
    qk_set_error("qk_batch_set_tape: invalid argument");
    return QK_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->stream)); /* launches in flight read the tape table this call rewrites */
  if (b->h_tapes[dist_id]) {
    CUDA_TRY(cudaFree(b->h_tapes[dist_id]));
    b->h_tapes[dist_id] = nullptr;
  }
  const size_t n = (size_t)per_rep_len * b->cfg.n_reps;
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, (n ? n : 1) * sizeof(double)));
  if (n) CUDA_TRY(cudaMemcpy(d, values, n * sizeof(double), cudaMemcpyHostToDevice));
  b->h_tapes[dist_id] = d;
  b->h_tape_len[dist_id] = per_rep_len;
  b->any_tape = true;
  CUDA_TRY(cudaMemcpy(b->d_tapes, b->h_tapes.data(), sizeof(double*) * b->h_tapes.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(b->d_tape_len, b->h_tape_len.data(), sizeof(uint64_t) * b->h_tape_len.size(),
                      cudaMemcpyHostToDevice));
  return QK_OK;
}

int qk_batch_init(qk_batch* b, uint64_t t0_ns) {
  if (!b) return QK_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  int rc = launch_step(b, 1, t0_ns, QK_TIME_NONE, 0);
  if (rc) return rc;
  b->initialised = true;
  return QK_OK;
}

int qk_batch_step_until(qk_batch* b, uint64_t t_ns) {
  if (!b) return QK_ERR_INVALID_ARG;
  if (!b->initialised) {
    qk_set_error("qk_batch_step_until: call qk_batch_init first");
    return QK_ERR_STATE;
  }
  if (t_ns == QK_TIME_NONE) return QK_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  return launch_step(b, 0, 0, t_ns, ~0ull);
}

/* Sliced stepping.  A launch of `grid` blocks runs in ceil(grid / wave_blocks) waves, and the last, partial wave runs
 * on a mostly idle GPU at the pace of a full one: 2048 blocks over 1480 resident ones (1M replications of
 * discrete_queue split over 8 GPUs) take two wave times for 1.38 waves of work.  Replications are independent and a
 * step of n events can be cut anywhere (qk_batch_step_events_chunked), so the step is cut into `c` chunks of events and
 * the tiles into QK_SLICE_PARTS parts; each part is advanced chunk by chunk by launches on its OWN stream.  Launches of
 * different parts overlap -- when one part's launch drains, blocks of the others fill the SMs -- while stream order
 * keeps a part's chunk k + 1 behind its chunk k.  The GPU stays full until the last chunk of the last parts: a tail of
 * about 1 / (waves * c) of the step instead of the last wave's.  Costs one state round trip through HBM per chunk
 * (2 * S bytes per replication per >= 200 events) and c * QK_SLICE_PARTS launches. */
static bool plan_slices(const qk_batch* b, uint64_t n_events, uint32_t* n_chunks) {
  static const char* env = getenv("QK_SLICE"); /* "0": never ; "1": whenever more than one wave ; default: when it pays */
  if (env && env[0] == '0') return false;
  const uint64_t W = b->wave_blocks, T = b->grid;
  if (W == 0 || T <= W || T < 4 * QK_SLICE_PARTS || n_events < 400 || n_events > 0xFFFFFFFEull) return false;
  const double waves = (double)T / (double)W;
  const double plain_eff = waves / (double)((T + W - 1) / W);
  if (!(env && env[0] == '1') && plain_eff > 0.97) return false;
  /* the tail is about one chunk of one wave: 1 / (waves * c) of the step; aim at 2 % */
  uint64_t c = (uint64_t)(1.0 / (0.02 * waves)) + 1;
  if (c < 2) c = 2;
  if (c > 32) c = 32;
  while (c > 2 && n_events / c < 200) --c;
  *n_chunks = (uint32_t)c;
  return true;
}

static int step_sliced(qk_batch* b, uint64_t n_events, uint32_t n_chunks) {
  if (!b->have_side) {
    CUDA_TRY(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < QK_SLICE_PARTS - 1; ++i) {
      CUDA_TRY(cudaStreamCreateWithFlags(&b->side[i], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&b->ev_join[i], cudaEventDisableTiming));
    }
    b->have_side = true;
  }
  const uint64_t chunk = (n_events + n_chunks - 1) / n_chunks;
  const uint32_t parts = b->grid >= QK_SLICE_PARTS ? QK_SLICE_PARTS : 1;
  b->last_chunks = (uint32_t)((n_events + chunk - 1) / chunk);
  CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
  b->last_launches = 0;
  CUDA_TRY(cudaEventRecord(b->ev_fork, b->stream));
  for (uint32_t p = 1; p < parts; ++p) CUDA_TRY(cudaStreamWaitEvent(b->side[p - 1], b->ev_fork, 0));
  for (uint64_t done = 0; done < n_events; done += chunk) {
    const uint64_t k = std::min<uint64_t>(chunk, n_events - done);
    for (uint32_t p = 0; p < parts; ++p) {
      QkSlice sl;
      sl.tile0 = (uint32_t)((uint64_t)b->grid * p / parts);
      sl.n_tiles = (uint32_t)((uint64_t)b->grid * (p + 1) / parts) - sl.tile0;
      sl.stream = p == 0 ? b->stream : b->side[p - 1];
      const int rc = launch_step(b, 0, 0, QK_TIME_NONE, k, false, false, &sl);
      if (rc) return rc;
    }
  }
  for (uint32_t p = 1; p < parts; ++p) {
    CUDA_TRY(cudaEventRecord(b->ev_join[p - 1], b->side[p - 1]));
    CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_join[p - 1], 0));
  }
  CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
  return QK_OK;
}

int qk_batch_step_events(qk_batch* b, uint64_t n_events) {
  if (!b) return QK_ERR_INVALID_ARG;
  if (!b->initialised) {
    qk_set_error("qk_batch_step_events: call qk_batch_init first");
    return QK_ERR_STATE;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  uint32_t n_chunks = 0;
  if (plan_slices(b, n_events, &n_chunks)) return step_sliced(b, n_events, n_chunks);
  b->last_chunks = 1;
  /* the kernel counts events in 32 bits: finite budgets go in launches of at most 2^32 - 2 events */
  const uint64_t cap = 0xFFFFFFFEull;
  bool first = true;
  do {
    const uint64_t k = n_events < cap ? n_events : cap;
    n_events -= k;
    const int rc = launch_step(b, 0, 0, QK_TIME_NONE, k, first, n_events == 0);
    if (rc) return rc;
    first = false;
  } while (n_events);
  return QK_OK;
}

/* the sliced schedule with an explicit number of chunks (tests, measurements); n_chunks == 0: as step_events */
int qk_batch_step_events_sliced(qk_batch* b, uint64_t n_events, uint32_t n_chunks) {
  if (!b) return QK_ERR_INVALID_ARG;
  if (n_chunks == 0) return qk_batch_step_events(b, n_events);
  if (!b->initialised) {
    qk_set_error("qk_batch_step_events_sliced: call qk_batch_init first");
    return QK_ERR_STATE;
  }
  if (n_events == 0 || n_events > 0xFFFFFFFEull || n_chunks > n_events) {
    qk_set_error("qk_batch_step_events_sliced: invalid argument");
    return QK_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  return step_sliced(b, n_events, n_chunks);
}

/* The same step cut into launches of `chunk` events each: every launch loads the hot state planes from HBM, advances
 * each replication by at most `chunk` scheduler actions and stores the planes back (events advanced per state
 * residency K = chunk; K = 1 is pure lockstep).  Results are identical to one launch (tests); this entry point exists
 * to MEASURE the kernel where it is HBM-bound (bench.py --events-per-launch). */
int qk_batch_step_events_chunked(qk_batch* b, uint64_t n_events, uint64_t chunk) {
  if (!b || chunk == 0 || chunk > 0xFFFFFFFEull) return QK_ERR_INVALID_ARG;
  if (!b->initialised) {
    qk_set_error("qk_batch_step_events_chunked: call qk_batch_init first");
    return QK_ERR_STATE;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  uint64_t left = n_events;
  bool first = true;
  while (left) {
    const uint64_t k = left < chunk ? left : chunk;
    left -= k;
    const int rc = launch_step(b, 0, 0, QK_TIME_NONE, k, first, left == 0);
    if (rc) return rc;
    first = false;
  }
  return QK_OK;
}

int qk_batch_synchronize(qk_batch* b) {
  if (!b) return QK_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return QK_OK;
}

int qk_batch_last_step_ms(qk_batch* b, float* ms, uint32_t* n_launches) {
  if (!b) return QK_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaEventSynchronize(b->ev1));
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, b->ev0, b->ev1));
  if (ms) *ms = t;
  if (n_launches) *n_launches = b->last_launches;
  return QK_OK;
}

int qk_batch_summary_get(qk_batch* b, qk_batch_summary* out) {
  if (!b || !out) return QK_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  uint64_t* h = reinterpret_cast<uint64_t*>(b->h_pinned);
  uint64_t* d = reinterpret_cast<uint64_t*>(b->d_scratch);
  h[0] = h[1] = h[2] = 0;
  h[3] = h[4] = ~0ull;
  CUDA_TRY(cudaMemcpyAsync(d, h, 5 * sizeof(uint64_t), cudaMemcpyHostToDevice, b->stream));
  qk_batch_summary_kernel<<<b->n_sm, 256, 0, b->stream>>>(b->g64, b->g32, b->rep_pad, b->cfg.n_reps, d);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(h, d, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  out->n_reps = b->cfg.n_reps;
  out->n_events = h[0];
  out->n_faulted = h[1];
  out->n_log_records = h[2];
  out->min_now_ns = h[3];
  out->max_now_ns = ~h[4];
  return QK_OK;
}

static int check_range(qk_batch* b, uint64_t rep0, uint64_t n) {
  if (!b || rep0 + n > b->cfg.n_reps || rep0 + n < rep0) {
    qk_set_error("replication range [%llu, +%llu) out of bounds", (unsigned long long)rep0, (unsigned long long)n);
    return QK_ERR_INVALID_ARG;
  }
  return QK_OK;
}

int qk_batch_read_status(qk_batch* b, uint64_t rep0, uint64_t n, uint32_t* status, uint64_t* now_ns,
                         uint64_t* n_events) {
  int rc = check_range(b, rep0, n);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  if (status)
    CUDA_TRY(cudaMemcpy(status, b->g32 + (size_t)G32_STATUS * b->rep_pad + rep0, n * 4, cudaMemcpyDeviceToHost));
  if (now_ns)
    CUDA_TRY(cudaMemcpy(now_ns, b->g64 + (size_t)G64_NOW * b->rep_pad + rep0, n * 8, cudaMemcpyDeviceToHost));
  if (n_events)
    CUDA_TRY(cudaMemcpy(n_events, b->g64 + (size_t)G64_NEVENTS * b->rep_pad + rep0, n * 8, cudaMemcpyDeviceToHost));
  return QK_OK;
}

int qk_batch_comp_stats(qk_batch* b, uint64_t rep0, uint64_t n, uint32_t comp, qk_comp_stats* out) {
  int rc = check_range(b, rep0, n);
  if (rc) return rc;
  if (comp >= b->low.hdr.n_comp || !out) return QK_ERR_INVALID_ARG;
  if (n == 0) return QK_OK;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  qk_comp_stats* d = reinterpret_cast<qk_comp_stats*>(b->d_scratch);
  const uint64_t piece = b->scratch_bytes / sizeof(qk_comp_stats);
  for (uint64_t done = 0; done < n; done += piece) {
    const uint64_t k = std::min<uint64_t>(piece, n - done);
    qk_comp_stats_kernel<<<(unsigned)((k + 255) / 256), 256, 0, b->stream>>>(b->d_tables, b->g64, b->g32, warm64(b), warm32(b), aos64(b), aos32(b), b->gs64, b->gs32,
                                                                              b->rep_pad, rep0 + done, k, comp, d);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out + done, d, k * sizeof(qk_comp_stats), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
  }
  return QK_OK;
}

int qk_batch_reduce_partials_device(qk_batch* b, uint64_t* d_partials, uint32_t n_comp) {
  if (!b || !d_partials || n_comp != b->low.hdr.n_comp) {
    qk_set_error("qk_batch_reduce_partials_device: invalid argument (n_comp must equal the model's)");
    return QK_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  const uint32_t n_words = n_comp * QK_PARTIAL_WORDS;
  qk_partials_clear_kernel<<<(n_words + 255) / 256, 256, 0, b->stream>>>(d_partials, n_words);
  CUDA_TRY(cudaGetLastError());
  dim3 grid(b->red_blocks, n_comp);
  qk_reduce_kernel<<<grid, 256, 0, b->stream>>>(b->d_tables, b->g64, b->g32, warm64(b), warm32(b), aos64(b), aos32(b), b->gs64, b->gs32, b->rep_pad, b->cfg.n_reps,
                                                d_partials);
  CUDA_TRY(cudaGetLastError());
  return QK_OK;
}

int qk_batch_combine_partials_device(qk_batch* b, const uint64_t* d_parts, uint32_t n_parts, uint64_t* d_out,
                                     uint32_t n_comp) {
  if (!b || !d_parts || !d_out || !n_parts || n_comp != b->low.hdr.n_comp) {
    qk_set_error("qk_batch_combine_partials_device: invalid argument");
    return QK_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  const uint32_t n_words = n_comp * QK_PARTIAL_WORDS;
  qk_partials_combine_kernel<<<(n_words + 255) / 256, 256, 0, b->stream>>>(d_parts, n_parts, n_words, d_out);
  CUDA_TRY(cudaGetLastError());
  return QK_OK;
}

int qk_batch_histograms_device(qk_batch* b, const uint64_t* d_partials, uint64_t* d_hist, uint32_t n_comp) {
  if (!b || !d_partials || !d_hist || n_comp != b->low.hdr.n_comp) return QK_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(uint64_t) * 2 * QK_HIST_BINS * n_comp, b->stream));
  dim3 grid(b->red_blocks, n_comp);
  qk_hist_kernel<<<grid, 256, 0, b->stream>>>(b->d_tables, b->g64, b->g32, warm64(b), warm32(b), aos64(b), aos32(b), b->gs64, b->gs32, b->rep_pad, b->cfg.n_reps,
                                              d_partials, d_hist);
  CUDA_TRY(cudaGetLastError());
  return QK_OK;
}

} /* extern "C" */

/* exact sum of squares from its six half-sums: S = AA * 2^64 + 2 * AB * 2^32 + BB as 32-bit digits with carries */
static void qk_sumsq_limbs(const uint64_t* w, uint64_t limbs[3]) {
  unsigned __int128 d[7] = {0, 0, 0, 0, 0, 0, 0};
  d[0] += w[QP_BB_LO];
  d[1] += w[QP_BB_HI];
  d[1] += (unsigned __int128)w[QP_AB_LO] * 2;
  d[2] += (unsigned __int128)w[QP_AB_HI] * 2;
  d[2] += w[QP_AA_LO];
  d[3] += w[QP_AA_HI];
  for (int k = 0; k < 6; ++k) {
    d[k + 1] += d[k] >> 32;
    d[k] &= 0xFFFFFFFFull;
  }
  for (int k = 0; k < 3; ++k) limbs[k] = (uint64_t)d[2 * k] | ((uint64_t)d[2 * k + 1] << 32);
}

void qk_fill_comp_summary(const uint64_t* w, const uint64_t* hist, qk_comp_summary* out) {
  qk_comp_summary& o = *out;
  memset(&o, 0, sizeof(o));
  o.sum_count = w[QP_COUNT];
  const uint64_t lo = w[QP_TIME_LO], hi = w[QP_TIME_HI];
  /* the kernel sums the low and high 32-bit halves separately; recombine into a 128-bit total */
  const uint64_t low64 = (hi << 32) + lo;
  o.sum_time_lo = low64;
  o.sum_time_hi = (hi >> 32) + (low64 < lo ? 1 : 0);
  o.sum_level = w[QP_LEVEL];
  o.min_count = w[QP_MIN_COUNT];
  o.min_time = w[QP_MIN_TIME];
  o.max_count = ~w[QP_NMAX_COUNT];
  o.max_time = ~w[QP_NMAX_TIME];
  qk_sumsq_limbs(w, o.sumsq_time_ns2);
  const unsigned __int128 cc = ((unsigned __int128)w[QP_CC_HI] << 32) + w[QP_CC_LO];
  o.sumsq_count_lo = (uint64_t)cc;
  o.sumsq_count_hi = (uint64_t)(cc >> 64);
  o.sumsq_level = w[QP_LEVEL_SQ];
  o.sumsq_time = ((long double)o.sumsq_time_ns2[2] * 18446744073709551616.0L * 18446744073709551616.0L +
                  (long double)o.sumsq_time_ns2[1] * 18446744073709551616.0L + (long double)o.sumsq_time_ns2[0]) *
                 1e-18L;
  o.sumsq_count = (double)((long double)o.sumsq_count_hi * 18446744073709551616.0L + (long double)o.sumsq_count_lo);
  if (hist) {
    for (int k = 0; k < QK_HIST_BINS; ++k) {
      o.hist_time[k] = hist[k];
      o.hist_count[k] = hist[QK_HIST_BINS + k];
    }
  }
  o.hist_lo_time = o.min_time;
  o.hist_hi_time = o.max_time;
  o.hist_lo_count = o.min_count;
  o.hist_hi_count = o.max_count;
}

extern "C" {

int qk_batch_reduce_partials(qk_batch* b, uint64_t* partials, uint32_t n_comp) {
  if (!b || !partials || n_comp != b->low.hdr.n_comp) return QK_ERR_INVALID_ARG;
  int rc = qk_batch_reduce_partials_device(b, b->d_part, n_comp);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(b->h_pinned, b->d_part, sizeof(uint64_t) * QK_PARTIAL_WORDS * n_comp, cudaMemcpyDeviceToHost,
                           b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  memcpy(partials, b->h_pinned, sizeof(uint64_t) * QK_PARTIAL_WORDS * n_comp);
  return QK_OK;
}

int qk_batch_reduce_stats(qk_batch* b, qk_comp_summary* out, uint32_t n_comp) {
  if (!b || !out || n_comp != b->low.hdr.n_comp) return QK_ERR_INVALID_ARG;
  int rc = qk_batch_reduce_partials_device(b, b->d_part, n_comp);
  if (rc) return rc;
  rc = qk_batch_histograms_device(b, b->d_part, b->d_hist, n_comp);
  if (rc) return rc;
  uint64_t* hp = reinterpret_cast<uint64_t*>(b->h_pinned);
  const size_t n_part = (size_t)QK_PARTIAL_WORDS * n_comp, n_hist = (size_t)2 * QK_HIST_BINS * n_comp;
  CUDA_TRY(cudaMemcpyAsync(hp, b->d_part, n_part * 8, cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaMemcpyAsync(hp + n_part, b->d_hist, n_hist * 8, cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  for (uint32_t c = 0; c < n_comp; ++c)
    qk_fill_comp_summary(hp + (size_t)c * QK_PARTIAL_WORDS, hp + n_part + (size_t)c * 2 * QK_HIST_BINS, &out[c]);
  return QK_OK;
}

int qk_batch_read_logs(qk_batch* b, uint64_t rep0, uint64_t n, qk_log_record* buf, uint32_t* kept, uint64_t* total) {
  int rc = check_range(b, rep0, n);
  if (rc) return rc;
  if (!b->log) {
    qk_set_error("qk_batch_read_logs: batch was created with log_capacity 0");
    return QK_ERR_STATE;
  }
  if (!buf || !kept || !total) return QK_ERR_INVALID_ARG;
  if (n == 0) return QK_OK;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  const uint32_t lc = b->cfg.log_capacity;
  const size_t per_rep = (size_t)lc * 32 + 16; /* records + (kept, total) */
  uint64_t piece = b->scratch_bytes / per_rep;
  uint8_t* d = b->d_scratch;
  bool pooled = false;
  if (piece < n && piece < 4096) { /* large requests: one temporary from the stream-ordered pool instead of many pieces */
    piece = std::min<uint64_t>(n, (256u << 20) / per_rep);
    if (piece == 0) piece = 1;
    CUDA_TRY(cudaMallocAsync(&d, piece * per_rep, b->stream));
    pooled = true;
  }
  for (uint64_t done = 0; done < n; done += piece) {
    const uint64_t k = std::min<uint64_t>(piece, n - done);
    uint4* d_rec = reinterpret_cast<uint4*>(d);
    uint64_t* d_total = reinterpret_cast<uint64_t*>(d + k * (size_t)lc * 32);
    uint32_t* d_kept = reinterpret_cast<uint32_t*>(d_total + k);
    qk_gather_logs_kernel<<<(unsigned)k, 64, 0, b->stream>>>(b->d_log, b->g32, b->rep_pad, rep0 + done, lc, d_rec, d_kept,
                                                             d_total);
    CUDA_TRY(cudaGetLastError());
    /* one gather kernel and one copy per array instead of a round trip per replication */
    CUDA_TRY(cudaMemcpyAsync(buf + done * lc, d_rec, k * (size_t)lc * 32, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaMemcpyAsync(total + done, d_total, k * 8, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaMemcpyAsync(kept + done, d_kept, k * 4, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
  }
  if (pooled) CUDA_TRY(cudaFreeAsync(d, b->stream));
  return QK_OK;
}

int qk_batch_read_log(qk_batch* b, uint64_t rep, qk_log_record* buf, uint64_t cap, uint64_t* n_out,
                      uint64_t* n_total) {
  int rc = check_range(b, rep, 1);
  if (rc) return rc;
  if (!b->log) {
    qk_set_error("qk_batch_read_log: batch was created with log_capacity 0");
    return QK_ERR_STATE;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  uint32_t cnt = 0;
  CUDA_TRY(cudaMemcpy(&cnt, b->g32 + (size_t)G32_LOGCNT * b->rep_pad + rep, 4, cudaMemcpyDeviceToHost));
  const uint32_t lc = b->cfg.log_capacity;
  const uint64_t kept = std::min<uint64_t>(cnt, lc);
  if (n_total) *n_total = cnt;
  if (n_out) *n_out = std::min<uint64_t>(kept, cap);
  if (!buf || cap == 0) return QK_OK;
  if (cap < kept) {
    qk_set_error("qk_batch_read_log: buffer holds %llu records, %llu needed", (unsigned long long)cap,
                 (unsigned long long)kept);
    return QK_ERR_BUFFER_TOO_SMALL;
  }
  std::vector<qk_log_record> ring(lc);
  CUDA_TRY(cudaMemcpy(ring.data(), reinterpret_cast<uint8_t*>(b->d_log) + (size_t)rep * lc * 32, (size_t)lc * 32,
                      cudaMemcpyDeviceToHost));
  const uint64_t first = cnt - kept; /* oldest record still in the ring */
  for (uint64_t i = 0; i < kept; ++i) {
    buf[i] = ring[(first + i) & (lc - 1)];
    if (buf[i].kind != QK_LOG_VEC_DATA) buf[i].aux = 0; /* continuation records carry fp64 bits there */
  }
  return QK_OK;
}

int qk_batch_read_vector(qk_batch* b, uint64_t rep, uint32_t comp, double* out3) {
  int rc = check_range(b, rep, 1);
  if (rc) return rc;
  if (comp >= b->low.hdr.n_comp || b->low.comps[comp].kind < QK_VECTOR_STOCK ||
      b->low.comps[comp].kind > QK_VECTOR_SPLITTER || !out3) {
    qk_set_error("qk_batch_read_vector: component %u is not a vector component", comp);
    return QK_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  const DevVec& v = b->low.vecs[b->low.comps[comp].vec_idx];
  out3[0] = out3[1] = out3[2] = 0.0;
  for (uint32_t k = 0; k < v.dim && k < 3; ++k)
    CUDA_TRY(cudaMemcpy(&out3[k], warm64_at(b, v.o64_res + k, rep), 8, cudaMemcpyDeviceToHost));
  return QK_OK;
}

int qk_batch_read_vector_n(qk_batch* b, uint64_t rep, uint32_t comp, double* out, uint32_t cap, uint32_t* dim_out) {
  int rc = check_range(b, rep, 1);
  if (rc) return rc;
  if (comp >= b->low.hdr.n_comp || b->low.comps[comp].kind < QK_VECTOR_STOCK ||
      b->low.comps[comp].kind > QK_VECTOR_SPLITTER || !out) {
    qk_set_error("qk_batch_read_vector_n: component %u is not a vector component", comp);
    return QK_ERR_INVALID_ARG;
  }
  const DevVec& v = b->low.vecs[b->low.comps[comp].vec_idx];
  if (dim_out) *dim_out = v.dim;
  if (cap < v.dim) return QK_ERR_BUFFER_TOO_SMALL;
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  for (uint32_t k = 0; k < v.dim; ++k)
    CUDA_TRY(cudaMemcpy(&out[k], warm64_at(b, v.o64_res + k, rep), 8, cudaMemcpyDeviceToHost));
  return QK_OK;
}

int qk_batch_read_stock(qk_batch* b, uint64_t rep, uint32_t comp, uint32_t* items, uint64_t cap, uint64_t* n_out) {
  int rc = check_range(b, rep, 1);
  if (rc) return rc;
  if (comp >= b->low.hdr.n_comp || b->low.comps[comp].kind != QK_DISCRETE_STOCK) {
    qk_set_error("qk_batch_read_stock: component %u is not a discrete stock", comp);
    return QK_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  const DevComp& c = b->low.comps[comp];
  uint32_t hc = 0;
  CUDA_TRY(cudaMemcpy(&hc, b->g32 + (size_t)(c.o32 + S32_HC) * b->rep_pad + rep, 4, cudaMemcpyDeviceToHost));
  const uint32_t head = QK_HC_HEAD(hc), cnt = QK_HC_CNT(hc);
  if (n_out) *n_out = cnt;
  if (!items) return QK_OK;
  if (cap < cnt) return QK_ERR_BUFFER_TOO_SMALL;
  std::vector<uint32_t> ring(c.ring_cap);
  if (b->split)
    CUDA_TRY(cudaMemcpy(ring.data(), cold32_at(b, c.ocold32, rep), 4u * c.ring_cap, cudaMemcpyDeviceToHost));
  else
    CUDA_TRY(cudaMemcpy2D(ring.data(), 4, b->gc32 + (size_t)c.ocold32 * b->rep_pad + rep, b->rep_pad * 4, 4,
                          c.ring_cap, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < cnt; ++i) items[i] = ring[(head + i) % c.ring_cap];
  return QK_OK;
}

int qk_batch_read_containers(qk_batch* b, uint64_t rep, uint32_t comp, uint32_t* handles, uint64_t* resources,
                             uint64_t cap, uint64_t* n_out) {
  int rc = check_range(b, rep, 1);
  if (rc) return rc;
  if (comp >= b->low.hdr.n_comp || !(b->low.comps[comp].flags & DC_CONTAINER)) {
    qk_set_error("qk_batch_read_containers: component %u does not hold containers", comp);
    return QK_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(b->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  const DevComp& c = b->low.comps[comp];
  /* (item slot in plane32 or cold32, payload slot in cold64) of every container held, in order */
  std::vector<uint32_t> item_slot, pay_slot;
  bool item_cold = false;
  if (c.kind == QK_DISCRETE_STOCK) {
    uint32_t hc = 0;
    CUDA_TRY(fetch(b->g32, b->rep_pad, c.o32 + S32_HC, rep, &hc));
    const uint32_t head = QK_HC_HEAD(hc), cnt = QK_HC_CNT(hc);
    item_cold = true;
    for (uint32_t i = 0; i < cnt; ++i) {
      const uint32_t pos = (head + i) % c.ring_cap;
      item_slot.push_back(c.ocold32 + pos);
      pay_slot.push_back(c.ocold64 + 3u * pos);
    }
  } else if (c.kind == QK_DISCRETE_PARALLEL_PROCESS || c.kind == QK_CONTAINER_LOAD_PROCESS ||
             c.kind == QK_CONTAINER_UNLOAD_PROCESS) {
    uint32_t nprog = 0, cq = 0;
    CUDA_TRY(cudaMemcpy(&nprog, warm32_at(b, c.o32 + PP32_NPROG, rep), 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&cq, warm32_at(b, c.o32 + PP32_COMPLETE, rep), 4, cudaMemcpyDeviceToHost));
    const uint32_t ring = c.ring_cap, item0 = c.o32 + PP32_BASE_COUNT;
    for (uint32_t i = 0; i < nprog; ++i) {
      item_slot.push_back(item0 + i);
      pay_slot.push_back(c.ocold64 + 3u * i);
    }
    for (uint32_t i = 0; i < (cq >> 16); ++i) {
      const uint32_t pos = ((cq & 0xFFFFu) + i) % ring;
      item_slot.push_back(item0 + ring + pos);
      pay_slot.push_back(c.ocold64 + 3u * (ring + pos));
    }
  } else {
    uint32_t flags = 0;
    CUDA_TRY(cudaMemcpy(&flags, warm32_at(b, c.o32 + P32_FLAGS, rep), 4, cudaMemcpyDeviceToHost));
    if (flags & PF_HAS_ITEM) {
      item_slot.push_back(c.o32 + P32_ITEM);
      pay_slot.push_back(c.ocold64);
    }
  }
  if (n_out) *n_out = item_slot.size();
  if (!handles || !resources) return QK_OK;
  if (cap < item_slot.size()) return QK_ERR_BUFFER_TOO_SMALL;
  for (size_t i = 0; i < item_slot.size(); ++i) {
    CUDA_TRY(cudaMemcpy(&handles[i], item_cold ? cold32_at(b, item_slot[i], rep) : warm32_at(b, item_slot[i], rep), 4,
                        cudaMemcpyDeviceToHost));
    for (uint32_t k = 0; k < 3; ++k)
      CUDA_TRY(cudaMemcpy(&resources[i * 3 + k], cold64_at(b, pay_slot[i] + k, rep), 8, cudaMemcpyDeviceToHost));
  }
  return QK_OK;
}

} /* extern "C" */
