/*
 * qk_device.cuh -- device-side numeric primitives for the batched-replication kernels (sm_100a).
 *
 *   - Duration::from_secs_f64 in 64-bit integer arithmetic (Rust core::time try_from_secs!, as used at
 *     discrete.rs:490,860,1107,1380 and delays.rs:102,132,163): round-to-nearest-ns, ties-to-even, from
 *     the binary mantissa.  Written independently of the oracle's u128 restatement.
 *   - Philox4x32-10 counter RNG and the native samplers for Distribution::sample (common.rs:152-178).
 *     Every floating-point step uses an explicit round-to-nearest intrinsic so that no FMA contraction can
 *     change the result (DESIGN.md "Native sampling arithmetic").
 */
#ifndef QK_DEVICE_CUH
#define QK_DEVICE_CUH
#include <stdint.h>

#include "qk_internal.h"

#ifdef QK_HOST_EMUL
/* tests/emul/qk_host_emul.h: lets the CPU-only test suite compile THIS source with g++ to check the kernel
 * logic against the oracle without a GPU.  Never defined in the product build. */
#include "qk_host_emul.h"
#else
#include <cuda_runtime.h>
/* file-scope dynamic shared memory: every device function sees a pointer whose address space is known */
extern __shared__ __align__(128) uint8_t qk_smem[]; /* 128: destination of TMA tensor copies */
#define QK_TID threadIdx.x
#define QK_NT blockDim.x
#define QK_BID blockIdx.x
#define QK_SYNCWARP() __syncwarp()
#define QK_ANY_SYNC(p) __any_sync(0xFFFFFFFFu, (p))
#define QK_BALLOT(p) __ballot_sync(0xFFFFFFFFu, (p))
#define QK_POPC(x) __popc(x)
__device__ __forceinline__ void qk_red_add64(uint64_t* p, uint64_t v) { atomicAdd((unsigned long long*)p, (unsigned long long)v); }
__device__ __forceinline__ void qk_red_add32(uint32_t* p, uint32_t v) { atomicAdd(p, v); }
__device__ __forceinline__ void qk_red_max32(uint32_t* p, uint32_t v) { atomicMax(p, v); }
/* fp64 accumulator stored in a u64 plane: one thread owns the address, so the sum order is the program order */
__device__ __forceinline__ void qk_red_addf64(uint64_t* p, double v) { atomicAdd(reinterpret_cast<double*>(p), v); }
#endif

#ifndef QK_HOST_EMUL
/* one 32-byte log record = ONE streaming store (evict-first: log lines are never read back by the kernel and must not
 * push the cold planes out of L2): sm_100 has 256-bit global stores (STG.E.EF.ENL2.256), which halves the L1 tag
 * work of the scattered per-replication log writes compared with two STG.128 */
__device__ __forceinline__ void qk_store_record(uint4* r, const uint4& a, const uint4& b) {
  asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(r), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
#endif

#ifndef QK_HOST_EMUL
/* asynchronous global -> shared copies for staging the state planes (cp.async = LDGSTS) */
__device__ __forceinline__ void qk_stage8(uint64_t* smem_dst, const uint64_t* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void qk_stage4(uint32_t* smem_dst, const uint32_t* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void qk_stage_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
/* every asynchronous copy this thread has issued has landed */
__device__ __forceinline__ void qk_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

#ifndef QK_HOST_EMUL
/* ---- bulk asynchronous copies (TMA engine, 1-D): one instruction moves a whole slot row ------------------------- */
__device__ __forceinline__ void qk_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void qk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void qk_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tQK_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra QK_DONE_%=;\n\t"
      "bra QK_WAIT_%=;\n\tQK_DONE_%=:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
/* global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on the mbarrier */
__device__ __forceinline__ void qk_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
/* 2-D tile of a state plane ([slot][replication], tensor map made by qk_batch_create) -> shared memory: ONE instruction
 * of the TMA engine moves the block's rows of EVERY slot (box = n_slots x block size, dense, no swizzle: exactly the
 * [slot][thread] layout the kernel works on); smem_dst 128-byte aligned; completes on the mbarrier */
__device__ __forceinline__ void qk_tma_load_2d(void* smem_dst, const void* tmap, uint32_t c0, uint32_t c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
/* ... and back */
__device__ __forceinline__ void qk_tma_store_2d(const void* tmap, uint32_t c0, uint32_t c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"((uint32_t)__cvta_generic_to_shared(smem_src))
               : "memory");
}
/* shared -> global */
__device__ __forceinline__ void qk_bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(smem_src)), "r"(bytes)
               : "memory");
}
/* before the block exits: the stores must have READ their shared-memory source (the block's shared memory is handed to
 * the next block the moment this one retires); that they have reached global memory is what the end of the grid
 * guarantees, and waiting for it here kept every block resident for one more trip to L2 -- at one event per launch the
 * kernel's rate is resident blocks / block lifetime (profiles/README.md R2.5) */
__device__ __forceinline__ void qk_bulk_store_wait() {
#ifdef QK_STORE_WAIT_FULL
  asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory");
#else
  asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
/* writes made through the generic proxy (ordinary STS) become visible to the asynchronous proxy (bulk stores) */
__device__ __forceinline__ void qk_fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

#define QK_DEV __device__ __forceinline__
#ifndef QK_REFILL_LANES
#define QK_REFILL_LANES 20 /* run the REFILL phase when this many lanes of the warp wait for a variate */
#endif

/* returns 0 and *ns, or a qk_rep_status fault code */
QK_DEV uint32_t qk_from_secs_f64(double secs, uint64_t* ns) {
  if (!(secs >= 0.0)) {
    /* negative -> TryFromFloatSecsErrorKind::Negative; NaN -> OverflowOrNan */
    return QK_REP_BAD_DURATION;
  }
  const uint64_t bits = (uint64_t)__double_as_longlong(secs);
  const int e = (int)((bits >> 52) & 0x7FFull) - 1023;
  const uint64_t mant = (bits & 0xFFFFFFFFFFFFFull) | 0x10000000000000ull;
  const uint64_t NPS = 1000000000ull;
  if (e < -31) {
    *ns = 0;
    return 0;
  }
  if (e >= 34) return QK_REP_BAD_DURATION; /* >= 2^34 s does not fit u64 nanoseconds (or inf/NaN) */
  uint64_t s, nanos;
  bool round_up;
  if (e >= 0) {
    s = mant >> (52 - e);
    const uint64_t frac = (mant << e) & 0xFFFFFFFFFFFFFull; /* 52-bit binary fraction */
    const uint64_t hi = __umul64hi(frac, NPS);
    const uint64_t lo = frac * NPS;
    nanos = (hi << 12) | (lo >> 52);
    const uint64_t rem = lo & 0xFFFFFFFFFFFFFull;
    const uint64_t half = 1ull << 51;
    round_up = rem > half || (rem == half && (nanos & 1ull));
  } else {
    s = 0;
    const int sh = 52 - e; /* 53..83 */
    const uint64_t hi = __umul64hi(mant, NPS);
    const uint64_t lo = mant * NPS;
    if (sh < 64) {
      nanos = (hi << (64 - sh)) | (lo >> sh);
      const uint64_t rem = lo & ((1ull << sh) - 1ull);
      const uint64_t half = 1ull << (sh - 1);
      round_up = rem > half || (rem == half && (nanos & 1ull));
    } else {
      const int k = sh - 64; /* 0..19 */
      nanos = hi >> k;
      const uint64_t rem_hi = hi & ((1ull << k) - 1ull);
      const uint64_t half_hi = k ? (1ull << (k - 1)) : 0ull;
      const uint64_t half_lo = k ? 0ull : (1ull << 63);
      const bool gt = rem_hi > half_hi || (rem_hi == half_hi && lo > half_lo);
      const bool eq = rem_hi == half_hi && lo == half_lo;
      round_up = gt || (eq && (nanos & 1ull));
    }
  }
  nanos += round_up ? 1ull : 0ull;
  if (nanos == NPS) {
    s += 1;
    nanos = 0;
  }
  if (s > 18446744073ull) return QK_REP_BAD_DURATION;
  *ns = s * NPS + nanos;
  return 0;
}

QK_DEV void qk_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                             uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0;
    const uint32_t n2 = h0 ^ c3 ^ k1;
    c0 = n0;
    c1 = l1;
    c2 = n2;
    c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

QK_DEV double qk_u01(uint32_t lo, uint32_t hi) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return __dmul_rn((double)(v >> 11), 1.0 / 9007199254740992.0);
}

/* natural log from + - * / only; x must be a positive normal double */
QK_DEV double qk_det_ln(double x) {
  uint64_t b = (uint64_t)__double_as_longlong(x);
  int e = (int)((b >> 52) & 0x7FFull) - 1022;
  double m = __longlong_as_double((long long)((b & 0x800FFFFFFFFFFFFFull) | 0x3FE0000000000000ull)); /* [0.5,1) */
  if (m < 0.70710678118654752440) {
    m = __dmul_rn(m, 2.0);
    e -= 1;
  }
  const double s = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0));
  const double z = __dmul_rn(s, s);
  double p = 1.0 / 27.0;
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 25.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 23.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 21.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 19.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 17.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 15.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 13.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 11.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 9.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 7.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 5.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0 / 3.0);
  p = __dadd_rn(__dmul_rn(p, z), 1.0);
  double r = __dmul_rn(2.0, s);
  r = __dmul_rn(r, p);
  const double el = __dmul_rn((double)e, 0.69314718055994530942);
  return __dadd_rn(r, el);
}

/* one native variate of a non-constant distribution; draws = the instance's draw counter (in and out).
 * Everything by value and NOT inlined: one copy of the fp64 code in the kernel, no stack traffic. */
struct QkVariate {
  double value;
  uint32_t draws;
};
static __device__ __noinline__ QkVariate qk_sample_native(uint32_t kind, double p0, double p1, double p2, double p3,
                                                   uint32_t stream, uint32_t key0, uint32_t key1, uint32_t rep_lo,
                                                   uint32_t rep_hi, uint32_t draws) {
  const double p[4] = {p0, p1, p2, p3};
  for (;;) {
    uint32_t x[4];
    qk_philox4x32_10(rep_lo, rep_hi, stream, draws, key0, key1, x);
    draws += 1;
    const double u = qk_u01(x[0], x[1]);
    if (kind == QK_DIST_UNIFORM) {
      const double w = __dsub_rn(p[1], p[0]);
      return QkVariate{__dadd_rn(p[0], __dmul_rn(u, w)), draws};
    } else if (kind == QK_DIST_TRIANGULAR) {
      const double mn = p[0], mx = p[1], mode = p[2];
      const double diff_mode_min = __dsub_rn(mode, mn);
      const double range = __dsub_rn(mx, mn);
      const double f_range = __dmul_rn(u, range);
      /* both arms of rand_distr's Triangular::sample, selected per lane: ONE square root for the whole warp instead
       * of two half-empty passes through the fp64 sqrt sequence (same operations on the selected arm, same bits) */
      const bool lo = f_range < diff_mode_min;
      const double a = __dsub_rn(range, f_range);
      const double b = __dsub_rn(mx, mode);
      const double arg = lo ? __dmul_rn(f_range, diff_mode_min) : __dmul_rn(a, b);
      const double r = __dsqrt_rn(arg);
      return QkVariate{lo ? __dadd_rn(mn, r) : __dsub_rn(mx, r), draws};
    } else if (kind == QK_DIST_EXPONENTIAL) {
      const double om = __dsub_rn(1.0, u);
      const double l = qk_det_ln(om);
      return QkVariate{__dmul_rn(__dsub_rn(0.0, l), p[0]), draws};
    } else { /* NORMAL / TRUNCNORMAL: Marsaglia polar on the two uniforms of one Philox block */
      const double u2 = qk_u01(x[2], x[3]);
      const double v1 = __dsub_rn(__dmul_rn(2.0, u), 1.0);
      const double v2 = __dsub_rn(__dmul_rn(2.0, u2), 1.0);
      const double s = __dadd_rn(__dmul_rn(v1, v1), __dmul_rn(v2, v2));
      if (s >= 1.0 || s == 0.0) continue;
      const double l = qk_det_ln(s);
      const double q = __ddiv_rn(__dmul_rn(-2.0, l), s);
      const double f = __dsqrt_rn(q);
      const double xv = __dadd_rn(p[0], __dmul_rn(p[1], __dmul_rn(v1, f)));
      if (kind == QK_DIST_TRUNCNORMAL && !(xv >= p[2] && xv <= p[3])) continue; /* common.rs:166-173 */
      return QkVariate{xv, draws};
    }
  }
}

#endif
