/*
 * qk_host_emul.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Shims that let g++ compile quokkasim_b200/csrc/qk_kernels.cuh on the host, one "thread" per block
 * (nt == 1), so the CPU-only test suite (-m "not gpu") can check the kernel LOGIC against the oracle
 * without a GPU.  It is never part of libquokka_b200.so: the product build does not define QK_HOST_EMUL
 * and does not have tests/emul on its include path.
 */
#ifndef QK_HOST_EMUL_H
#define QK_HOST_EMUL_H
#include <sched.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__

struct uint4 {
  uint32_t x, y, z, w;
};
struct uint2 {
  uint32_t x, y;
};

extern thread_local uint8_t* qk_emul_smem;
extern thread_local uint32_t qk_emul_bid;
/* The thread-per-replication kernel runs as one "thread" per block: every collective is trivial.  The lane-group kernel
 * (qk_group.cuh) runs as a TEAM of host threads, one per lane of one small block that is also its only warp; the warp
 * collectives go through an exchange array between two barriers, which also gives the memory ordering __syncwarp
 * gives on the device. */
struct QkEmulTeam {
  uint32_t n;
  std::atomic<uint32_t> count;
  std::atomic<uint32_t> gen;
  uint64_t slot[64];
};
extern thread_local QkEmulTeam* qk_emul_team;
extern thread_local uint32_t qk_emul_tid, qk_emul_nt;
static inline void qk_emul_barrier() {
  QkEmulTeam* t = qk_emul_team;
  if (!t) return;
  const uint32_t g = t->gen.load(std::memory_order_acquire);
  if (t->count.fetch_add(1, std::memory_order_acq_rel) + 1 == t->n) {
    t->count.store(0, std::memory_order_relaxed);
    t->gen.store(g + 1, std::memory_order_release);
  } else {
    uint32_t spins = 0;
    while (t->gen.load(std::memory_order_acquire) == g)
      if ((++spins & 63u) == 0) sched_yield();
  }
}
static inline uint64_t qk_emul_shfl(uint64_t v, uint32_t src) {
  QkEmulTeam* t = qk_emul_team;
  if (!t) return v;
  t->slot[qk_emul_tid] = v;
  qk_emul_barrier();
  const uint64_t r = t->slot[src % t->n];
  qk_emul_barrier();
  return r;
}
static inline uint32_t qk_emul_ballot(bool p) {
  QkEmulTeam* t = qk_emul_team;
  if (!t) return p ? 1u : 0u;
  t->slot[qk_emul_tid] = p ? 1u : 0u;
  qk_emul_barrier();
  uint32_t m = 0;
  for (uint32_t i = 0; i < t->n; ++i) m |= (uint32_t)t->slot[i] << i;
  qk_emul_barrier();
  return m;
}
extern std::atomic<uint64_t> qk_emul_counters[4];
static inline void qk_emul_count(int k) { qk_emul_counters[k].fetch_add(1, std::memory_order_relaxed); }
#define qk_smem qk_emul_smem
#define QK_TID qk_emul_tid
#define QK_NT qk_emul_nt
#define QK_BID qk_emul_bid
#define QK_WARP qk_emul_nt
#define QK_SYNCWARP() qk_emul_barrier()
#define QK_SYNCBLOCK() qk_emul_barrier()
#define QK_ANY_SYNC(p) (qk_emul_ballot(p) != 0u)
#define QK_BALLOT(p) qk_emul_ballot(p)
#define QK_SHFL32(v, src) ((uint32_t)qk_emul_shfl((uint64_t)(v), (uint32_t)(src)))
#define QK_SHFL_XOR32(v, m) ((uint32_t)qk_emul_shfl((uint64_t)(v), qk_emul_tid ^ (uint32_t)(m)))
static inline uint32_t qk_emul_reduce_min(uint32_t mask, uint32_t v) {
  QkEmulTeam* t = qk_emul_team;
  if (!t) return v;
  t->slot[qk_emul_tid] = v;
  qk_emul_barrier();
  uint32_t m = 0xFFFFFFFFu;
  for (uint32_t i = 0; i < t->n; ++i)
    if ((mask >> i) & 1u) m = (uint32_t)t->slot[i] < m ? (uint32_t)t->slot[i] : m;
  qk_emul_barrier();
  return m;
}
#define QK_REDUCE_MIN32(mask, v) qk_emul_reduce_min((mask), (v))
#define QK_POPC(x) __builtin_popcount(x)
static inline void qk_red_add64(uint64_t* p, uint64_t v) { *p += v; }
static inline void qk_red_add32(uint32_t* p, uint32_t v) { *p += v; }
static inline void qk_red_max32(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
static inline void qk_red_addf64(uint64_t* p, double v) {
  double d;
  std::memcpy(&d, p, 8);
  d += v;
  std::memcpy(p, &d, 8);
}
static inline void __syncthreads() { qk_emul_barrier(); }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) {
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline long long __double_as_longlong(double d) {
  long long v;
  std::memcpy(&v, &d, 8);
  return v;
}
static inline double __longlong_as_double(long long v) {
  double d;
  std::memcpy(&d, &v, 8);
  return d;
}
/* built with -ffp-contract=off: each of these is one IEEE round-to-nearest operation */
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline int __ffs(int v) { return __builtin_ffs(v); }


static inline void qk_stage8(uint64_t* dst, const uint64_t* src) { *dst = *src; }
static inline void qk_stage4(uint32_t* dst, const uint32_t* src) { *dst = *src; }
static inline void qk_stage_wait() {}
static inline void qk_async_wait() {}
static inline void qk_store_record(uint4* r, const uint4& a, const uint4& b) {
  r[0] = a;
  r[1] = b;
}
#endif
