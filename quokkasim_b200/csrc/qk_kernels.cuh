/*
 * qk_kernels.cuh -- the per-event state-update kernel of the batched-replication executor (sm_100a).
 *
 * One thread owns one replication ("thread group" of size 1): its scheduler queue, component state
 * machines and stock rings live in shared memory, replication-fastest, for the whole launch, so HBM is
 * touched only to load/store the state planes once per launch and to stream log records.
 *
 * What it replaces in the reference (paths relative to /root/reference/quokkasim/src):
 *   - NeXosim's scheduler queue + executor behind Simulation::step_until / step
 *     (call sites quokkasim_examples/src/bin/discrete_queue.rs:111; cx.schedule_event discrete.rs:194,219;
 *     cx.schedule_keyed_event discrete.rs:549,556; action_key.cancel() discrete.rs:548): here every
 *     component has ONE fire-time slot, the queue is "min over fire[]" (+ a stale bit mask for the
 *     dropped-but-not-cancelled handles of SURVEY A.6 Q3), ordered by (time, registration rank).
 *   - Process::update_state = pre -> impl -> post for DiscreteProcess / DiscreteSource / DiscreteSink /
 *     DiscreteParallelProcess (core.rs:370-376; discrete.rs:404-564, 779-920, 1025-1167, 1280-1458)
 *   - DiscreteStock add / remove / emit_change (core.rs:177-204; discrete.rs:161-233)
 *   - DelayModes::update_state (delays.rs:86-142), BasicEnvironment::set_state (core.rs:256-265)
 *   - Distribution::sample (common.rs:152-178) via Philox4x32-10 or a pre-drawn variate tape
 *   - every `self.log(..)` (discrete.rs:235-251, 566-582) as a 32-byte binary record
 *
 * Control flow is arranged so that all lanes of a warp run the SAME code most of the time: an action
 * (keyed event, stock emit_change, scheduled environment change) is expanded into a short list of
 * update_state(p) calls, and the warp iterates "one update_state per lane per trip"; which component p is,
 * is data, not control.
 */
#ifndef QK_KERNELS_CUH
#define QK_KERNELS_CUH
#include "qk_device.cuh"

/* fields a vector resource may have in the kernels of feature set FT (qk_vector.cuh) */
template <int FT> struct QkVd { static constexpr int v = FT == 3 ? QK_VEC_MAX_DIM : 3; };
#define qk_vd(ft) (QkVd<(ft)>::v)
#define QK_D2U(d) ((uint64_t)__double_as_longlong(d))
#define QK_U2D(u) (__longlong_as_double((long long)(u)))

/* a CUtensorMap (128 opaque bytes, 64-byte aligned) without <cuda.h>: the host emulation compiles this file too */
struct alignas(64) QkTmap {
  uint64_t opaque[16];
};

struct KParams {
  uint64_t* g64; /* [n64][rep_pad] */
  uint32_t* g32; /* [n32][rep_pad] */
  uint64_t* gs64; /* statistics planes [n_comp * 2][rep_pad], HBM only */
  uint32_t* gs32;
  uint64_t* gc64; /* cold state planes [n64c][rep_pad] / [n32c][rep_pad], global memory only */
  uint32_t* gc32;
  uint32_t n64c, n32c;
  const uint8_t* tables;
  uint32_t table_bytes; /* multiple of 16 */
  uint32_t n64, n32;
  uint64_t n_reps, rep_pad, rep_offset, seed;
  uint4* log;        /* [rep_pad][log_cap] records (2 x uint4 each) or NULL */
  uint32_t log_cap;  /* power of two */
  const double* const* tapes; /* [n_dist] device pointers or NULL entries */
  const uint64_t* tape_len;   /* [n_dist] per-replication length */
  uint64_t t0, t_until, event_budget;
  uint32_t do_init;
  /* sliced stepping (qk_batch.cu, step_sliced): a launch may cover only the tiles [tile0, tile0 + gridDim.x) of the
   * batch (a tile = one block's worth of replications) */
  uint32_t tile0;
  /* lane-group kernel (qk_group.cuh): replications per block, row strides of the shared-memory state (elements),
   * byte offset of the mbarrier */
  uint32_t grp_reps, grp_stride64, grp_stride32, grp_bar_off;
  uint32_t grp_x64; /* extra 64-bit rows after the state rows (the phased kernel's loop state); the 32-bit rows follow */
  /* staging by TMA tensor copies (qk_step_kernel, on-chip and split placements): tensor maps of the two hot planes as
   * 2-D arrays [slot][replication] with a box of (all slots) x (one block's replications) */
  uint32_t use_tma;
  QkTmap tm64, tm32;
};

struct Ctx {
  uint32_t base64, base32; /* on-chip: element offsets of this lane's slot 0 inside qk_smem */
  uint64_t* g64;           /* HBM-resident state: plane pointers offset to this replication */
  uint32_t* g32;
  uint32_t stride;         /* on-chip: threads per block ; HBM: padded replication count */
  uint32_t stride32;       /* run-time-stride on-chip mode only (SM < 0): row stride of the 32-bit slots */
  uint64_t* gs64;          /* statistics planes, offset to this replication; row stride = rep_pad */
  uint32_t* gs32;
  uint64_t* gc64;          /* cold planes, offset to this replication */
  uint32_t* gc32;
  uint32_t rep_pad; /* < 2^32 replications per batch */
  const DevHeader* hdr;
  const DevComp* comps;
  const DevDist* dists;
  const DevDelayMode* dms;
  const uint16_t* subs;
  const uint16_t* sched;
  const DevExt* ext;
  const DevVec* vecs;
  const DevHot* hot;
  const DevSubQ* subq;   /* quick-path descriptor of every subs[] entry */
  const double* ccap;    /* container capacity by handle (DevHeader.off_ccap) */
  uint32_t pf32, n_pf_words; /* hoisted DevHeader fields (re-reading them after every state store is wasted LDS) */
  uint64_t now;
  uint32_t status;
  uint4* log_base;
  uint32_t log_mask;
  uint32_t log_cnt;
  uint32_t key0, key1, rep_lo, rep_hi;
  uint32_t npend; /* consumed pre-drawn durations waiting for the REFILL phase */
  /* two-tier scheduler queue (QK_SM_QSPLIT): first cold64 slot of the process-like fire array, first hot32 slot of the busy
   * mask, and the cached minimum over that array by (time, rank) -- exact while kvalid */
  uint32_t cfire0, busy32, emask32, kidx;
  uint32_t gshift, n_groups, gmin64, gidx32, gdirty32; /* group minima of the cold fire array (DevHeader) */
  uint64_t kbest;
  bool kvalid;
  uint64_t rep_local;
  const double* const* tapes;
  const uint64_t* tape_len;
};

/* state accessors: the address space is a compile-time fact (shared via the file-scope qk_smem symbol, or
 * global), so every access is an LDS/STS or LDG/STG -- never a generic load */
template <int SM>
__device__ __forceinline__ uint64_t* qk_p64(const Ctx& cx) {
  if constexpr (SM == 0) return cx.g64; else return reinterpret_cast<uint64_t*>(qk_smem) + cx.base64;
}
template <int SM>
__device__ __forceinline__ uint32_t* qk_p32(const Ctx& cx) {
  if constexpr (SM == 0) return cx.g32; else return reinterpret_cast<uint32_t*>(qk_smem) + cx.base32;
}
/* SM is the state-placement mode of a kernel instantiation:
 *   SM  > 0 : on chip, block size == SM known at compile time (slot offsets fold into LDS/STS immediates)
 *   SM  < 0 : on chip, block size read at run time (host emulation / unusual sizes)
 *   SM == 0 : state operated on in place in the HBM planes (models too large for shared memory) */
template <int SM>
__device__ __forceinline__ auto qk_ix(uint32_t slot, uint32_t stride) {
  if constexpr (SM == 0)
    return (size_t)slot * stride;
  else if constexpr (SM > 0)
    return slot * (uint32_t)SM;
  else
    return slot * stride;
}
#define S64(slot) (qk_p64<SM>(cx)[qk_ix<SM>((uint32_t)(slot), cx.stride)])
#define S32(slot) (qk_p32<SM>(cx)[qk_ix<SM>((uint32_t)(slot), SM < 0 ? cx.stride32 : cx.stride)])
/* component records ("warm" state: countdowns, flags, items, pre-drawn durations, draw counters, resources): hot-plane
 * slots, except under the SPLIT placement (SM == QK_SM_SPLIT, DevHeader.split), where the lowering allocated them in the
 * cold planes so that only the scheduler queue and the stocks' state words occupy shared memory */
#define QK_SM_SPLIT (-2)
/* ... and QSPLIT: the split placement with the two-tier scheduler queue (qk_lower, placement 2): the hot fire array holds
 * the stocks' slots only, the keyed events of the process-like components live in a cold array by rank */
#define QK_SM_QSPLIT (-3)
#define qk_is_split(sm) ((sm) == QK_SM_SPLIT || (sm) == QK_SM_QSPLIT)
#define qk_is_two_tier(sm) ((sm) == QK_SM_QSPLIT)
/* keyed-event slot of rank idx under the two-tier queue (cold plane) */
#define QK_CFIRE(idx) C64(cx.cfire0 + (idx))
template <int SM>
__device__ __forceinline__ uint64_t& qk_w64(const Ctx& cx, uint32_t slot) {
  if constexpr (qk_is_split(SM)) return cx.gc64[slot];
  else return qk_p64<SM>(cx)[qk_ix<SM>(slot, cx.stride)];
}
template <int SM>
__device__ __forceinline__ uint32_t& qk_w32(const Ctx& cx, uint32_t slot) {
  if constexpr (qk_is_split(SM)) return cx.gc32[slot];
  else return qk_p32<SM>(cx)[qk_ix<SM>(slot, SM < 0 ? cx.stride32 : cx.stride)];
}
#define W64(slot) (qk_w64<SM>(cx, (uint32_t)(slot)))
#define W32(slot) (qk_w32<SM>(cx, (uint32_t)(slot)))
/* cold state: always global memory, whatever the placement of the hot planes */
/* layout: [slot][replication] like the hot planes -- except under the split placements, where the cold planes hold the
 * component records and are laid out [replication][slot] (DevHeader.split): the handlers of a warp work on 32 different
 * components, so a 32-byte sector of a [slot][replication] plane would carry one useful word, while a replication's own
 * record -- countdown, previous-check time, pre-drawn duration / flags, item, draw counter -- is two sectors here, and
 * its keyed-event array (two-tier queue) one contiguous run */
#define C64(slot) (cx.gc64[qk_is_split(SM) ? (uint64_t)(uint32_t)(slot) : (uint64_t)(uint32_t)(slot) * cx.rep_pad])
#define C32(slot) (cx.gc32[qk_is_split(SM) ? (uint64_t)(uint32_t)(slot) : (uint64_t)(uint32_t)(slot) * cx.rep_pad])
/* statistics accumulators: fire-and-forget RED atomics on the HBM planes (see qk_internal.h) */
#ifdef QK_EXPERIMENT_NO_STATS /* perf experiment only: results are wrong */
#define QK_ST64_ADD(comp, k, v) ((void)0)
#define QK_ST32_ADD(comp, k, v) ((void)0)
#define QK_ST32_MAX(comp, k, v) ((void)0)
#else
#define QK_STROW(comp, k) ((uint64_t)(uint32_t)((comp)*ST_PER_COMP + (k)) * cx.rep_pad)
#define QK_ST64_ADD(comp, k, v) qk_red_add64(&cx.gs64[QK_STROW(comp, k)], (v))
#define QK_ST32_ADD(comp, k, v) qk_red_add32(&cx.gs32[QK_STROW(comp, k)], (v))
#define QK_ST32_MAX(comp, k, v) qk_red_max32(&cx.gs32[QK_STROW(comp, k)], (v))
#endif
#define SRC_INIT (((uint64_t)QK_SRC_INIT) << 32)
#define SRC_SCH (((uint64_t)QK_SRC_SCH) << 32)

/* one LDS.128 -> a 16-byte table quad in registers */
template <class T>
__device__ __forceinline__ T qk_ldq(const T* p) {
  static_assert(sizeof(T) == 16, "table quads are 16 bytes");
  union {
    uint4 v;
    T t;
  } u;
  u.v = *reinterpret_cast<const uint4*>(p);
  return u.t;
}
#define QK_HOT_A(comp) qk_ldq(&cx.hot[comp].a)
#define QK_HOT_B(comp) qk_ldq(&cx.hot[comp].b)
#define QK_HOT_SB(comp) qk_ldq(&cx.hot[comp].sb)
#define QK_HOT_C(comp) qk_ldq(&cx.hot[comp].c)

/* ---- log(): allocate EventId "{code}_{seq:06}", emit one 32-byte record, return the new id ------------ */
template <bool LOG, int SM>
__device__ __forceinline__ uint64_t qk_log(Ctx& cx, const DevHotA& c, uint32_t comp, uint64_t src, uint32_t kind,
                                           uint32_t reason, uint64_t payload) {
  if (!LOG) return 0;
  const uint32_t seq = S32(c.olog32 + L32_SEQ);
  S32(c.olog32 + L32_SEQ) = seq + 1;
  if (c.flags & DC_LOGGED) {
    uint4* r = cx.log_base + 2 * (size_t)(cx.log_cnt & cx.log_mask);
    uint4 a, b;
    a.x = (uint32_t)cx.now;
    a.y = (uint32_t)(cx.now >> 32);
    a.z = comp | (kind << 16) | (reason << 24);
    a.w = seq;
    b.x = (uint32_t)(src >> 32) & 0xFFFFu; /* src_comp | aux(0) << 16 */
    b.y = (uint32_t)src;
    b.z = (uint32_t)payload;
    b.w = (uint32_t)(payload >> 32);
    qk_store_record(r, a, b);
    cx.log_cnt += 1;
  }
  return ((uint64_t)comp << 32) | seq;
}

/* continuation record of a record that serialises a Vector3Container (ProcessStart / ProcessFinish / Add / Remove):
 * the container's resource as three raw fp64 words (w[0] == QK_CONTAINER_NONE_BITS for None).  Same layout as the
 * vector components' continuation records (quokka_b200.h, QK_LOG_VEC_DATA). */
template <bool LOG, int SM>
__device__ __forceinline__ void qk_log_pay(Ctx& cx, uint32_t flags, uint32_t comp, uint64_t id, const uint64_t* w) {
  if (!LOG) return;
  if (flags & DC_LOGGED) {
    uint4* r = cx.log_base + 2 * (size_t)(cx.log_cnt & cx.log_mask);
    uint4 a, b;
    a.x = (uint32_t)w[0];
    a.y = (uint32_t)(w[0] >> 32);
    a.z = comp | ((uint32_t)QK_LOG_VEC_DATA << 16);
    a.w = (uint32_t)id;
    b.x = (uint32_t)w[1];
    b.y = (uint32_t)(w[1] >> 32);
    b.z = (uint32_t)w[2];
    b.w = (uint32_t)(w[2] >> 32);
    qk_store_record(r, a, b);
    cx.log_cnt += 1;
  }
}

/* ---- Distribution::sample (common.rs:152-178) + Duration::from_secs_f64 --------------------------------
 * ONE out-of-line copy of "next variate of this instance (tape or Philox), optionally converted to a Duration": it is
 * needed at some fifteen places of the general kernels, and inlined there it was a fifth of their code (the general
 * kernels are 140-240 KB of SASS against a 32 KB L1.5 instruction cache, and the container model spent 61 % of its stall
 * samples waiting for instructions).  Everything travels by value, so no caller state is forced into memory. */
struct QkDrawn {
  double value;
  uint64_t ns;
  uint32_t draws, fault;
};
static __device__ __noinline__ QkDrawn qk_draw(uint32_t kind, double p0, double p1, double p2, double p3, uint32_t stream,
                                               uint32_t key0, uint32_t key1, uint32_t rep_lo, uint32_t rep_hi,
                                               uint32_t draws, const double* tape, uint64_t tape_len, uint64_t rep_local,
                                               uint32_t want_duration) {
  QkDrawn r;
  r.ns = 0;
  r.fault = 0;
  if (kind == QK_DIST_CONSTANT) { /* a tape on a Constant is never consulted (common.rs:155) */
    r.value = p0;
    r.draws = draws;
  } else if (tape) {
    if (draws >= tape_len) {
      r.value = 1.0;
      r.draws = draws;
      r.fault = QK_REP_TAPE_EXHAUSTED;
      return r;
    }
    r.value = tape[rep_local * tape_len + draws];
    r.draws = draws + 1;
  } else {
    const QkVariate v = qk_sample_native(kind, p0, p1, p2, p3, stream, key0, key1, rep_lo, rep_hi, draws);
    r.value = v.value;
    r.draws = v.draws;
  }
  if (want_duration) r.fault = qk_from_secs_f64(r.value, &r.ns);
  return r;
}

template <int SM>
__device__ __forceinline__ QkDrawn qk_draw_for(Ctx& cx, uint32_t dist_id, uint32_t want_duration) {
  const DevDist* d = &cx.dists[dist_id];
  const bool constant = d->kind == QK_DIST_CONSTANT; /* has no draw counter */
  const double* tape = cx.tapes ? cx.tapes[dist_id] : nullptr;
  const QkDrawn r = qk_draw(d->kind, d->p[0], d->p[1], d->p[2], d->p[3], d->stream, cx.key0, cx.key1, cx.rep_lo, cx.rep_hi,
                            constant ? 0u : W32(d->draw_slot), tape, tape ? cx.tape_len[dist_id] : 0ull, cx.rep_local,
                            want_duration);
  if (!constant) W32(d->draw_slot) = r.draws;
  return r;
}

template <int SM>
__device__ __forceinline__ double qk_sample(Ctx& cx, uint32_t dist_id) {
  const DevDist* d = &cx.dists[dist_id];
  if (d->kind == QK_DIST_CONSTANT) return d->p[0];
  const QkDrawn r = qk_draw_for<SM>(cx, dist_id, 0u);
  if (r.fault) cx.status = r.fault;
  return r.value;
}

/* sample + convert; on a fault sets cx.status and returns 1 (callers abandon the handler) */
template <int SM>
__device__ __forceinline__ uint64_t qk_sample_duration(Ctx& cx, uint32_t dist_id) {
  const QkDrawn r = qk_draw_for<SM>(cx, dist_id, 1u);
  if (r.fault) {
    cx.status = r.fault;
    return 1;
  }
  return r.ns;
}

/* sample + convert for the prefetch slots: never raises, faults are encoded and raised when the value is consumed */
template <int SM>
__device__ __forceinline__ uint64_t qk_sample_duration_code(Ctx& cx, uint32_t dist_id) {
  const uint32_t saved = cx.status;
  cx.status = 0;
  const uint64_t ns = qk_sample_duration<SM>(cx, dist_id);
  const uint32_t f = cx.status;
  cx.status = saved;
  return f ? QK_DUR_FAULT_BASE + f : ns;
}

/* does the next update_state(comp) possibly start a process whose pre-drawn duration is still missing? */
template <int SM>
__device__ __forceinline__ bool qk_needs_refill_now(Ctx& cx, uint32_t comp) {
  /* stocks never appear in a call list, but a BasicEnvironment does (the init list), and its row has no DevHotB: only
   * the single-server discrete kinds pre-draw their process time */
  const DevHotA a = QK_HOT_A(comp);
  if (a.kind < QK_DISCRETE_PROCESS || a.kind > QK_DISCRETE_SINK) return false;
  const DevHotB b = QK_HOT_B(comp);
  if (b.o64_nextdur == 0xFFFF) return false;
  if constexpr (qk_is_split(SM)) {
    /* the component record is in global memory: ask the pending mask (a hot word) instead -- its bit is set exactly while
     * the slot holds QK_DUR_EMPTY -- and do not look at the countdown at all.  Refilling a little early only changes WHEN
     * a variate is drawn, never which one (every distribution instance has its own counter / tape position) */
    return ((S32(cx.pf32 + (b.pf_idx >> 5)) >> (b.pf_idx & 31)) & 1u) != 0;
  }
  if (W64(b.o64_nextdur) != QK_DUR_EMPTY) return false;
  const uint32_t flags = W32(a.o32 + P32_FLAGS);
  if (!(flags & PF_HAS_ITEM)) return true;
  return W64(a.o64 + P64_LEFT) <= cx.now - W64(a.o64 + P64_PREV); /* finishes in this call, may restart */
}

/* REFILL: every lane with consumed slots draws their next durations (Distribution::sample, common.rs:152-178) */
template <int SM>
__device__ __forceinline__ void qk_refill(Ctx& cx) {
  const uint32_t nw = cx.n_pf_words;
  const uint16_t* pf_comp = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(cx.hdr) + cx.hdr->off_pf);
  for (uint32_t w = 0; w < nw; ++w) {
    uint32_t m = S32(cx.pf32 + w);
    if (m) S32(cx.pf32 + w) = 0;
    while (m) {
      const uint32_t i = (uint32_t)__ffs((int)m) - 1;
      m &= m - 1;
      const DevHotB b = QK_HOT_B(pf_comp[w * 32 + i]);
      W64(b.o64_nextdur) = qk_sample_duration_code<SM>(cx, b.dist_time);
    }
  }
  cx.npend = 0;
}

/* ---- the fire slot of a process-like component (its pending keyed event, NeXosim's queue entry behind
 * cx.schedule_keyed_event, discrete.rs:549,556), by rank.  One hot slot per component, or -- two-tier queue -- a cold slot
 * plus a busy bit, with the cached minimum kept exact: every write goes through these three functions. */
template <int SM>
__device__ __forceinline__ bool qk_fire_busy(Ctx& cx, uint32_t idx) {
  return ((S32(cx.busy32 + (idx >> 5)) >> (idx & 31)) & 1u) != 0;
}
template <int SM>
__device__ __forceinline__ uint64_t qk_fire_get(Ctx& cx, uint32_t idx) {
  if constexpr (qk_is_two_tier(SM)) return qk_fire_busy<SM>(cx, idx) ? QK_CFIRE(idx) : QK_TIME_NONE;
  else return S64(G64_FIRE0 + idx);
}
template <int SM>
__device__ __forceinline__ void qk_fire_set(Ctx& cx, uint32_t idx, uint64_t t) { /* t != NONE */
  if constexpr (qk_is_two_tier(SM)) {
    QK_CFIRE(idx) = t;
    S32(cx.busy32 + (idx >> 5)) |= 1u << (idx & 31);
    const uint32_t g = idx >> cx.gshift;
    if (!((S32(cx.gdirty32) >> g) & 1u)) { /* (a dirty group is re-scanned before it is used: it will see this write) */
      const uint64_t gt = S64(cx.gmin64 + g);
      if (t < gt || (t == gt && idx < S32(cx.gidx32 + g))) {
        S64(cx.gmin64 + g) = t;
        S32(cx.gidx32 + g) = idx;
      }
    }
    if (cx.kvalid && (t < cx.kbest || (t == cx.kbest && idx < cx.kidx))) {
      cx.kbest = t;
      cx.kidx = idx;
    }
  } else {
    S64(G64_FIRE0 + idx) = t;
  }
}
template <int SM>
__device__ __forceinline__ void qk_fire_clear(Ctx& cx, uint32_t idx) {
  if constexpr (qk_is_two_tier(SM)) {
    QK_CFIRE(idx) = QK_TIME_NONE; /* (the scan reads the array, not the mask) */
    S32(cx.busy32 + (idx >> 5)) &= ~(1u << (idx & 31));
    const uint32_t g = idx >> cx.gshift;
    if (S64(cx.gmin64 + g) != QK_TIME_NONE && S32(cx.gidx32 + g) == idx) S32(cx.gdirty32) |= 1u << g; /* its group's minimum */
    if (cx.kvalid && idx == cx.kidx) cx.kvalid = false; /* the minimum is gone: the next pop re-scans the dirty groups */
  } else {
    S64(G64_FIRE0 + idx) = QK_TIME_NONE;
  }
}

/* ---- DiscreteStock (discrete.rs:158-252) ------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t qk_stock_cat(uint32_t cnt, uint32_t low, uint32_t max) {
  /* get_state discrete.rs:161-171: Empty test first */
  return cnt <= low ? 0u : (cnt >= max ? 2u : 1u);
}

/* cx.schedule_event(now + 1ns, Self::emit_change, ..) discrete.rs:193-194, 218-219.
 * HOT (discrete stocks): the source id of the OLDEST pending emit lives in the hot slot S32_EMITSRC, the cold ring only
 * holds the ones queued behind it -- almost every emit_change is alone in its queue, so the pop (whose id goes
 * straight into a log record) no longer waits for a global-memory load */
/* Faults travel UP as return values (true = this replication has faulted, cx.status says why): the lean kernel keeps
 * cx.status in local memory, and re-reading it after every call cost a local-memory round trip each time. */
template <bool LOG, int SM, bool HOT = false>
__device__ __forceinline__ bool qk_emit_push(Ctx& cx, const DevHotA& c, uint32_t emit_depth, uint32_t oemit,
                                             uint32_t src_seq) {
  const uint32_t fire = G64_FIRE0 + c.fire_idx;
  uint32_t e = S32(c.o32 + S32_EMIT);
  uint32_t n = e & 0xFFu, split = (e >> 8) & 0xFFu, head = (e >> 16) & 0xFFu;
  const uint64_t T = cx.now + 1;
  if (n == 0) {
    S64(fire) = T;
    if constexpr (qk_is_two_tier(SM)) S32(cx.emask32 + (c.fire_idx >> 5)) |= 1u << (c.fire_idx & 31); /* pending */
    split = 1;
    head = 0;
  } else if (S64(fire) == T) {
    split += 1;
  }
  n += 1;
  if (n > emit_depth) {
    cx.status = QK_REP_EMIT_OVERFLOW;
    return true;
  }
  if (LOG) {
    if (HOT && n == 1) {
      S32(c.o32 + S32_EMITSRC) = src_seq;
    } else {
      uint32_t pos = head + n - 1;
      if (pos >= emit_depth) pos -= emit_depth;
      C32(oemit + pos) = src_seq;
    }
  }
  S32(c.o32 + S32_EMIT) = n | (split << 8) | (head << 16);
  return false;
}

template <bool LOG, int SM, bool HOT = false>
__device__ __forceinline__ uint32_t qk_emit_pop(Ctx& cx, const DevHotA& c, uint32_t emit_depth, uint32_t oemit) {
  const uint32_t fire = G64_FIRE0 + c.fire_idx;
  uint32_t e = S32(c.o32 + S32_EMIT);
  uint32_t n = e & 0xFFu, split = (e >> 8) & 0xFFu, head = (e >> 16) & 0xFFu;
  uint32_t src_seq = 0;
  if (LOG) src_seq = HOT ? S32(c.o32 + S32_EMITSRC) : C32(oemit + head);
  head += 1;
  if (head >= emit_depth) head = 0;
  n -= 1;
  split -= 1;
  if (LOG && HOT && n != 0) S32(c.o32 + S32_EMITSRC) = C32(oemit + head); /* the next one moves up (rare) */
  if (n == 0) {
    S64(fire) = QK_TIME_NONE;
    if constexpr (qk_is_two_tier(SM)) S32(cx.emask32 + (c.fire_idx >> 5)) &= ~(1u << (c.fire_idx & 31));
  } else if (split == 0) {
    S64(fire) = S64(fire) + 1;
    split = n;
  }
  S32(c.o32 + S32_EMIT) = n | (split << 8) | (head << 16);
  return src_seq;
}

/* Stock::add = pre_add, add_impl, post_add (core.rs:177-183; discrete.rs:181-198). No capacity check. */
/* pay != nullptr: the item is a Vector3Container and pay[3] is what it carries (the stock is a container stock) */
template <bool LOG, int SM>
__device__ __forceinline__ bool qk_stock_add(Ctx& cx, uint32_t sidx, uint32_t item, uint64_t src,
                                             const uint64_t* pay = nullptr) {
  const DevHotA c = QK_HOT_A(sidx);
  const DevHotSB sb = QK_HOT_SB(sidx);
  const uint32_t hc = S32(c.o32 + S32_HC);
  uint32_t head = QK_HC_HEAD(hc), cnt = QK_HC_CNT(hc);
  const uint32_t prev = QK_HC_CAT(hc);
  if (cnt >= sb.ring_cap) {
    cx.status = QK_REP_STOCK_OVERFLOW;
    return true;
  }
  uint32_t pos = head + cnt;
  if (pos >= sb.ring_cap) pos -= sb.ring_cap;
  C32(c.olog64 + pos) = item; /* item ring: cold */
  if (LOG && SM != 0 && cnt == 0) S32(c.o32 + S32_HEADITEM) = item; /* the ring was empty: this is its head */
  if (pay) {
    const uint32_t op = cx.comps[sidx].ocold64 + 3u * pos;
    C64(op) = pay[0];
    C64(op + 1) = pay[1];
    C64(op + 2) = pay[2];
  }
  cnt += 1;
  const uint32_t cur = qk_stock_cat(cnt, sb.low, sb.max);
  S32(c.o32 + S32_HC) = QK_HC_PACK(head, cnt, cur);
  QK_ST64_ADD(sidx, ST64_TIME, 0ull - cx.now); /* occupancy integral = sum(t_remove) - sum(t_add) + cnt * T */
  QK_ST32_ADD(sidx, ST32_COUNT, 1u);
  QK_ST32_MAX(sidx, ST32_MAXOCC, cnt);
  const uint64_t id = qk_log<LOG, SM>(cx, c, sidx, src, QK_LOG_STOCK_ADD, 0, item);
  if (pay) qk_log_pay<LOG, SM>(cx, c.flags, sidx, id, pay);
  if (prev != cur)
    return qk_emit_push<LOG, SM, true>(cx, c, sb.emit_depth, c.olog64 + sb.ring_cap, (uint32_t)id);
  return false;
}

/* Stock::remove (core.rs:197-204; discrete.rs:200-225); returns QK_RM_GOT unless Remove(None), | QK_RM_FAULT */
#define QK_RM_GOT 1u
#define QK_RM_FAULT 2u
template <bool LOG, int SM>
__device__ __forceinline__ uint32_t qk_stock_remove(Ctx& cx, uint32_t sidx, uint64_t src, uint32_t* item_out,
                                                    uint64_t* pay_out = nullptr, const uint32_t* head_item = nullptr) {
  const DevHotA c = QK_HOT_A(sidx);
  const DevHotSB sb = QK_HOT_SB(sidx);
  const uint32_t hc = S32(c.o32 + S32_HC);
  uint32_t head = QK_HC_HEAD(hc), cnt = QK_HC_CNT(hc);
  const uint32_t prev = QK_HC_CAT(hc);
  uint32_t cur = prev;
  const bool some = cnt != 0;
  uint32_t item = 0;
  if (some) {
    /* With logging the head item goes straight into the Remove record, so a global-memory load here is a stall of
     * a few hundred cycles in every start.  On-chip kernels therefore keep a copy of the HEAD item in a hot slot:
     * reading it is a shared-memory load, and the copy of the NEXT head is fetched by an asynchronous
     * global->shared copy (LDGSTS) that nobody waits for until this stock's next remove. */
    if (LOG && SM != 0) {
      qk_async_wait();
      item = S32(c.o32 + S32_HEADITEM);
    } else if (head_item) {
      item = *head_item; /* requested ahead of the call (qk_prefetch_record) */
    } else {
      item = C32(c.olog64 + head);
    }
    if (pay_out) {
      const uint32_t op = cx.comps[sidx].ocold64 + 3u * head;
      pay_out[0] = C64(op);
      pay_out[1] = C64(op + 1);
      pay_out[2] = C64(op + 2);
    }
    head += 1;
    if (head >= sb.ring_cap) head = 0;
    cnt -= 1;
    cur = qk_stock_cat(cnt, sb.low, sb.max);
    S32(c.o32 + S32_HC) = QK_HC_PACK(head, cnt, cur);
    if (LOG && SM != 0 && cnt != 0) qk_stage4(&S32(c.o32 + S32_HEADITEM), &C32(c.olog64 + head));
    QK_ST64_ADD(sidx, ST64_TIME, cx.now);
  }
  const uint64_t id = qk_log<LOG, SM>(cx, c, sidx, src, QK_LOG_STOCK_REMOVE, 0, some ? (uint64_t)item : QK_ITEM_NONE);
  if (pay_out && some) qk_log_pay<LOG, SM>(cx, c.flags, sidx, id, pay_out);
  *item_out = item;
  uint32_t rc = some ? QK_RM_GOT : 0u;
  if (prev != cur &&
      qk_emit_push<LOG, SM, true>(cx, c, sb.emit_depth, c.olog64 + sb.ring_cap, (uint32_t)id))
    rc |= QK_RM_FAULT;
  return rc;
}

/* state category through DevHotB.up_hc / down_hc (3 = not connected): one state word, no table access */
template <int SM>
__device__ __forceinline__ uint32_t qk_cat_at(Ctx& cx, uint32_t hc_slot) {
  return hc_slot == QK_HC_NONE ? 3u : QK_HC_CAT(S32(hc_slot));
}

/* state category of stock `sidx` (3 = not connected): one table quad + one state word */
template <int SM>
__device__ __forceinline__ uint32_t qk_stock_cat_of(Ctx& cx, int sidx) {
  if (sidx < 0) return 3u;
  const DevHotSB sb = QK_HOT_SB(sidx);
  return QK_HC_CAT(S32(sb.o32 + S32_HC));
}

/* ---- DelayModes::update_state (delays.rs:86-142) + DelayEnd/DelayStart logs (discrete.rs:441-451) ------- */
template <bool LOG, int SM>
__device__ __forceinline__ void qk_tick_delays(Ctx& cx, const DevHotA& c, const DevHotB& cb, uint32_t comp,
                                               uint32_t* flags_io, uint64_t dt, uint64_t* src) {
  uint32_t flags = *flags_io;
  uint32_t mask = (flags >> PF_FIX_SHIFT) & 0xFFu;
  int from = -1, to = -1;
  if (mask) {
    const int a = __ffs((int)mask) - 1; /* active_delay: first mode in TimeUntilFix, insertion order */
    uint64_t dur = W64(cb.o64_dm + a);
    const uint64_t applied = dt < dur ? dt : dur;
    QK_ST64_ADD(comp, ST64_DELAY, applied);
    dur -= applied;
    if (dur == 0) {
      dur = qk_sample_duration<SM>(cx, cx.dms[cb.dm_off + a].until_delay);
      if (cx.status) return;
      mask &= ~(1u << a);
    }
    W64(cb.o64_dm + a) = dur;
    from = a;
  } else {
    for (int k = 0; k < c.n_dm; ++k) {
      const uint64_t dur = W64(cb.o64_dm + k);
      W64(cb.o64_dm + k) = dur - (dt < dur ? dt : dur);
    }
  }
  if (mask) {
    to = __ffs((int)mask) - 1;
  } else {
    for (int k = 0; k < c.n_dm; ++k) {
      if (W64(cb.o64_dm + k) == 0) {
        const uint64_t dur = qk_sample_duration<SM>(cx, cx.dms[cb.dm_off + k].until_fix);
        if (cx.status) return;
        W64(cb.o64_dm + k) = dur;
        mask |= 1u << k;
        to = k;
        break;
      }
    }
  }
  flags = (flags & ~(0xFFu << PF_FIX_SHIFT)) | (mask << PF_FIX_SHIFT);
  *flags_io = flags;
  if (from != to) { /* DelayStateTransition::has_changed delays.rs:35-42 */
    if (from >= 0) *src = qk_log<LOG, SM>(cx, c, comp, *src, QK_LOG_DELAY_END, 0, (uint64_t)from);
    if (to >= 0) *src = qk_log<LOG, SM>(cx, c, comp, *src, QK_LOG_DELAY_START, 0, (uint64_t)to);
  }
}

/* "Update cached environment state" (discrete.rs:455-471 and siblings) */
template <bool LOG, int SM>
__device__ __forceinline__ void qk_refresh_env(Ctx& cx, const DevHotA& c, int env, uint32_t comp, uint32_t* flags,
                                               uint64_t* src) {
  const bool new_stopped = env >= 0 ? (W32(QK_HOT_SB(env).o32 + E32_STATE) == 1u) : false;
  const bool old_stopped = (*flags & PF_ENV_STOPPED) != 0;
  if (!old_stopped && new_stopped) {
    *src = qk_log<LOG, SM>(cx, c, comp, *src, QK_LOG_PROCESS_STOPPED, QK_REASON_STOPPED_BY_ENV, 0);
    *flags |= PF_ENV_STOPPED;
  } else if (old_stopped && !new_stopped) {
    *src = qk_log<LOG, SM>(cx, c, comp, *src, QK_LOG_PROCESS_CONTINUE, QK_REASON_RESUMED_BY_ENV, 0);
    *flags &= ~PF_ENV_STOPPED;
  }
}

/* post_update_state (discrete.rs:535-564): keep the earlier of the pending keyed event and now+ttnpe.
 * Cancelling == overwriting the single fire slot. */
/* LEAN (feature set 0, SIMPLE components only): a keyed event is only ever scheduled right after the component logged
 * its ProcessStart, and a busy SIMPLE component logs nothing until that event fires (qk_try_quick), so the event's
 * source id is always (component, seq counter - 1) and need not be stored */
template <bool LOG, int SM, bool LEAN = false>
__device__ __forceinline__ bool qk_post(Ctx& cx, const DevHotA& c, uint64_t ttnpe, uint64_t src) {
  if (ttnpe != QK_TIME_NONE) {
    if (ttnpe == 0) {
      cx.status = QK_REP_ZERO_DURATION; /* panic!("Time until next event is zero!") */
      return true;
    }
    const uint64_t next = cx.now + ttnpe;
    if (next < qk_fire_get<SM>(cx, c.fire_idx)) { /* NONE is the largest u64 */
      qk_fire_set<SM>(cx, c.fire_idx, next);
      if (LOG && !LEAN) C64(c.olog64 + PL64_SCHED_SRC) = src;
    }
  }
  return false;
}

/* pre_update_state (discrete.rs:404-412): a handle whose time has come is forgotten, NOT cancelled; if its
 * event has not fired yet it stays queued (Q3) -> move it to the stale mask (it fires at `now`, in rank order) */
template <bool LOG, int SM, bool LEAN = false>
__device__ __forceinline__ void qk_pre(Ctx& cx, const DevHotA& c, uint32_t comp) {
  if constexpr (qk_is_two_tier(SM)) {
    /* nothing pending, or (cached minimum) nothing of any component due yet: no need to fetch the slot */
    if (!qk_fire_busy<SM>(cx, c.fire_idx) || (cx.kvalid && cx.kbest > cx.now)) return;
  }
  if (qk_fire_get<SM>(cx, c.fire_idx) <= cx.now) {
    qk_fire_clear<SM>(cx, c.fire_idx);
    S32(G32_STALE0 + (c.fire_idx >> 5)) |= 1u << (c.fire_idx & 31);
    if (LOG)
      C64(c.olog64 + PL64_STALE_SRC) =
          LEAN ? (((uint64_t)comp << 32) | (uint32_t)(S32(c.olog32 + L32_SEQ) - 1u)) : C64(c.olog64 + PL64_SCHED_SRC);
  }
}

/* ---- quick path ----------------------------------------------------------------------------------------------
 * Most update_state calls change nothing: a stock's emit_change is broadcast to every connected process, and a
 * process that is busy, or idle and still blocked, has nothing to do.  The event loop disposes of those calls while
 * it walks a call list, so the UPDATE phase -- the long, divergent handler -- only sees calls that finish or start
 * something.  Quick handling is limited to SIMPLE components (single-server, no delay modes, no environment), for
 * which two invariants hold between calls:
 *     busy  <=>  fire != NONE, and then fire == previous_check_time + left (the finish time);
 *     idle   =>  the sink's persistent ttnpe is NONE.
 * (left, previous_check_time) is only ever used as the pair "left at previous_check_time", so a busy no-op call may
 * leave it untouched: the next real call subtracts the whole elapsed time at once (saturating subtraction composes),
 * and statistics read-outs evaluate the countdown at the clock (qk_left_at_clock).  Hence a quick call of a busy
 * component is one table quad + one state load, and it writes nothing.
 * Returns false (state untouched) when the call needs the full handler. */
/* Quick handling of a single-server discrete component WITH delay modes (SQ_DM; no environment).  Its update_state
 * (discrete.rs:404-564) changes nothing but countdowns in two frequent cases:
 *   idle   -- no item, no mode in TimeUntilFix, no pending keyed event: the delay modes do not tick (discrete.rs:436,
 *             `is_in_delay || is_in_process`), so the call is the start test alone; if the component stays blocked it
 *             logs ProcessNonStart and nothing else (previous_check_time is not read again before the next start, which
 *             rewrites it);
 *   busy   -- an item in process, no mode in TimeUntilFix, the pending keyed event neither due nor later than the finish
 *             time, and dt shorter than every until-delay countdown: the process countdown and the until-delay countdowns
 *             lose dt, previous_check_time = now, no record, and post_update_state keeps the pending event
 *             (now + left is not earlier than it).  Not for sinks, whose ttnpe is not recomputed while busy
 *             (discrete.rs:1127-1129).
 * Returns 0 = needs the full handler, 1 = idle and blocked (reason in *reason), 2 = busy: *dt_out to be applied with
 * qk_dm_apply.  Reads only. */
template <int SM>
__device__ __forceinline__ uint32_t qk_dm_quick(Ctx& cx, const DevSubQ& q, uint32_t comp, uint64_t now, uint32_t* reason,
                                                uint64_t* dt_out, uint64_t* left_out = nullptr, uint64_t* dm0_out = nullptr) {
  const DevHotA a = QK_HOT_A(comp);
  const uint32_t flags = W32(a.o32 + P32_FLAGS);
  const uint64_t fire = qk_is_two_tier(SM) ? qk_fire_get<SM>(cx, q.fire_slot) : S64(q.fire_slot);
  /* record in global memory (split placements, HBM planes): the countdown pair and the smallest until-delay countdown are
   * requested together with the flags -- one trip to L2 for the whole test instead of three dependent ones (38 % of the
   * split kernel's stall samples on the 67-component line sat on these lines) */
  constexpr bool EAGER_REC = qk_is_split(SM) || SM == 0;
  uint64_t e_prev = 0, e_left = 0, e_mindm = QK_TIME_NONE;
  if constexpr (EAGER_REC) {
    e_prev = W64(a.o64 + P64_PREV);
    e_left = W64(a.o64 + P64_LEFT);
    const DevHotB eb = QK_HOT_B(comp);
    for (uint32_t k = 0; k < a.n_dm; ++k) {
      const uint64_t d = W64(eb.o64_dm + k);
      e_mindm = d < e_mindm ? d : e_mindm;
    }
  }
  if (flags & ~PF_HAS_ITEM) return 0u; /* a mode in TimeUntilFix (or environment-stopped) */
  if (!(flags & PF_HAS_ITEM)) {
    if (fire != QK_TIME_NONE) return 0u;
    const uint32_t us = (q.flags & SQ_US_FORCED) ? (q.flags >> SQ_US_SHIFT) & 3u : QK_HC_CAT(S32(q.up_hc));
    const uint32_t ds = (q.flags & SQ_DS_FORCED) ? (q.flags >> SQ_DS_SHIFT) & 3u : QK_HC_CAT(S32(q.down_hc));
    if ((us == 1u || us == 2u) && (ds == 0u || ds == 1u)) return 0u; /* starts */
    *reason = us == 0u   ? QK_REASON_UPSTREAM_EMPTY
              : us == 3u ? QK_REASON_UPSTREAM_NOT_CONNECTED
              : ds == 2u ? QK_REASON_DOWNSTREAM_FULL
                         : QK_REASON_DOWNSTREAM_NOT_CONNECTED;
    return 1u;
  }
  if (a.kind == QK_DISCRETE_SINK || fire == QK_TIME_NONE || fire <= now) return 0u;
  const uint64_t prev = EAGER_REC ? e_prev : W64(a.o64 + P64_PREV), left = EAGER_REC ? e_left : W64(a.o64 + P64_LEFT);
  if (prev + left < fire) return 0u; /* post_update_state would move the keyed event */
  const uint64_t dt = now - prev;
  if constexpr (EAGER_REC) {
    if (e_mindm <= dt) return 0u; /* a breakdown starts in this call */
    if (left_out) *left_out = left; /* what qk_dm_apply is about to rewrite: it need not fetch them again */
    if (dm0_out) *dm0_out = e_mindm;
  } else {
    const DevHotB b = QK_HOT_B(comp);
    for (uint32_t k = 0; k < a.n_dm; ++k)
      if (W64(b.o64_dm + k) <= dt) return 0u; /* a breakdown starts in this call */
  }
  *dt_out = dt;
  return 2u;
}
template <int SM>
__device__ __forceinline__ void qk_dm_apply(Ctx& cx, uint32_t comp, uint64_t now, uint64_t dt,
                                            const uint64_t* left_known = nullptr, const uint64_t* dm0_known = nullptr) {
  const DevHotA a = QK_HOT_A(comp);
  const DevHotB b = QK_HOT_B(comp);
  if (dm0_known && a.n_dm == 1) W64(b.o64_dm) = *dm0_known - dt; /* (one mode: its countdown is the minimum just tested) */
  else for (uint32_t k = 0; k < a.n_dm; ++k) W64(b.o64_dm + k) -= dt;
  if (left_known) W64(a.o64 + P64_LEFT) = *left_known - dt;
  else W64(a.o64 + P64_LEFT) -= dt;
  W64(a.o64 + P64_PREV) = now;
#ifdef QK_HOST_EMUL
  qk_emul_count(0); /* the CPU suite checks that its delay-mode models actually take this path */
#endif
}

template <bool LOG, int SM, int FT>
__device__ __forceinline__ bool qk_try_quick(Ctx& cx, uint32_t ptr, uint64_t src) {
  union {
    uint2 v;
    DevSubQ q;
  } u;
  u.v = *reinterpret_cast<const uint2*>(&cx.subq[ptr]); /* one LDS.64 */
  const DevSubQ q = u.q;
  if (FT != 0 && !(q.flags & SQ_QUICK)) {
    if (!(q.flags & SQ_DM)) return false;
    const uint32_t comp = cx.subs[ptr];
    uint32_t reason = 0;
    uint64_t dt = 0;
    uint64_t left_k = 0, dm0_k = 0;
    constexpr bool KNOWN = qk_is_split(SM) || SM == 0; /* qk_dm_quick fetched them */
    const uint32_t code = qk_dm_quick<SM>(cx, q, comp, cx.now, &reason, &dt, &left_k, &dm0_k);
    if (code == 0u) return false;
    if (code == 2u)
      qk_dm_apply<SM>(cx, comp, cx.now, dt, KNOWN ? &left_k : nullptr, KNOWN ? &dm0_k : nullptr);
    else if (LOG)
      qk_log<LOG, SM>(cx, QK_HOT_A(comp), comp, src, QK_LOG_PROCESS_NONSTART, reason, 0);
#ifdef QK_HOST_EMUL
    if (code == 1u) qk_emul_count(1);
#endif
    return true;
  }
  /* the three state words are independent loads: one shared-memory latency, not three.  (On the HBM planes the
   * neighbour words are only fetched when the callee turns out to be idle: there a load is a memory transaction; the
   * logging kernels do the same because they have no registers to hold the two words: measured.) */
  constexpr bool EAGER = SM != 0 && !LOG;
  uint64_t fire;
  if constexpr (qk_is_two_tier(SM)) {
    /* two-tier queue: q.fire_slot is the callee's rank.  Busy (bit set) and, by the cached minimum, nothing of ANY
     * component due yet: busy and not due without fetching the slot; only at a timestamp tie, or with the cache
     * invalid, the slot itself is read */
    if (!qk_fire_busy<SM>(cx, q.fire_slot)) fire = QK_TIME_NONE;
    else if (cx.kvalid && cx.kbest > cx.now) fire = cx.kbest; /* some time > now: all the code below asks */
    else fire = QK_CFIRE(q.fire_slot);
#ifdef QK_HOST_EMUL
    if (qk_fire_busy<SM>(cx, q.fire_slot)) qk_emul_count((cx.kvalid && cx.kbest > cx.now) ? 2 : 3);
#endif
  } else {
    fire = S64(q.fire_slot);
  }
  uint32_t uw = 0, dw = 0;
  if constexpr (EAGER) {
    uw = S32(q.up_hc);
    dw = S32(q.down_hc);
  }
  if (fire != QK_TIME_NONE) {
#ifdef QK_HOST_EMUL
    const DevHotA c = QK_HOT_A(cx.subs[ptr]);
    const uint64_t fx = qk_is_two_tier(SM) ? QK_CFIRE(q.fire_slot) : fire; /* the slot itself */
    if (qk_is_two_tier(SM) && (fx > cx.now) != (fire > cx.now)) cx.status = 97; /* the cached minimum lied */
    if (fx > cx.now && (!(W32(c.o32 + P32_FLAGS) & PF_HAS_ITEM) ||
                        W64(c.o64 + P64_PREV) + W64(c.o64 + P64_LEFT) != fx))
      cx.status = 98; /* invariant check, CPU test build only */
#endif
    return fire > cx.now; /* busy and not due: nothing to do ; due (<= now): pre_update_state drops the handle (Q3) */
  }
#ifdef QK_HOST_EMUL
  if (W32(QK_HOT_A(cx.subs[ptr]).o32 + P32_FLAGS) & PF_HAS_ITEM) cx.status = 98;
#endif
  /* idle: quick only if it stays idle */
  if constexpr (!EAGER) {
    uw = S32(q.up_hc);
    dw = S32(q.down_hc);
  }
  const uint32_t us = (q.flags & SQ_US_FORCED) ? (q.flags >> SQ_US_SHIFT) & 3u : QK_HC_CAT(uw);
  const uint32_t ds = (q.flags & SQ_DS_FORCED) ? (q.flags >> SQ_DS_SHIFT) & 3u : QK_HC_CAT(dw);
  if ((us == 1u || us == 2u) && (ds == 0u || ds == 1u)) return false; /* starts */
  if (LOG) {
    /* a source's forced upstream category is 1, so the first two tests can only fire for components that need one */
    const uint32_t reason = us == 0u   ? QK_REASON_UPSTREAM_EMPTY
                            : us == 3u ? QK_REASON_UPSTREAM_NOT_CONNECTED
                            : ds == 2u ? QK_REASON_DOWNSTREAM_FULL
                                       : QK_REASON_DOWNSTREAM_NOT_CONNECTED;
    const uint32_t comp = cx.subs[ptr];
    qk_log<LOG, SM>(cx, QK_HOT_A(comp), comp, src, QK_LOG_PROCESS_NONSTART, reason, 0);
  }
  return true;
}

/* ---- component record of a single-server discrete component, requested ahead of its update_state ------------------
 * With the records in global memory (split placements, HBM planes) a finish-and-restart was a chain of dependent trips to
 * L2 / DRAM -- flags and countdown, then the item, then the pre-drawn duration, then the head of the upstream ring -- 40 %
 * of the split kernel's stall samples.  All of it is known by address once the lane knows which component its next call
 * is for, so the event loop asks for everything at once, BEFORE the REFILL phase (whose own trip for the draw counters
 * then overlaps), and the handler finds the values in registers.  Nothing but the REFILL phase runs in between, and that
 * writes pre-drawn durations only: an EMPTY one is fetched again. */
struct QkPre {
  uint32_t comp, flags, item, idx, up_item;
  uint64_t left, prev, dur;
  bool has, has_up;
};
template <int SM>
__device__ __forceinline__ void qk_prefetch_record(Ctx& cx, uint32_t comp, QkPre& pre) {
  const DevHotA c = QK_HOT_A(comp);
  if (c.kind < QK_DISCRETE_PROCESS || c.kind > QK_DISCRETE_SINK) return;
  const DevHotB cb = QK_HOT_B(comp);
  pre.comp = comp;
  pre.has = true;
  pre.flags = W32(c.o32 + P32_FLAGS);
  pre.left = W64(c.o64 + P64_LEFT);
  pre.prev = W64(c.o64 + P64_PREV);
  pre.item = W32(c.o32 + P32_ITEM);
  if (cb.o64_nextdur != 0xFFFF) pre.dur = W64(cb.o64_nextdur);
  if (c.kind == QK_DISCRETE_SOURCE) pre.idx = W32(QK_HOT_C(comp).o32_next_item);
  /* the item it would withdraw: the head of the upstream ring, if there is one now (a ring that is empty now can only
   * receive its head from this very call -- a process whose downstream is its upstream -- and is then read in place) */
  if (c.kind != QK_DISCRETE_SOURCE && cb.up_hc != QK_HC_NONE) {
    const uint32_t hc = S32(cb.up_hc);
    if (QK_HC_CNT(hc) != 0) {
      pre.has_up = true;
      pre.up_item = C32(QK_HOT_A((uint32_t)cx.comps[comp].up).olog64 + QK_HC_HEAD(hc));
    }
  }
}

/* ---- DiscreteProcess / DiscreteSource / DiscreteSink update_state, one code path ----------------------------- */
template <bool LOG, int SM, int FT>
__device__ __forceinline__ bool qk_update_single(Ctx& cx, const DevHotA& c, uint32_t comp, uint64_t src,
                                                 const QkPre* pre = nullptr) {
  const DevHotB cb = QK_HOT_B(comp);
  const uint32_t kind = c.kind;
  /* component record in global memory (split placements, HBM planes): the event loop has requested it ahead of the call
   * (qk_prefetch_record); a caller without a prefetch gets the same loads here, side by side */
  constexpr bool REC_GLOBAL = qk_is_split(SM) || SM == 0;
  QkPre mine;
  mine.has = false;
  mine.has_up = false;
  mine.dur = 0;
  mine.idx = 0;
  if constexpr (REC_GLOBAL) {
    if (pre && pre->has && pre->comp == comp) {
      mine = *pre;
      if (cb.o64_nextdur != 0xFFFF && mine.dur == QK_DUR_EMPTY) mine.dur = W64(cb.o64_nextdur); /* refilled since */
    } else {
      qk_prefetch_record<SM>(cx, comp, mine);
    }
  }
  uint32_t flags = REC_GLOBAL ? mine.flags : W32(c.o32 + P32_FLAGS);
  uint64_t left = REC_GLOBAL ? mine.left : W64(c.o64 + P64_LEFT);
  const uint64_t dt = cx.now - (REC_GLOBAL ? mine.prev : W64(c.o64 + P64_PREV));
  const uint32_t pre_item = mine.item, pre_idx = mine.idx;
  const uint64_t pre_dur = mine.dur;
  uint64_t ttnpe = QK_TIME_NONE;
  /* Vector3ContainerProcess / Source / Sink (core.rs:550-553): the item carries a resource, kept at cold64 `opay` */
  const bool ctr = FT == 2 && (c.flags & DC_CONTAINER) != 0;
  const uint32_t opay = ctr ? cx.comps[comp].ocold64 : 0u;
  uint64_t pay[3];
  {
    const bool is_in_delay = ((flags >> PF_FIX_SHIFT) & 0xFFu) != 0;
    const bool has_item0 = (flags & PF_HAS_ITEM) != 0;
    const bool is_in_process = has_item0 && !is_in_delay;
    const bool is_env_blocked = (flags & PF_ENV_STOPPED) != 0;
    if (!(is_in_delay || is_env_blocked) && has_item0) {
      left -= (dt < left ? dt : left); /* saturating_sub */
      if (left == 0) {
        flags &= ~PF_HAS_ITEM;
        const uint32_t item = REC_GLOBAL ? pre_item : W32(c.o32 + P32_ITEM);
        src = qk_log<LOG, SM>(cx, c, comp, src, QK_LOG_PROCESS_FINISH, 0, item);
        if (ctr) {
          pay[0] = C64(opay);
          pay[1] = C64(opay + 1);
          pay[2] = C64(opay + 2);
          qk_log_pay<LOG, SM>(cx, c.flags, comp, src, pay);
        }
        QK_ST32_ADD(comp, ST32_COUNT, 1u);
        if (kind != QK_DISCRETE_SINK && cb.down_hc != QK_HC_NONE) { /* push_downstream: no downstream check at finish */
          if (qk_stock_add<LOG, SM>(cx, (uint32_t)cx.comps[comp].down, item, src, ctr ? pay : nullptr)) return true;
        }
      }
    }
    if (FT != 0 && c.n_dm && !is_env_blocked && (is_in_delay || is_in_process)) {
      qk_tick_delays<LOG, SM>(cx, c, cb, comp, &flags, dt, &src);
      if (cx.status) return true;
    }
  }
  if (FT != 0 && (cb.env >= 0 || (flags & PF_ENV_STOPPED))) qk_refresh_env<LOG, SM>(cx, c, cb.env, comp, &flags, &src);

  const bool has_item = (flags & PF_HAS_ITEM) != 0;
  const uint32_t fixmask = FT != 0 ? (flags >> PF_FIX_SHIFT) & 0xFFu : 0u;
  const bool has_active_delay = FT != 0 && (fixmask != 0 || (flags & PF_ENV_STOPPED) != 0);
  if (!has_item && !has_active_delay) {
    /* match (&us_state, &ds_state) discrete.rs:480-516 / 854-872 / 1100-1125; 3 == not connected */
    const bool need_us = kind != QK_DISCRETE_SOURCE, need_ds = kind != QK_DISCRETE_SINK;
    const uint32_t us = need_us ? qk_cat_at<SM>(cx, cb.up_hc) : 1u;
    const uint32_t ds = need_ds ? qk_cat_at<SM>(cx, cb.down_hc) : 1u;
    const bool ok = (us == 1u || us == 2u) && (ds == 0u || ds == 1u);
    if (ok) {
      bool got = true;
      uint32_t item;
      if (need_us) {
        src = qk_log<LOG, SM>(cx, c, comp, src, QK_LOG_WITHDRAW_REQUEST, 0, 0);
        const uint32_t rc = qk_stock_remove<LOG, SM>(cx, (uint32_t)cx.comps[comp].up, src, &item, ctr ? pay : nullptr,
                                                     REC_GLOBAL && mine.has_up ? &mine.up_item : nullptr);
        if (rc & QK_RM_FAULT) return true;
        got = (rc & QK_RM_GOT) != 0;
      }
      if (got) {
        /* process_time_distr.sample() + Duration::from_secs_f64: Constant distributions were converted at
         * lowering; random ones were drawn ahead of time by the REFILL phase of the event loop */
        const DevHotC cc = QK_HOT_C(comp);
        uint64_t d;
        if (cb.o64_nextdur != 0xFFFF) {
          d = REC_GLOBAL ? pre_dur : W64(cb.o64_nextdur);
          W64(cb.o64_nextdur) = QK_DUR_EMPTY;
          S32(cx.pf32 + (cb.pf_idx >> 5)) |= 1u << (cb.pf_idx & 31);
          cx.npend += 1;
        } else {
          d = ((uint64_t)cc.const_dur_hi << 32) | cc.const_dur_lo;
        }
        if (!need_us) { /* ItemFactory::create_item discrete.rs:680-686 */
          const uint32_t idx = REC_GLOBAL ? pre_idx : W32(cc.o32_next_item);
          if (idx > 0x3FFFFFFu) {
            cx.status = QK_REP_ITEM_INDEX_OVERFLOW;
            return true;
          }
          W32(cc.o32_next_item) = idx + 1;
          item = ((uint32_t)c.src_ord << 26) | idx;
          if (ctr) { /* Vector3ContainerFactory::create_item (vector_container.rs:136-156) */
            const DevVec* v = &cx.vecs[cc.vec_idx];
            pay[0] = QK_CONTAINER_NONE_BITS;
            pay[1] = pay[2] = 0;
            if (v->dist_qty != 0xFFFF) {
              const double q = qk_sample<SM>(cx, v->dist_qty);
              if (cx.status) return true;
              for (int k = 0; k < 3; ++k) pay[k] = QK_D2U(__dmul_rn(q, v->vinit[k])); /* quantity * split[k] */
            }
          }
        }
        if (d >= QK_DUR_FAULT_BASE) { /* from_secs_f64 panicked / tape exhausted / (EMPTY: refill logic bug) */
          cx.status = d == QK_DUR_EMPTY ? 99u : (uint32_t)(d - QK_DUR_FAULT_BASE);
          return true;
        }
        left = d;
        flags |= PF_HAS_ITEM;
        W32(c.o32 + P32_ITEM) = item;
        QK_ST64_ADD(comp, ST64_TIME, d);
        src = qk_log<LOG, SM>(cx, c, comp, src, QK_LOG_PROCESS_START, 0, item);
        if (ctr) {
          C64(opay) = pay[0];
          C64(opay + 1) = pay[1];
          C64(opay + 2) = pay[2];
          qk_log_pay<LOG, SM>(cx, c.flags, comp, src, pay);
        }
        ttnpe = d;
      } else {
        src = qk_log<LOG, SM>(cx, c, comp, src, QK_LOG_PROCESS_NONSTART, QK_REASON_UPSTREAM_NO_RESOURCE, 0);
      }
    } else {
      uint32_t reason;
      if (need_us && us == 0u)
        reason = QK_REASON_UPSTREAM_EMPTY;
      else if (need_us && us == 3u)
        reason = QK_REASON_UPSTREAM_NOT_CONNECTED;
      else if (ds == 2u)
        reason = QK_REASON_DOWNSTREAM_FULL;
      else
        reason = QK_REASON_DOWNSTREAM_NOT_CONNECTED;
      src = qk_log<LOG, SM>(cx, c, comp, src, QK_LOG_PROCESS_NONSTART, reason, 0);
    }
  } else if (kind == QK_DISCRETE_SINK) {
    /* (Some, _) leaves ttnpe untouched (discrete.rs:1127-1129); (_, true) clears it (Q2, :1130-1132) */
    ttnpe = has_item ? W64(QK_HOT_C(comp).o64_ttnpe) : QK_TIME_NONE;
  } else if (has_item && (!has_active_delay || kind == QK_DISCRETE_SOURCE)) {
    ttnpe = left; /* Process (Some,false) :518-520 ; Source (Some,_) :874-876 */
  } else {
    ttnpe = fixmask ? W64(cb.o64_dm + (__ffs((int)fixmask) - 1)) : QK_TIME_NONE; /* (_, true) :521-523 */
  }
  if (kind == QK_DISCRETE_SINK) W64(QK_HOT_C(comp).o64_ttnpe) = ttnpe;
  W32(c.o32 + P32_FLAGS) = flags;
  W64(c.o64 + P64_LEFT) = left;
  const bool fault = qk_post<LOG, SM, FT == 0>(cx, c, ttnpe, src);
  W64(c.o64 + P64_PREV) = cx.now;
  return fault;
}

#include "qk_vector.cuh"

/* ---- multi-server process-like components ------------------------------------------------------------------
 * DiscreteParallelProcess::update_state_impl (discrete.rs:1280-1417), ContainerLoadingProcess (vector_container.rs:
 * 263-393) and ContainerUnloadingProcess (vector_container.rs:550-707) share one skeleton: a Vec of (countdown, item)
 * in progress, a deque of completed items drained while downstream accepts, a withdraw loop, ttnpe = min countdown.
 * They differ in the drain condition, in what a finish/start moves, and in the withdraw test; `kind` selects. */
template <bool LOG, int SM, int FT>
__device__ __forceinline__ void qk_update_multi(Ctx& cx, const DevComp* c, uint32_t comp, uint64_t src) {
  const DevHotA ha = QK_HOT_A(comp);
  const DevHotB hb = QK_HOT_B(comp);
  const uint32_t kind = c->kind;
  /* the container family is compiled in only for models that have it (feature set 2) */
  const bool ctr = FT == 2 && (c->flags & DC_CONTAINER) != 0;
  const bool is_load = FT == 2 && kind == QK_CONTAINER_LOAD_PROCESS;
  const bool is_unload = FT == 2 && kind == QK_CONTAINER_UNLOAD_PROCESS;
  uint32_t flags = W32(c->o32 + PP32_FLAGS);
  const uint64_t dt = cx.now - W64(c->o64 + PP64_PREV);
  const uint32_t cap = c->ring_cap;
  const uint32_t dur0 = c->o64_dm + c->n_dm;
  const uint32_t item0 = c->o32 + PP32_BASE_COUNT;
  const uint32_t comp0 = item0 + cap;
  const uint32_t pay0 = c->ocold64, cpay0 = c->ocold64 + 3u * cap; /* container payloads: in progress / complete ring */
  uint32_t nprog = W32(c->o32 + PP32_NPROG);
  uint32_t cq = W32(c->o32 + PP32_COMPLETE);
  uint32_t chead = cq & 0xFFFFu, ccnt = cq >> 16;
  uint64_t pay[3];
  {
    const bool is_in_delay = ((flags >> PF_FIX_SHIFT) & 0xFFu) != 0;
    const bool is_in_process = nprog > 0 && !is_in_delay;
    const bool is_env_blocked = (flags & PF_ENV_STOPPED) != 0;
    if (!(is_in_delay || is_env_blocked)) {
      uint32_t w = 0;
      for (uint32_t i = 0; i < nprog; ++i) { /* retain_mut discrete.rs:1298-1306 */
        uint64_t d = W64(dur0 + i);
        const uint32_t it = W32(item0 + i);
        d -= (dt < d ? dt : d);
        if (d == 0) {
          if (ccnt >= cap) {
            cx.status = QK_REP_PARALLEL_OVERFLOW;
            return;
          }
          uint32_t pos = chead + ccnt;
          if (pos >= cap) pos -= cap;
          W32(comp0 + pos) = it;
          if (ctr)
            for (uint32_t k = 0; k < 3; ++k) C64(cpay0 + 3u * pos + k) = C64(pay0 + 3u * i + k);
          ccnt += 1;
        } else {
          W64(dur0 + w) = d;
          W32(item0 + w) = it;
          if (ctr && w != i)
            for (uint32_t k = 0; k < 3; ++k) C64(pay0 + 3u * w + k) = C64(pay0 + 3u * i + k);
          w += 1;
        }
      }
      nprog = w;
      while (ccnt) { /* discrete.rs:1311-1327 -- the popped item is lost when downstream does not accept it */
        const uint32_t it = W32(comp0 + chead);
        if (ctr)
          for (uint32_t k = 0; k < 3; ++k) pay[k] = C64(cpay0 + 3u * chead + k);
        chead += 1;
        if (chead >= cap) chead = 0;
        ccnt -= 1;
        const uint32_t ds = c->down >= 0 ? qk_stock_cat_of<SM>(cx, c->down) : 3u;
        if (!is_unload) {
          if (ds == 0u || ds == 1u) {
            src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_PROCESS_FINISH, 0, it);
            if (ctr) qk_log_pay<LOG, SM>(cx, ha.flags, comp, src, pay);
            QK_ST32_ADD(comp, ST32_COUNT, 1u);
            qk_stock_add<LOG, SM>(cx, (uint32_t)c->down, it, src, ctr ? pay : nullptr);
            if (cx.status) return;
          } else {
            src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_PROCESS_NONSTART,
                                  ds == 2u ? QK_REASON_DOWNSTREAM_FULL : QK_REASON_DOWNSTREAM_NOT_CONNECTED, 0);
            break;
          }
        } else { /* vector_container.rs:574-613 */
          const DevVec* v = &cx.vecs[c->vec_idx];
          const uint32_t dsr = qk_vstock_cat_of<SM>(cx, v->downs[0]);
          if ((ds == 0u || ds == 1u) && (dsr == 0u || dsr == 1u)) {
            src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_PROCESS_FINISH, 0, it); /* the container as it arrived */
            qk_log_pay<LOG, SM>(cx, ha.flags, comp, src, pay);
            QK_ST32_ADD(comp, ST32_COUNT, 1u);
            const double proportion = qk_sample<SM>(cx, v->dist_qty); /* process_quantity_ratio_distr */
            if (cx.status) return;
            if (pay[0] != QK_CONTAINER_NONE_BITS) {
              /* item.take_resource() (:582) leaves item.resource == None ... */
              double taken[3] = {QK_U2D(pay[0]), QK_U2D(pay[1]), QK_U2D(pay[2])};
              double moved[3] = {0.0, 0.0, 0.0};
              pay[0] = QK_CONTAINER_NONE_BITS;
              pay[1] = pay[2] = 0;
              if (!(proportion >= 1.0)) {
                qk_vremove(taken, __dmul_rn(qk_vtotal(taken, 3), proportion), 3, moved);
                for (int k = 0; k < 3; ++k) pay[k] = QK_D2U(taken[k]); /* item.set_resource(Some(resource)) :590 */
              }
              /* ... so for proportion >= 1 item.remove_all() (:585) yields Vector3::default(): zeros reach the
               * resource stock and what the container carried is dropped (reproduced on purpose) */
              qk_vstock_add<LOG, SM>(cx, (uint32_t)v->downs[0], moved, src);
              if (cx.status) return;
            }
            qk_stock_add<LOG, SM>(cx, (uint32_t)c->down, it, src, pay);
            if (cx.status) return;
          } else {
            const uint32_t reason = ds == 2u ? QK_REASON_DS_CONTAINERS_FULL
                                             : (dsr == 2u ? QK_REASON_DS_RESOURCE_FULL : QK_REASON_DOWNSTREAM_NOT_CONNECTED);
            src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_PROCESS_NONSTART, reason, 0);
            break;
          }
        }
      }
    }
    if (c->n_dm && !is_env_blocked && (is_in_delay || is_in_process)) {
      qk_tick_delays<LOG, SM>(cx, ha, hb, comp, &flags, dt, &src);
      if (cx.status) return;
    }
  }
  if (c->env >= 0 || (flags & PF_ENV_STOPPED)) qk_refresh_env<LOG, SM>(cx, ha, c->env, comp, &flags, &src);
  uint64_t ttnpe = QK_TIME_NONE;
  if (!(flags & PF_ENV_STOPPED)) {
    for (;;) { /* discrete.rs:1373-1398 ; vector_container.rs:348-381, 659-688 */
      const uint32_t us = c->up >= 0 ? qk_stock_cat_of<SM>(cx, c->up) : 3u;
      /* Parallel and Load withdraw while upstream is Empty|Normal (Q4); Unload while it is Normal|Full */
      const bool us_ok = is_unload ? (us == 1u || us == 2u) : (us == 0u || us == 1u);
      const bool slots = !(is_load || is_unload) || nprog + ccnt < c->max; /* max_process_count - in progress - complete */
      if (us_ok && slots) {
        src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_WITHDRAW_REQUEST, 0, 0);
        uint32_t it;
        const uint32_t rc = qk_stock_remove<LOG, SM>(cx, (uint32_t)c->up, src, &it, ctr ? pay : nullptr);
        if (rc & QK_RM_FAULT) return;
        if (!(rc & QK_RM_GOT)) break;
        if (is_load) { /* vector_container.rs:356-360 */
          const DevVec* v = &cx.vecs[c->vec_idx];
          const double proportion = qk_sample<SM>(cx, v->dist_qty); /* process_capacity_ratio_distr */
          if (cx.status) return;
          const bool has = pay[0] != QK_CONTAINER_NONE_BITS;
          double held[3] = {has ? QK_U2D(pay[0]) : 0.0, has ? QK_U2D(pay[1]) : 0.0, has ? QK_U2D(pay[2]) : 0.0};
          const uint32_t ord = it >> 26;
          const double capacity = cx.ccap[ord == 63u ? 64u + (it & 0x3FFFFFFu) : ord];
          const double quantity = __dmul_rn(__dsub_rn(capacity, has ? qk_vtotal(held, 3) : 0.0), proportion);
          if (v->ups[0] < 0) {
            cx.status = QK_REP_NOT_CONNECTED; /* withdraw_us_resource ... .next().unwrap() :358 */
            return;
          }
          double got3[3];
          qk_vstock_remove<LOG, SM>(cx, (uint32_t)v->ups[0], quantity, src, got3); /* no state check on that stock */
          if (cx.status) return;
          for (int k = 0; k < 3; ++k) pay[k] = QK_D2U(has ? __dadd_rn(held[k], got3[k]) : got3[k]); /* item.add :84-92 */
        }
        const uint64_t d = qk_sample_duration<SM>(cx, c->dist_time);
        if (cx.status) return;
        if (nprog >= cap) {
          cx.status = QK_REP_PARALLEL_OVERFLOW;
          return;
        }
        W64(dur0 + nprog) = d;
        W32(item0 + nprog) = it;
        if (ctr)
          for (uint32_t k = 0; k < 3; ++k) C64(pay0 + 3u * nprog + k) = pay[k];
        nprog += 1;
        QK_ST64_ADD(comp, ST64_TIME, d);
        src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_PROCESS_START, 0, it);
        if (ctr) qk_log_pay<LOG, SM>(cx, ha.flags, comp, src, pay);
      } else {
        uint32_t reason;
        if (!slots)
          reason = QK_REASON_NO_PROCESS_SLOTS; /* the (_, 0) arm precedes the state arms */
        else if (us == 3u)
          reason = QK_REASON_UPSTREAM_NOT_CONNECTED;
        else
          reason = is_unload ? QK_REASON_UPSTREAM_EMPTY : QK_REASON_UPSTREAM_FULL;
        src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_PROCESS_NONSTART, reason, 0);
        break;
      }
    }
    for (uint32_t i = 0; i < nprog; ++i) { /* discrete.rs:1400-1406 */
      const uint64_t d = W64(dur0 + i);
      if (d < ttnpe) ttnpe = d;
    }
  }
  W32(c->o32 + PP32_FLAGS) = flags;
  W32(c->o32 + PP32_NPROG) = nprog;
  W32(c->o32 + PP32_COMPLETE) = chead | (ccnt << 16);
  qk_post<LOG, SM>(cx, ha, ttnpe, src);
  W64(c->o64 + PP64_PREV) = cx.now;
}


/* Process::update_state (core.rs:370-376) */
/* returns true when the replication has faulted (cx.status) */
template <bool LOG, int SM, int FT>
__device__ __forceinline__ bool qk_update_state(Ctx& cx, uint32_t comp, uint64_t src, const QkPre* pre = nullptr) {
  const DevHotA a = QK_HOT_A(comp);
  if (FT == 0 || (a.kind >= QK_DISCRETE_PROCESS && a.kind <= QK_DISCRETE_SINK)) {
    qk_pre<LOG, SM, FT == 0>(cx, a, comp);
    return qk_update_single<LOG, SM, FT>(cx, a, comp, src, pre);
  }
  if constexpr (FT != 0) {
    const DevComp* c = &cx.comps[comp];
    if (a.kind == QK_ENVIRONMENT) { /* only reached from the init list: BasicEnvironment::init logs its state */
      qk_log<LOG, SM>(cx, a, comp, src, QK_LOG_ENV_STATE, W32(a.o32 + E32_STATE), 0);
      return cx.status != 0;
    }
    if (a.kind >= QK_VECTOR_PROCESS && a.kind <= QK_VECTOR_SPLITTER) {
      /* Combiner / Splitter / VectorSource / VectorSink init with their own next id instead of INIT_000000
       * (vector.rs:774,1111,1393,1619); SRC_INIT only ever arrives from the init list */
      if (LOG && src == SRC_INIT && a.kind != QK_VECTOR_PROCESS)
        src = ((uint64_t)comp << 32) | S32(a.olog32 + L32_SEQ);
      qk_pre<LOG, SM>(cx, a, comp);
      qk_update_vector<LOG, SM, qk_vd(FT)>(cx, c, comp, src);
      return cx.status != 0;
    }
    qk_pre<LOG, SM>(cx, a, comp);
    qk_update_multi<LOG, SM, FT>(cx, c, comp, src);
    return cx.status != 0;
  }
  return false;
}

/* ============================================ the kernel ============================================ */
#define QK_MAX_NT 256
#ifndef QK_SCAN_ITERS
#define QK_SCAN_ITERS 2 /* scheduler pops a lane may spend per trip looking for a call that needs the full handler (measured: 2 beats 1, 3, 4) */
#endif

constexpr int qk_lb_threads(int sm) { return sm > 0 ? sm : QK_MAX_NT; }
/* resident threads per SM the register allocation must allow: the kernel is latency-bound and its throughput is
 * proportional to resident warps (profiles/README.md), so the lean kernel (FT == 0: small models, ~250-280 B of hot
 * state per replication) is held to 96 registers = 20 warps (10 blocks of 64); the general kernels to 128
 * registers = 16 warps */
#ifndef QK_LEAN_THREADS
#define QK_LEAN_THREADS 640
#endif
#ifndef QK_LEAN_THREADS_LOG
#define QK_LEAN_THREADS_LOG QK_LEAN_THREADS /* the lean kernel with logging compiled in */
#endif
constexpr int qk_lb_resident(int ft, bool log) { return ft == 0 ? (log ? QK_LEAN_THREADS_LOG : QK_LEAN_THREADS) : 512; }
constexpr int qk_lb_blocks(int sm, int ft, bool log) {
  /* split placement: shared memory holds at most ~240 replications per SM, so registers are not what limits residency */
  /* ... but with the two-tier queue 500-600 replications of a lean model fit: two blocks of 256 */
  return sm == QK_SM_QSPLIT ? (ft == 0 ? 2 : 1) : qk_is_split(sm) ? 1 : sm > 0 ? (qk_lb_resident(ft, log) / sm > 0 ? qk_lb_resident(ft, log) / sm : 1) : (ft == 0 ? 3 : 2);
}
/* FT = feature set of the model: 0 = only SIMPLE single-server discrete components and stocks (no delay modes, no
 * environment, no scheduled events, no parallel or vector components): the handlers for everything else are
 * compiled out, which frees registers ; 1 = everything but the container family ; 2 = everything */
template <bool LOG, int SM, int FT>
__global__ void __launch_bounds__(qk_lb_threads(SM), qk_lb_blocks(SM, FT, LOG)) qk_step_kernel(const __grid_constant__ KParams P) {
  uint8_t* const smem = qk_smem;
  const uint32_t tid = QK_TID, nt = SM > 0 ? (uint32_t)SM : QK_NT;

  /* model tables -> shared memory (read with data-dependent component indices afterwards).  A launch that stages the
   * block's hot planes (every launch but the init sweep of an on-chip placement) brings the tables in by the same
   * bulk asynchronous copies, on the same mbarrier: the requests for tables and state leave together and the block
   * waits once -- as LDG/STS + __syncthreads in front of the staging the table copy was a dependent trip to L2 of its
   * own, 10 % of the stall samples of a one-event launch (profiles/README.md R2.3) */
#if !defined(QK_HOST_EMUL) && !defined(QK_TABLES_BY_LDG)
  constexpr bool BULK_PROLOGUE = SM != 0;
  __shared__ __align__(8) uint64_t stage_bar;
#else
  constexpr bool BULK_PROLOGUE = false;
#endif
  if (!(BULK_PROLOGUE && !P.do_init)) {
    const uint4* srcp = reinterpret_cast<const uint4*>(P.tables);
    uint4* dstp = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = tid; i < P.table_bytes / 16; i += nt) dstp[i] = srcp[i];
    __syncthreads();
  }
#if !defined(QK_HOST_EMUL) && !defined(QK_TABLES_BY_LDG)
  else {
    const uint64_t block_rep0 = (uint64_t)(P.tile0 + QK_BID) * nt;
    if (tid == 0) {
      qk_mbar_init(&stage_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t* const row64 = reinterpret_cast<uint64_t*>(qk_smem) + P.table_bytes / 8;
    uint32_t* const row32 = reinterpret_cast<uint32_t*>(qk_smem) + (P.table_bytes + P.n64 * nt * 8) / 4;
    if (tid == 0) {
      qk_mbar_expect_tx(&stage_bar, P.table_bytes + (P.n64 * 8u + P.n32 * 4u) * nt);
      qk_bulk_load(smem, P.tables, P.table_bytes, &stage_bar);
      if (P.use_tma) {
        /* the block's rows of every slot as ONE 2-D tile per plane (UTMALDG): two instructions instead of one bulk copy
         * per slot row, each of which the compiler has to issue lane by lane (its operands are uniform registers) --
         * a third of the instructions of a one-event launch */
        qk_tma_load_2d(row64, &P.tm64, (uint32_t)block_rep0, 0u, &stage_bar);
        if (P.n32) qk_tma_load_2d(row32, &P.tm32, (uint32_t)block_rep0, 0u, &stage_bar);
      }
    }
    for (uint32_t s = P.use_tma ? P.n64 + P.n32 : tid; s < P.n64 + P.n32; s += nt) {
      if (s < P.n64)
        qk_bulk_load(row64 + (size_t)s * nt, P.g64 + (size_t)s * P.rep_pad + block_rep0, nt * 8u, &stage_bar);
      else
        qk_bulk_load(row32 + (size_t)(s - P.n64) * nt, P.g32 + (size_t)(s - P.n64) * P.rep_pad + block_rep0, nt * 4u,
                     &stage_bar);
    }
    qk_mbar_wait(&stage_bar, 0);
  }
#endif

  Ctx cx;
  cx.hdr = reinterpret_cast<const DevHeader*>(smem);
  cx.comps = reinterpret_cast<const DevComp*>(smem + cx.hdr->off_comps);
  cx.dists = reinterpret_cast<const DevDist*>(smem + cx.hdr->off_dists);
  cx.dms = reinterpret_cast<const DevDelayMode*>(smem + cx.hdr->off_dms);
  cx.subs = reinterpret_cast<const uint16_t*>(smem + cx.hdr->off_subs);
  cx.sched = reinterpret_cast<const uint16_t*>(smem + cx.hdr->off_sched);
  cx.ext = reinterpret_cast<const DevExt*>(smem + cx.hdr->off_ext);
  cx.vecs = reinterpret_cast<const DevVec*>(smem + cx.hdr->off_vecs);
  cx.hot = reinterpret_cast<const DevHot*>(smem + cx.hdr->off_hot);
  cx.subq = reinterpret_cast<const DevSubQ*>(smem + cx.hdr->off_subq);
  cx.ccap = reinterpret_cast<const double*>(smem + cx.hdr->off_ccap);
  cx.pf32 = cx.hdr->pf32;
  cx.n_pf_words = cx.hdr->n_pf_words;
  const uint64_t rep_local = (uint64_t)(P.tile0 + QK_BID) * nt + tid;
  if constexpr (SM == 0) {
    cx.g64 = P.g64 + rep_local;
    cx.g32 = P.g32 + rep_local;
    cx.stride = (uint32_t)P.rep_pad;
    cx.stride32 = cx.stride;
    cx.base64 = cx.base32 = 0;
  } else {
    cx.g64 = nullptr;
    cx.g32 = nullptr;
    cx.stride = nt;
    cx.stride32 = nt;
    cx.base64 = P.table_bytes / 8 + tid;
    cx.base32 = (P.table_bytes + P.n64 * nt * 8) / 4 + tid;
  }
  const bool live = rep_local < P.n_reps;
  const uint64_t rep_global = P.rep_offset + rep_local;
  cx.rep_local = rep_local;
  cx.gs64 = P.gs64 + rep_local;
  cx.gs32 = P.gs32 + rep_local;
  if constexpr (qk_is_split(SM)) { /* cold planes [replication][slot] */
    cx.gc64 = P.gc64 + rep_local * P.n64c;
    cx.gc32 = P.gc32 + rep_local * P.n32c;
  } else {
    cx.gc64 = P.gc64 + rep_local;
    cx.gc32 = P.gc32 + rep_local;
  }
  cx.rep_pad = (uint32_t)P.rep_pad;
  cx.key0 = (uint32_t)P.seed;
  cx.key1 = (uint32_t)(P.seed >> 32);
  cx.rep_lo = (uint32_t)rep_global;
  cx.rep_hi = (uint32_t)(rep_global >> 32);
  cx.tapes = P.tapes;
  cx.tape_len = P.tape_len;
  cx.log_mask = P.log_cap ? P.log_cap - 1 : 0;
  cx.log_base = LOG ? P.log + 2 * (size_t)rep_local * P.log_cap : nullptr;
  cx.status = 0;
  cx.npend = 0;
  cx.cfire0 = cx.hdr->cfire0;
  cx.busy32 = cx.hdr->busy32;
  cx.emask32 = cx.hdr->emask32;
  cx.gshift = cx.hdr->gshift;
  cx.n_groups = cx.hdr->n_groups;
  cx.gmin64 = cx.hdr->gmin64;
  cx.gidx32 = cx.hdr->gidx32;
  cx.gdirty32 = cx.hdr->gdirty32;
  cx.kbest = QK_TIME_NONE;
  cx.kidx = 0;
  cx.kvalid = false; /* two-tier queue: the first pop of a launch scans the cold fire array */
  const uint32_t n_comp = cx.hdr->n_comp, n_sched = cx.hdr->n_sched, n_ext = cx.hdr->n_ext;
  const uint32_t n_stale_words = cx.hdr->n_stale_words;
  const uint32_t n_sched_pad = (n_sched + 3u) & ~3u;
  /* entries of the hot fire array: every schedulable component, or (two-tier queue) the stocks */
  const uint32_t n_hot_pad = qk_is_two_tier(SM) ? ((cx.hdr->n_hot + 3u) & ~3u) : n_sched_pad;
  const uint16_t* const hot_rank =
      reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(cx.hdr) + cx.hdr->off_hot_rank);
  (void)hot_rank;

  if (P.do_init) {
    /* fresh replication: zero state, no pending events, stocks hold their initial items */
    for (uint32_t s = 0; s < P.n64; ++s) S64(s) = 0;
    for (uint32_t s = 0; s < P.n32; ++s) S32(s) = 0;
    for (uint32_t i = 0; i < n_hot_pad; ++i) S64(G64_FIRE0 + i) = QK_TIME_NONE; /* incl. the padding slots */
    for (uint32_t s = 0; s < P.n64c; ++s) C64(s) = 0;
    for (uint32_t s = 0; s < P.n32c; ++s) C32(s) = 0;
    if constexpr (qk_is_two_tier(SM)) {
      for (uint32_t i = 0; i < n_sched_pad; ++i) QK_CFIRE(i) = QK_TIME_NONE;
      for (uint32_t g = 0; g < cx.n_groups; ++g) S64(cx.gmin64 + g) = QK_TIME_NONE;
    }
    cx.now = P.t0;
    cx.log_cnt = 0;
    for (uint32_t s = 0; s < n_comp * ST_PER_COMP; ++s) {
      cx.gs64[(size_t)s * P.rep_pad] = 0;
      cx.gs32[(size_t)s * P.rep_pad] = 0;
    }
    if (live) {
      uint32_t initial_base = 0;
      for (uint32_t i = 0; i < n_comp; ++i) {
        const DevComp* c = &cx.comps[i];
        if (c->o64_ttnpe != 0xFFFF) W64(c->o64_ttnpe) = QK_TIME_NONE;
        if (c->o64_nextdur != 0xFFFF) {
          W64(c->o64_nextdur) = QK_DUR_EMPTY;
          S32(cx.pf32 + (c->pf_idx >> 5)) |= 1u << (c->pf_idx & 31);
        }
        /* DelayModes::modify(Add) samples until_delay when the mode is added (delays.rs:159-164) */
        for (int k = 0; k < c->n_dm && !cx.status; ++k)
          W64(c->o64_dm + k) = qk_sample_duration<SM>(cx, cx.dms[c->dm_off + k].until_delay);
        if (c->kind == QK_DISCRETE_STOCK) {
          /* with_initial_resource: the count rides in DevComp.dist_time (unused for stocks) */
          const uint32_t n_init = c->dist_time;
          for (uint32_t k = 0; k < n_init; ++k) C32(c->ocold32 + k) = (63u << 26) | (initial_base + k);
          if (FT == 2 && (c->flags & DC_CONTAINER)) { /* with_initial_resource(ItemDeque<Vector3Container>) */
            const uint64_t* cinit = reinterpret_cast<const uint64_t*>(smem + cx.hdr->off_cinit) + 3u * initial_base;
            for (uint32_t k = 0; k < 3u * n_init; ++k) C64(c->ocold64 + k) = cinit[k];
          }
          initial_base += n_init;
          S32(c->o32 + S32_HC) = QK_HC_PACK(0u, n_init, qk_stock_cat(n_init, c->low, c->max));
          if (LOG && n_init) S32(c->o32 + S32_HEADITEM) = (63u << 26) | (initial_base - n_init);
          cx.gs32[(size_t)(i * ST_PER_COMP + ST32_MAXOCC) * P.rep_pad] = n_init;
          cx.gs64[(size_t)(i * ST_PER_COMP + ST64_TIME) * P.rep_pad] = 0ull - (uint64_t)n_init * P.t0;
        } else if (c->kind == QK_ENVIRONMENT) {
          W32(c->o32 + E32_STATE) = c->low; /* initial state rides in `low` */
        } else if (c->kind == QK_VECTOR_STOCK) {
          const DevVec* v = &cx.vecs[c->vec_idx];
          qk_vstore<SM, qk_vd(FT)>(cx, v->o64_res, v->dim, v->vinit);
          W64(v->o64_low) = QK_D2U(v->vlow);
          W64(v->o64_last) = P.t0;
        } else if (FT != 0 && c->kind == QK_VECTOR_PROCESS && cx.vecs[c->vec_idx].o32_qty != 0xFFFF) {
          W32(cx.vecs[c->vec_idx].o32_qty) = cx.vecs[c->vec_idx].dist_qty;
        }
      }
    }
    S64(G64_NEVENTS) = 0;
  } else {
    if constexpr (SM != 0) {
      /* stage the block's hot planes.  Slot row s of the block -- nt consecutive replications -- is contiguous both in
       * the HBM plane ([slot][replication]) and in shared memory (slot * nt + tid), so ONE bulk asynchronous copy per row
       * (cp.async.bulk: the TMA engine moves 256-512 bytes per instruction, no register round trip) brings it in, all
       * rows in flight at once, completing on an mbarrier.  With few events per launch this load IS the kernel: as
       * dependent LDG/STS pairs it ran at 44 % of the HBM roofline, as one 4/8-byte cp.async (LDGSTS) per thread and slot
       * at 50-55 % (profiles/README.md). */
#if !defined(QK_HOST_EMUL) && !defined(QK_TABLES_BY_LDG)
      /* (done at the top of the kernel, together with the tables) */
#elif !defined(QK_HOST_EMUL)
      __shared__ __align__(8) uint64_t stage_bar;
      const uint64_t block_rep0 = (uint64_t)(P.tile0 + QK_BID) * nt;
      if (tid == 0) {
        qk_mbar_init(&stage_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncthreads();
      if (tid == 0) qk_mbar_expect_tx(&stage_bar, (P.n64 * 8u + P.n32 * 4u) * nt);
      {
        uint64_t* const row64 = reinterpret_cast<uint64_t*>(qk_smem) + P.table_bytes / 8;
        uint32_t* const row32 = reinterpret_cast<uint32_t*>(qk_smem) + (P.table_bytes + P.n64 * nt * 8) / 4;
        for (uint32_t s = tid; s < P.n64 + P.n32; s += nt) {
          if (s < P.n64)
            qk_bulk_load(row64 + (size_t)s * nt, P.g64 + (size_t)s * P.rep_pad + block_rep0, nt * 8u, &stage_bar);
          else
            qk_bulk_load(row32 + (size_t)(s - P.n64) * nt, P.g32 + (size_t)(s - P.n64) * P.rep_pad + block_rep0, nt * 4u,
                         &stage_bar);
        }
      }
      qk_mbar_wait(&stage_bar, 0);
#else
      for (uint32_t s = 0; s < P.n64; ++s) qk_stage8(&S64(s), &P.g64[(size_t)s * P.rep_pad + rep_local]);
      for (uint32_t s = 0; s < P.n32; ++s) qk_stage4(&S32(s), &P.g32[(size_t)s * P.rep_pad + rep_local]);
      qk_stage_wait();
#endif
    }
    cx.now = S64(G64_NOW);
    cx.status = S32(G32_STATUS);
    cx.log_cnt = S32(G32_LOGCNT);
  }

  for (uint32_t w = 0; w < cx.n_pf_words; ++w) cx.npend += (uint32_t)QK_POPC(S32(cx.pf32 + w));

  /* ---- event loop ----------------------------------------------------------------------------------------
   * A lane's work is a stream of update_state(p) calls: an action popped from its scheduler queue expands into a
   * call list ([p] for a keyed event, the subscriber list for an emit_change / environment change, the init list
   * at t0).  Each trip of the loop has three warp-converged phases:
   *   SCAN    every lane walks forward through its stream -- popping the next action when the list is empty and
   *           disposing of the calls that only refresh a countdown (qk_try_quick) -- until it stands on a call that
   *           needs the full handler (at most QK_SCAN_ITERS pops per trip, so no lane holds the warp for long);
   *   REFILL  deferred variate prefetch (see below);
   *   UPDATE  one full update_state per lane that has one. */
  /* events left in this launch, as a 32-bit count-down (registers are what limits residency): the host cuts finite
   * budgets into launches of at most 2^32 - 2 events; ~0 means no limit, and then the counter is re-armed when it
   * runs out (4.29e9 events of one replication in one launch) */
  /* (which counter is cheaper is a matter of register allocation, measured per kernel family on the box: the logging
   * kernels gain 1.5 % from the 32-bit count-down, the statistics-only kernels lose 2 % with it) */
  constexpr bool EV32 = LOG;
  uint64_t events_done = 0;
  const uint64_t budget = live ? P.event_budget : 0;
  const bool unlimited = P.event_budget == ~0ull;
  const uint32_t ev_initial = live ? (unlimited ? 0xFFFFFFFFu : (uint32_t)P.event_budget) : 0u;
  uint32_t ev_left = ev_initial;
  const uint64_t t_until = P.t_until;
  /* Model::init of every component in registration order is the first call list of the loop */
  uint32_t rem = (P.do_init && live) ? cx.hdr->n_init : 0, ptr = cx.hdr->init_off;
  const uint32_t ident_off = cx.hdr->ident_off;
  uint64_t src = SRC_INIT;
  uint32_t ext_cursor = FT != 0 ? S32(G32_EXT_CURSOR) : 0u; /* the lean kernel has no external events */
  bool done = cx.status != 0; /* a faulted replication stays frozen; later faults arrive as return values */
  if (done) rem = 0;
  bool keyed = false; /* the pending call is a component's own keyed event: its fire slot was just cleared, so the
                         quick path's "fire == NONE means idle" does not apply -- always the full handler */
  for (;;) {
    QK_SYNCWARP();
#pragma unroll 1
    for (int it = 0; it < QK_SCAN_ITERS; ++it) {
      /* dispose of the quick calls at the head of the list */
      while (rem != 0 && !keyed && qk_try_quick<LOG, SM, FT>(cx, ptr, src)) {
        ptr += 1;
        rem -= 1;
      }
      if (rem == 0 && !done) {
        if constexpr (EV32) {
          if (ev_left == 0 && unlimited && live) { /* re-arm (see above) */
            S64(G64_NEVENTS) += 0xFFFFFFFFull;
            ev_left = 0xFFFFFFFFu;
          }
          done = ev_left == 0;
        } else {
          done = events_done >= budget;
        }
      }
      if (!QK_ANY_SYNC(rem == 0 && !done)) break; /* every lane stands on a full call or has finished */
      if (rem == 0 && !done) {
        /* pop the scheduler queue: min over (time, rank); rank 0 = externally scheduled events */
        uint64_t best = (FT != 0 && ext_cursor < n_ext) ? cx.ext[ext_cursor].t : QK_TIME_NONE;
        int bi = -1;
        /* fire[] is padded with NONE to a multiple of four slots (lowering): groups of four independent loads and a
         * two-level tournament instead of one long dependent chain of 64-bit compares.  Strict '<' everywhere keeps
         * the LOWEST index among equal times, i.e. registration-rank order. */
        if constexpr (qk_is_two_tier(SM)) {
          /* two-tier queue.  Keyed events: the cached minimum over the cold array, re-established by a scan (independent,
           * coalesced global loads) only after the minimum itself was popped or dropped. */
          if (!cx.kvalid) {
            /* groups whose minimum fired or was dropped: fetch that group of the cold array (2^gshift consecutive ranks,
             * one contiguous run of the replication's record) and re-establish its minimum */
            uint32_t dm = S32(cx.gdirty32);
            while (dm) {
              const uint32_t g = (uint32_t)__ffs((int)dm) - 1u;
              dm &= dm - 1;
              const uint32_t lo = g << cx.gshift;
              uint32_t hi = lo + (1u << cx.gshift);
              if (hi > n_sched_pad) hi = n_sched_pad;
              uint64_t gb = QK_TIME_NONE;
              uint32_t gi = 0;
              for (uint32_t i = lo; i < hi; i += 4) {
                const uint64_t t0 = QK_CFIRE(i), t1 = QK_CFIRE(i + 1);
                const uint64_t t2 = QK_CFIRE(i + 2), t3 = QK_CFIRE(i + 3);
                const bool p01 = t1 < t0, p23 = t3 < t2;
                const uint64_t a = p01 ? t1 : t0, b = p23 ? t3 : t2;
                const uint32_t ai = p01 ? i + 1 : i, bj = p23 ? i + 3 : i + 2;
                const bool pab = b < a;
                const uint64_t m = pab ? b : a;
                if (m < gb) {
                  gb = m;
                  gi = pab ? bj : ai;
                }
              }
              S64(cx.gmin64 + g) = gb;
              S32(cx.gidx32 + g) = gi;
            }
            S32(cx.gdirty32) = 0;
            /* the overall minimum: the group minima, ascending = rank order (strict '<' keeps the lowest rank) */
            uint64_t kb = QK_TIME_NONE;
            uint32_t ki = 0;
            for (uint32_t g = 0; g < cx.n_groups; ++g) {
              const uint64_t t = S64(cx.gmin64 + g);
              if (t < kb) {
                kb = t;
                ki = S32(cx.gidx32 + g);
              }
            }
            cx.kbest = kb;
            cx.kidx = ki;
            cx.kvalid = true;
          }
          /* the stocks' pending emit_change events: the hot array, in rank order */
          /* (one or two are pending at a time: the pending mask names them, ascending = rank order) */
          uint64_t sb = QK_TIME_NONE;
          uint32_t sj = 0;
          for (uint32_t w = 0; w * 32u < n_hot_pad; ++w) {
            uint32_t em = S32(cx.emask32 + w);
            while (em) {
              const uint32_t j = w * 32u + (uint32_t)__ffs((int)em) - 1u;
              em &= em - 1;
              const uint64_t t = S64(G64_FIRE0 + j);
              if (t < sb) {
                sb = t;
                sj = j;
              }
            }
          }
          /* min over (time, rank) of the two tiers; an external event (rank 0) keeps winning ties */
          uint64_t m = cx.kbest;
          uint32_t mi = cx.kidx;
          if (sb != QK_TIME_NONE) {
            const uint32_t sr = hot_rank[sj];
            if (sb < m || (sb == m && sr < mi)) {
              m = sb;
              mi = sr;
            }
          }
          if (m < best) {
            best = m;
            bi = (int)mi;
          }
        } else {
        for (uint32_t i = 0; i < n_sched_pad; i += 4) {
          const uint64_t t0 = S64(G64_FIRE0 + i), t1 = S64(G64_FIRE0 + i + 1);
          const uint64_t t2 = S64(G64_FIRE0 + i + 2), t3 = S64(G64_FIRE0 + i + 3);
          const bool p01 = t1 < t0, p23 = t3 < t2;
          const uint64_t a = p01 ? t1 : t0, b = p23 ? t3 : t2;
          const uint32_t ai = p01 ? i + 1 : i, bj = p23 ? i + 3 : i + 2;
          const bool pab = b < a;
          const uint64_t m = pab ? b : a;
          if (m < best) {
            best = m;
            bi = (int)(pab ? bj : ai);
          }
        }
        }
        bool stale = false;
        uint32_t stale_word = 0, stale_mask = 0;
        for (uint32_t w = 0; w < n_stale_words; ++w) {
          const uint32_t m = S32(G32_STALE0 + w);
          if (m) {
            const int is = (int)(w * 32) + __ffs((int)m) - 1;
            if (best > cx.now || is < bi) { /* a stale handle fires at `now`, in rank order */
              stale = true;
              bi = is;
              best = cx.now;
              stale_word = w;
              stale_mask = m & (m - 1);
            }
            break;
          }
        }
        if (best == QK_TIME_NONE || best > t_until) {
          done = true; /* queue empty, or next action is beyond the step_until horizon */
        } else {
          if (stale) S32(G32_STALE0 + stale_word) = stale_mask;
          cx.now = best;
          if constexpr (EV32) ev_left -= 1; else events_done += 1;
          if (FT != 0 && bi < 0) {
            /* ScheduledEventConfig::SetEnvironmentState -> BasicEnvironment::set_state (core.rs:256-265, 1393-1396) */
            const DevExt* e = &cx.ext[ext_cursor];
            ext_cursor += 1;
            const uint32_t target = e->target;
            const DevHotA a = QK_HOT_A(target);
            if (e->type == 1) {
              /* SetLowCapacity on an F64Stock: with_low_capacity_inplace, then add((0., SCH_000000)) (core.rs:1382-1386) */
              W64(cx.vecs[cx.comps[target].vec_idx].o64_low) = QK_D2U(e->value);
              double zero[qk_vd(FT)];
#pragma unroll
              for (int k = 0; k < qk_vd(FT); ++k) zero[k] = 0.0;
              qk_vstock_add<LOG, SM, qk_vd(FT)>(cx, target, zero, SRC_SCH);
              if (cx.status) done = true;
            } else if (e->type == 2) {
              /* SetProcessQuantity on an F64Process: with_process_quantity_distr_inplace(instance), then
               * update_state(SCH_000000) (core.rs:1389-1390) */
              W32(cx.vecs[cx.comps[target].vec_idx].o32_qty) = e->state;
              src = SRC_SCH;
              rem = 1;
              ptr = ident_off + target;
            } else if (W32(a.o32 + E32_STATE) != e->state) {
              const uint32_t st = e->state;
              W32(a.o32 + E32_STATE) = st;
              src = qk_log<LOG, SM>(cx, a, target, SRC_SCH, QK_LOG_ENV_STATE, st, 0);
              rem = a.n_subs;
              ptr = QK_HOT_SB(target).subs_off;
            }
          } else {
            const uint32_t comp = cx.sched[bi];
            const DevHotA a = QK_HOT_A(comp);
            if (a.kind == QK_DISCRETE_STOCK) {
              /* emit_change (discrete.rs:227-233): log the state AT EMISSION TIME, then update_state on every
               * subscriber in connection order */
              const DevHotSB sb = QK_HOT_SB(comp);
              const uint32_t src_seq = qk_emit_pop<LOG, SM, true>(cx, a, sb.emit_depth, a.olog64 + sb.ring_cap);
              if (LOG) {
                const uint32_t cnt = QK_HC_CNT(S32(a.o32 + S32_HC));
                const uint32_t empty = sb.max > cnt ? sb.max - cnt : 0;
                src = qk_log<LOG, SM>(cx, a, comp, ((uint64_t)comp << 32) | src_seq, QK_LOG_STOCK_STATE_CHANGE,
                                      qk_stock_cat(cnt, sb.low, sb.max), ((uint64_t)cnt << 32) | empty);
              }
              rem = a.n_subs;
              ptr = sb.subs_off;
            } else if (FT != 0 && a.kind == QK_VECTOR_STOCK) {
              /* VectorStock::emit_change (vector.rs:167-171) logs the state carried by the event */
              const DevHotSB sb = QK_HOT_SB(comp);
              const DevComp* c = &cx.comps[comp];
              uint32_t packed = 0;
              uint64_t occ_bits = 0;
              if (LOG) {
                const uint32_t head = (S32(a.o32 + S32_EMIT) >> 16) & 0xFFu;
                occ_bits = C64(a.olog64 + head);
              }
              packed = qk_emit_pop<LOG, SM>(cx, a, sb.emit_depth, c->ocold32);
              src = qk_log<LOG, SM>(cx, a, comp, ((uint64_t)comp << 32) | (packed & 0x3FFFFFFFu),
                                    QK_LOG_VSTOCK_STATE_CHANGE, packed >> 30, occ_bits);
              qk_log_vec<LOG, SM>(cx, c, comp, src, 0, __dsub_rn(cx.vecs[c->vec_idx].vmax, QK_U2D(occ_bits)), 0.0, 0.0);
              rem = a.n_subs;
              ptr = sb.subs_off;
            } else {
              /* keyed update_state(source_event_id) (discrete.rs:549,556) */
              if (stale) {
                if (LOG) src = C64(a.olog64 + PL64_STALE_SRC);
              } else {
                if (LOG)
                  src = FT == 0 ? (((uint64_t)comp << 32) | (uint32_t)(S32(a.olog32 + L32_SEQ) - 1u))
                                : C64(a.olog64 + PL64_SCHED_SRC);
                qk_fire_clear<SM>(cx, (uint32_t)bi);
                keyed = true;
              }
              rem = 1;
              ptr = ident_off + comp;
            }
          }
        } /* an action was popped */
      }
      QK_SYNCWARP();
    }
    QK_SYNCWARP();
    if (!QK_ANY_SYNC(!done || rem != 0)) break; /* uniform exit: every lane of the warp has finished */
    /* REFILL phase (deferred variate prefetch): the fp64 sampler is the most expensive code on the path, and only
     * the few lanes that start a process need a variate on a given trip.  So starts consume a pre-drawn duration,
     * and the sampler runs as its own phase -- for all lanes with consumed slots at once -- when enough lanes are
     * waiting, or when some lane's next call could start a process whose slot is still empty. */
    QkPre pre;
    pre.has = false;
    pre.has_up = false;
    pre.dur = 0;
    pre.idx = 0;
    if constexpr (qk_is_split(SM) || SM == 0) {
      if (rem != 0) qk_prefetch_record<SM>(cx, cx.subs[ptr], pre); /* the loads travel while the REFILL phase runs */
    }
    {
      const bool must = rem != 0 && cx.npend != 0 && qk_needs_refill_now<SM>(cx, cx.subs[ptr]);
      const uint32_t waiting = QK_BALLOT(cx.npend != 0);
      if (QK_ANY_SYNC(must) || QK_POPC(waiting) >= QK_REFILL_LANES) {
        if (cx.npend != 0) qk_refill<SM>(cx);
      }
    }
    QK_SYNCWARP();
    if (rem != 0) {
      const uint32_t p = cx.subs[ptr];
      ptr += 1;
      rem -= 1;
      const bool fault = qk_update_state<LOG, SM, FT>(cx, p, src, &pre);
      keyed = false;
      if (fault) {
        rem = 0;
        done = true;
      }
    }
  }
  /* step_until(t) leaves the clock at t */
  if (!cx.status && live && P.t_until != QK_TIME_NONE && P.t_until > cx.now) cx.now = P.t_until;

  if constexpr (LOG && SM != 0) qk_async_wait(); /* head-item copies still in flight land before the write-back */
  S64(G64_NOW) = cx.now;
  S64(G64_NEVENTS) += EV32 ? (uint64_t)(ev_initial - ev_left) : events_done;
  S32(G32_STATUS) = cx.status;
  if (FT != 0) S32(G32_EXT_CURSOR) = ext_cursor;
  S32(G32_LOGCNT) = cx.log_cnt;
  if constexpr (SM != 0) {
#ifndef QK_HOST_EMUL
    /* write-back: one bulk asynchronous store per slot row (shared -> global) */
    qk_fence_async_proxy();
    __syncthreads();
    {
      const uint64_t block_rep0 = (uint64_t)(P.tile0 + QK_BID) * nt;
      uint64_t* const row64 = reinterpret_cast<uint64_t*>(qk_smem) + P.table_bytes / 8;
      uint32_t* const row32 = reinterpret_cast<uint32_t*>(qk_smem) + (P.table_bytes + P.n64 * nt * 8) / 4;
      bool stored = false;
      if (P.use_tma && tid == 0) { /* one tile per plane (UTMASTG) */
        qk_tma_store_2d(&P.tm64, (uint32_t)block_rep0, 0u, row64);
        if (P.n32) qk_tma_store_2d(&P.tm32, (uint32_t)block_rep0, 0u, row32);
        stored = true;
      }
      for (uint32_t s = P.use_tma ? P.n64 + P.n32 : tid; s < P.n64 + P.n32; s += nt) {
        if (s < P.n64)
          qk_bulk_store(P.g64 + (size_t)s * P.rep_pad + block_rep0, row64 + (size_t)s * nt, nt * 8u);
        else
          qk_bulk_store(P.g32 + (size_t)(s - P.n64) * P.rep_pad + block_rep0, row32 + (size_t)(s - P.n64) * nt, nt * 4u);
        stored = true;
      }
      if (stored) qk_bulk_store_wait();
    }
#else
#pragma unroll 4
    for (uint32_t s = 0; s < P.n64; ++s) P.g64[(size_t)s * P.rep_pad + rep_local] = S64(s);
#pragma unroll 4
    for (uint32_t s = 0; s < P.n32; ++s) P.g32[(size_t)s * P.rep_pad + rep_local] = S32(s);
#endif
  }
}

/* what is left of a countdown at the CLOCK: state keeps (left, previous_check_time) as of the component's last
 * update_state; busy time is reported up to `now`, so the elapsed-but-unapplied part counts as consumed unless a
 * delay or the environment has paused the countdown */
__device__ __forceinline__ uint64_t qk_left_at_clock(uint32_t flags, uint64_t left, uint64_t dt) {
  if (((flags >> PF_FIX_SHIFT) & 0xFFu) != 0 || (flags & PF_ENV_STOPPED)) return left;
  return left - (dt < left ? dt : left);
}

/* per-replication (count, time_ns, aux, level) of component c, from the HBM state planes; w64 / w32 = the planes the
 * component records live in (the hot planes, or the cold ones under the split placement: DevHeader.split) */
__device__ __forceinline__ void qk_comp_stats_dev(const DevComp* c, const DevVec* vecs, uint32_t comp, const uint64_t* g64,
                                                  const uint32_t* g32, const uint64_t* w64, const uint32_t* w32,
                                                  uint32_t w_aos64, uint32_t w_aos32,
                                                  const uint64_t* gs64, const uint32_t* gs32,
                                                  uint64_t rep_pad, uint64_t rep, qk_comp_stats* o) {
  /* w_aos64 / w_aos32: 0 = the record planes are [slot][replication] ; else [replication][slot] with that many slots */
#define G64(slot) g64[(size_t)(slot)*rep_pad + rep]
#define G32(slot) g32[(size_t)(slot)*rep_pad + rep]
#define WG64(slot) w64[w_aos64 ? (size_t)rep * w_aos64 + (slot) : (size_t)(slot)*rep_pad + rep]
#define WG32(slot) w32[w_aos32 ? (size_t)rep * w_aos32 + (slot) : (size_t)(slot)*rep_pad + rep]
#define GS64(k) gs64[(size_t)(comp * ST_PER_COMP + (k)) * rep_pad + rep]
#define GS32(k) gs32[(size_t)(comp * ST_PER_COMP + (k)) * rep_pad + rep]
  const uint64_t now = G64(G64_NOW);
  o->fvalue = 0.0;
  o->faux = 0.0;
  if (c->kind == QK_VECTOR_STOCK) {
    const DevVec* v = &vecs[c->vec_idx];
    double r[QK_VEC_MAX_DIM];
    for (uint32_t k = 0; k < QK_VEC_MAX_DIM; ++k) r[k] = k < v->dim ? QK_U2D(WG64(v->o64_res + k)) : 0.0;
    const double total = qk_vtotaln<QK_VEC_MAX_DIM>(r, v->dim, v->tmask);
    o->count = GS32(ST32_COUNT);
    o->time_ns = 0;
    o->aux = 0;
    o->level = qk_vcat(total, QK_U2D(WG64(v->o64_low)), v->vmax);
    o->fvalue = total;
    o->faux = __dadd_rn(QK_U2D(GS64(ST64_TIME)), __dmul_rn(total, __dmul_rn((double)(now - WG64(v->o64_last)), 1e-9)));
  } else if (c->kind >= QK_VECTOR_PROCESS && c->kind <= QK_VECTOR_SPLITTER) {
    const DevVec* v = &vecs[c->vec_idx];
    const uint32_t flags = WG32(c->o32 + P32_FLAGS);
    const bool has = flags & PF_HAS_ITEM;
    double r[QK_VEC_MAX_DIM];
    for (uint32_t k = 0; k < QK_VEC_MAX_DIM; ++k) r[k] = k < v->dim ? QK_U2D(WG64(v->o64_res + k)) : 0.0;
    o->count = GS32(ST32_COUNT);
    o->time_ns = GS64(ST64_TIME) - (has ? qk_left_at_clock(flags, WG64(c->o64 + P64_LEFT), now - WG64(c->o64 + P64_PREV)) : 0);
    o->aux = GS64(ST64_DELAY);
    o->level = has ? 1 : 0;
    o->fvalue = QK_U2D(GS64(ST64_FQ));
    o->faux = !has ? 0.0 : (c->kind == QK_VECTOR_COMBINER ? QK_U2D(WG64(v->o64_res + v->dim)) : qk_vtotaln<QK_VEC_MAX_DIM>(r, v->dim, v->tmask));
  } else if (c->kind == QK_DISCRETE_STOCK) {
    const uint32_t cnt = QK_HC_CNT(G32(c->o32 + S32_HC));
    o->count = GS32(ST32_COUNT);
    o->time_ns = GS64(ST64_TIME) + (uint64_t)cnt * now;
    o->aux = GS32(ST32_MAXOCC);
    o->level = cnt;
  } else if (c->kind == QK_DISCRETE_PARALLEL_PROCESS || c->kind == QK_CONTAINER_LOAD_PROCESS ||
             c->kind == QK_CONTAINER_UNLOAD_PROCESS) {
    const uint32_t nprog = WG32(c->o32 + PP32_NPROG);
    const uint32_t pflags = WG32(c->o32 + PP32_FLAGS);
    const uint64_t pdt = now - WG64(c->o64 + PP64_PREV);
    uint64_t busy = GS64(ST64_TIME);
    for (uint32_t i = 0; i < nprog; ++i) busy -= qk_left_at_clock(pflags, WG64(c->o64_dm + c->n_dm + i), pdt);
    o->count = GS32(ST32_COUNT);
    o->time_ns = busy;
    o->aux = GS64(ST64_DELAY);
    o->level = nprog;
  } else if (c->kind == QK_ENVIRONMENT) {
    o->count = 0;
    o->time_ns = 0;
    o->aux = 0;
    o->level = WG32(c->o32 + E32_STATE);
  } else {
    const uint32_t flags = WG32(c->o32 + P32_FLAGS);
    const bool has = flags & PF_HAS_ITEM;
    o->count = GS32(ST32_COUNT);
    o->time_ns = GS64(ST64_TIME) - (has ? qk_left_at_clock(flags, WG64(c->o64 + P64_LEFT), now - WG64(c->o64 + P64_PREV)) : 0);
    o->aux = GS64(ST64_DELAY);
    o->level = has ? 1 : 0;
  }
#undef WG64
#undef WG32
#undef G64
#undef G32
#undef GS64
#undef GS32
}

#endif
