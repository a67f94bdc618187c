/*
 * qk_group.cuh -- the per-event state-update kernel for LARGE models: a lane group per replication (sm_100a).
 *
 * The thread-per-replication kernel (qk_kernels.cuh) keeps a replication's state in shared memory only while 12 or more
 * warps of replications fit per SM; a 65-90 component model needs 4-4.5 KB per replication, and on the HBM planes
 * every scheduler pop was ~90 dependent global loads and every emit_change broadcast a serial walk over 9-16
 * subscribers (84 % of the stall samples waiting for memory, profiles/README.md section 7).  Here G = 4 / 8 / 16 lanes
 * own one replication, so that the ~45 replications whose state fits the 227 KB of an SM are advanced by 6-23 warps:
 *
 *   - scheduler queue (NeXosim's queue behind Simulation::step; call sites discrete.rs:194,219,549,556): the fire slots
 *     live one-per-lane-strided in shared memory; a pop is n_sched / G independent loads per lane and a log2(G)-step
 *     shuffle reduction on (time, rank) -- registration-rank tie-break as in the thread-per-replication kernel;
 *   - emit_change broadcast (discrete.rs:227-233): the subscribers of the list are evaluated G at a time, one lane
 *     each (busy? would it start?) with qk_quick_eval; a ballot finds the first one that needs the full handler, the
 *     ones before it are disposed of at once (their ProcessNonStart records written by their lanes, in order);
 *   - variates: consumed pre-drawn durations are re-drawn G at a time (REFILL);
 *   - the full update_state handler (the same source as the other kernel) runs on the group's lane 0.
 *
 * State placement: shared memory, slot-major like the HBM planes -- row `slot` holds that slot of the block's R
 * replications -- so a block stages its state with ONE bulk asynchronous copy (cp.async.bulk, the TMA engine) per
 * slot row, global -> shared, completing on an mbarrier, and stores it back the same way.  The row stride of the 64-bit
 * rows is chosen = 2 (mod 4) elements: the G lanes of a group read G consecutive slots of one replication, i.e.
 * addresses one row stride apart, and with that stride the lanes of a half-warp hit distinct bank pairs.
 */
#ifndef QK_GROUP_CUH
#define QK_GROUP_CUH
#include "qk_kernels.cuh"

#ifndef QK_HOST_EMUL
#define QK_WARP 32u
#define QK_LANE (threadIdx.x & 31u)
#define QK_SHFL32(v, src) __shfl_sync(0xFFFFFFFFu, (v), (int)(src))
#define QK_SHFL_XOR32(v, m) __shfl_xor_sync(0xFFFFFFFFu, (v), (int)(m))
#define QK_REDUCE_MIN32(mask, v) __reduce_min_sync((mask), (v)) /* REDUX over the lanes of `mask` (the caller's group) */
#define QK_SYNCBLOCK() __syncthreads()

#endif

__device__ __forceinline__ uint64_t qk_shfl64(uint64_t v, uint32_t src) {
  const uint32_t lo = QK_SHFL32((uint32_t)v, src), hi = QK_SHFL32((uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t qk_shfl_xor64(uint64_t v, uint32_t m) {
  const uint32_t lo = QK_SHFL_XOR32((uint32_t)v, m), hi = QK_SHFL_XOR32((uint32_t)(v >> 32), m);
  return ((uint64_t)hi << 32) | lo;
}

/* What would update_state(subs[ptr]) do right now?  The side-effect-free half of qk_try_quick:
 *   0              it finishes or starts something (or has no quick path): the full handler
 *   1              busy and not due: nothing
 *   2              a busy component with delay modes: its countdowns lose *dt (qk_dm_apply)
 *   0x100 | reason idle and it stays blocked: only its ProcessNonStart record */
template <bool LOG, int SM, int FT>
__device__ __forceinline__ uint32_t qk_quick_eval(Ctx& cx, uint32_t ptr, uint64_t now, uint64_t* dt) {
  union {
    uint2 v;
    DevSubQ q;
  } u;
  u.v = *reinterpret_cast<const uint2*>(&cx.subq[ptr]);
  const DevSubQ q = u.q;
  if (LOG && (q.flags & SQ_DUP)) return 0u; /* the component is in this list twice: its records must be written in turn */
  if (FT != 0 && (q.flags & SQ_DM)) {
    uint32_t reason = 0;
    const uint32_t code = qk_dm_quick<SM>(cx, q, cx.subs[ptr], now, &reason, dt);
    return code == 1u ? (0x100u | reason) : code;
  }
  if (!(q.flags & SQ_QUICK)) return 0u;
  const uint64_t fire = S64(q.fire_slot);
  if (fire != QK_TIME_NONE) return fire > now ? 1u : 0u;
  const uint32_t us = (q.flags & SQ_US_FORCED) ? (q.flags >> SQ_US_SHIFT) & 3u : QK_HC_CAT(S32(q.up_hc));
  const uint32_t ds = (q.flags & SQ_DS_FORCED) ? (q.flags >> SQ_DS_SHIFT) & 3u : QK_HC_CAT(S32(q.down_hc));
  if ((us == 1u || us == 2u) && (ds == 0u || ds == 1u)) return 0u; /* starts */
  const uint32_t reason = us == 0u   ? QK_REASON_UPSTREAM_EMPTY
                          : us == 3u ? QK_REASON_UPSTREAM_NOT_CONNECTED
                          : ds == 2u ? QK_REASON_DOWNSTREAM_FULL
                                     : QK_REASON_DOWNSTREAM_NOT_CONNECTED;
  return 0x100u | reason;
}

/* the (time, rank) minimum over a group's lanes: xor-butterfly shuffles (default) or three REDUX reductions over the
 * group's lane mask.  Measured (A/B, one box, 65,536 x 2,000 events, statistics only): shuffles 1.83e9 / 2.03e9 events/s
 * (haulage / serial line), REDUX 1.58e9 / 1.72e9 -- a reduction over a partial member mask is not the fast path */
#ifndef QK_GROUP_REDUX
#define QK_GROUP_REDUX 0
#endif
#ifndef QK_GROUP_REFILL_EAGER
#define QK_GROUP_REFILL_EAGER 0
#endif

/* GROUP-mode state accessors are the run-time-stride ones of qk_kernels.cuh (SM < 0): slot * row stride + replication */
#define QK_SMG (-1)

/* ---- SCAN: what every trip of a group's event loop begins with -- up to QK_SCAN_ITERS rounds of POP (groups whose call
 * list is empty take the next action from their scheduler queue, the group's lanes scanning the fire slots side by side)
 * and QUICK (the calls at the head of the list evaluated G at a time, one lane each), until every group of the warp stands
 * on a call that needs the full handler or has finished.  The loop state (rem, ptr, src, stand, done, event counters,
 * ext_cursor) is the group leader's; the other lanes carry rem = 0, done = true.  Warp-convergent: every lane of the warp
 * executes every collective. */
template <bool LOG, int G, int FT>
__device__ __forceinline__ void qk_scan_phase(Ctx& cx, uint32_t& rem, uint32_t& ptr, uint64_t& src, bool& stand, bool& done,
                                              uint32_t& ev_left, uint64_t& events_done, uint32_t& ext_cursor,
                                              const bool leader, const uint32_t lead, const uint32_t j,
                                              const uint32_t gshift, const uint32_t gmask, const bool live,
                                              const bool unlimited, const uint64_t budget, const uint64_t t_until,
                                              const uint32_t n_sched, const uint32_t n_ext, const uint32_t n_stale_words,
                                              const uint32_t ident_off) {
  constexpr int SM = QK_SMG;
  constexpr bool EV32 = LOG;
#pragma unroll 1
  for (int it = 0; it < QK_SCAN_ITERS; ++it) {
    /* -- POP: groups whose list is empty take the next action from their scheduler queue */
    if (leader && rem == 0 && !done) {
      if constexpr (EV32) {
        if (ev_left == 0 && unlimited && live) { /* re-arm (qk_kernels.cuh) */
          S64(G64_NEVENTS) += 0xFFFFFFFFull;
          ev_left = 0xFFFFFFFFu;
        }
        done = ev_left == 0;
      } else {
        done = events_done >= budget;
      }
    }
    const bool want = QK_SHFL32((leader && rem == 0 && !done) ? 1u : 0u, lead) != 0;
    const bool any_want = QK_ANY_SYNC(want);
    if (!any_want && !QK_ANY_SYNC(leader && rem != 0 && !stand)) break; /* every group stands on a full call or has finished */
    if (any_want) {
    uint64_t best = QK_TIME_NONE;
    uint32_t bidx = 0xFFFFFFFFu;
    if (want) {
      /* lane j owns fire slots j, j + G, ...: independent loads, ascending, strict '<' keeps the lowest index */
#pragma unroll 4
      for (uint32_t i = j; i < n_sched; i += G) {
        const uint64_t t = S64(G64_FIRE0 + i);
        if (t < best) {
          best = t;
          bidx = i;
        }
      }
    }
#if QK_GROUP_REDUX
    {
      /* min over (time, registration rank) of the group's lanes: three warp reductions (REDUX) on the 32-bit halves
       * and the index, each over this group's lanes only -- the groups of a warp reduce side by side */
      const uint32_t wm = gmask << gshift;
      const uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
      const uint32_t mhi = QK_REDUCE_MIN32(wm, hi);
      const uint32_t mlo = QK_REDUCE_MIN32(wm, hi == mhi ? lo : 0xFFFFFFFFu);
      const uint32_t mid = QK_REDUCE_MIN32(wm, (hi == mhi && lo == mlo) ? bidx : 0xFFFFFFFFu);
      best = ((uint64_t)mhi << 32) | mlo;
      bidx = mid;
    }
#else
#pragma unroll
    for (uint32_t m = G / 2; m > 0; m >>= 1) { /* min over (time, registration rank): xor butterfly inside the group */
      const uint64_t ot = qk_shfl_xor64(best, m);
      const uint32_t oi = QK_SHFL_XOR32(bidx, m);
      if (ot < best || (ot == best && oi < bidx)) {
        best = ot;
        bidx = oi;
      }
    }
#endif
    if (leader && rem == 0 && !done) {
      int bi = best == QK_TIME_NONE ? -1 : (int)bidx;
      if (FT != 0 && ext_cursor < n_ext) { /* rank 0 = externally scheduled events: they win ties */
        const uint64_t te = cx.ext[ext_cursor].t;
        if (te <= best) {
          best = te;
          bi = -1;
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
          /* ScheduledEventConfig -> BasicEnvironment::set_state / SetLowCapacity (core.rs:256-265, 1382-1396) */
          const DevExt* e = &cx.ext[ext_cursor];
          ext_cursor += 1;
          const uint32_t target = e->target;
          const DevHotA a = QK_HOT_A(target);
          if (e->type == 1) {
            W64(cx.vecs[cx.comps[target].vec_idx].o64_low) = QK_D2U(e->value);
            const double zero[3] = {0.0, 0.0, 0.0};
            qk_vstock_add<LOG, SM>(cx, target, zero, SRC_SCH);
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
              S64(G64_FIRE0 + bi) = QK_TIME_NONE;
              stand = true;
            }
            rem = 1;
            ptr = ident_off + comp;
          }
        }
      } /* an action was popped */
    }
    QK_SYNCWARP();
    } /* any_want */

    /* -- QUICK: the calls at the head of each group's list, G at a time, one lane each */
    for (;;) {
      const uint32_t g_rem = QK_SHFL32((leader && !stand) ? rem : 0u, lead);
      if (!QK_ANY_SYNC(g_rem != 0)) break;
      const uint32_t g_ptr = QK_SHFL32(ptr, lead);
      const uint64_t g_now = qk_shfl64(cx.now, lead);
      const uint32_t n_here = g_rem < (uint32_t)G ? g_rem : (uint32_t)G;
      uint32_t code = 1u;
      uint64_t tick_dt = 0;
      if (j < n_here) code = qk_quick_eval<LOG, SM, FT>(cx, g_ptr + j, g_now, &tick_dt);
      const uint32_t full = (QK_BALLOT(code == 0u) >> gshift) & gmask;
      const uint32_t n_take = full ? (uint32_t)__ffs((int)full) - 1u : n_here; /* calls disposed of right here */
      if (FT != 0 && code == 2u && j < n_take) qk_dm_apply<SM>(cx, cx.subs[g_ptr + j], g_now, tick_dt);
#ifdef QK_HOST_EMUL
      if (FT != 0 && code >= 0x100u && j < n_take && (cx.subq[g_ptr + j].flags & SQ_DM)) qk_emul_count(1);
#endif
      if (LOG) {
        /* their ProcessNonStart records, in list order: lane j's is record number (popcount of the writing lanes
         * below it) after the replication's current count */
        const uint64_t g_src = qk_shfl64(src, lead);
        const uint32_t g_cnt = QK_SHFL32(cx.log_cnt, lead);
        bool writes = false;
        uint32_t comp = 0, seq = 0;
        if (j < n_take && code >= 0x100u) {
          comp = cx.subs[g_ptr + j];
          const DevHotA a = QK_HOT_A(comp);
          seq = S32(a.olog32 + L32_SEQ);
          S32(a.olog32 + L32_SEQ) = seq + 1;
          writes = (a.flags & DC_LOGGED) != 0;
        }
        const uint32_t wmask = (QK_BALLOT(writes) >> gshift) & gmask;
        if (writes) {
          const uint32_t pos = g_cnt + (uint32_t)QK_POPC(wmask & ((1u << j) - 1u));
          uint4* r = cx.log_base + 2 * (size_t)(pos & cx.log_mask);
          uint4 a, b;
          a.x = (uint32_t)g_now;
          a.y = (uint32_t)(g_now >> 32);
          a.z = comp | ((uint32_t)QK_LOG_PROCESS_NONSTART << 16) | ((code & 0xFFu) << 24);
          a.w = seq;
          b.x = (uint32_t)(g_src >> 32) & 0xFFFFu;
          b.y = (uint32_t)g_src;
          b.z = 0;
          b.w = 0;
          qk_store_record(r, a, b);
        }
        if (leader) cx.log_cnt += (uint32_t)QK_POPC(wmask);
      }
      if (leader && g_rem != 0) {
        ptr += n_take;
        rem -= n_take;
        if (full) stand = true; /* the head of the list is now a call for the full handler */
      }
      QK_SYNCWARP();
    }
  }
}

constexpr int qk_group_max_threads(int g) { return g <= 4 ? 256 : (g == 8 ? 512 : 768); }

template <bool LOG, int G, int FT>
__global__ void __launch_bounds__(qk_group_max_threads(G), 1) qk_group_kernel(const __grid_constant__ KParams P) {
  constexpr int SM = QK_SMG;
  static_assert(G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group width");
  uint8_t* const smem = qk_smem;
  const uint32_t tid = QK_TID, nt = QK_NT;
  const uint32_t R = P.grp_reps;           /* replications per block: nt == R * G */
  const uint32_t grp = tid / G, j = tid % G; /* replication slot in the block ; lane within the group */
  const uint32_t lane = tid % QK_WARP;
  const uint32_t lead = lane & ~(uint32_t)(G - 1); /* warp lane of this group's lane 0 */
  const bool leader = j == 0;
  const uint32_t gshift = lead; /* this group's bits inside a warp ballot */
  const uint32_t gmask = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);

  const uint64_t rep0 = (uint64_t)(P.tile0 + QK_BID) * R; /* first replication of this block */

  /* ---- stage in: model tables (plain copies) and, unless this launch initialises, every slot row of the block's
   * replications with one bulk asynchronous copy each */
  uint64_t* const s64 = reinterpret_cast<uint64_t*>(smem + P.table_bytes);
  uint32_t* const s32 = reinterpret_cast<uint32_t*>(smem + P.table_bytes + (size_t)(P.n64 + P.grp_x64) * P.grp_stride64 * 8);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + P.grp_bar_off);
  (void)bar;
#ifndef QK_HOST_EMUL
  if (!P.do_init) {
    if (tid == 0) {
      qk_mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    QK_SYNCBLOCK();
    if (tid == 0) qk_mbar_expect_tx(bar, (P.n64 * 8u + P.n32 * 4u) * R);
    QK_SYNCBLOCK();
    for (uint32_t s = tid; s < P.n64 + P.n32; s += nt) {
      if (s < P.n64)
        qk_bulk_load(s64 + (size_t)s * P.grp_stride64, P.g64 + (size_t)s * P.rep_pad + rep0, R * 8u, bar);
      else
        qk_bulk_load(s32 + (size_t)(s - P.n64) * P.grp_stride32, P.g32 + (size_t)(s - P.n64) * P.rep_pad + rep0, R * 4u,
                     bar);
    }
  }
#else
  if (!P.do_init && tid == 0) {
    for (uint32_t s = 0; s < P.n64; ++s)
      memcpy(s64 + (size_t)s * P.grp_stride64, P.g64 + (size_t)s * P.rep_pad + rep0, R * 8u);
    for (uint32_t s = 0; s < P.n32; ++s)
      memcpy(s32 + (size_t)s * P.grp_stride32, P.g32 + (size_t)s * P.rep_pad + rep0, R * 4u);
  }
#endif
  {
    const uint4* srcp = reinterpret_cast<const uint4*>(P.tables);
    uint4* dstp = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = tid; i < P.table_bytes / 16; i += nt) dstp[i] = srcp[i];
  }
#ifndef QK_HOST_EMUL
  if (!P.do_init) qk_mbar_wait(bar, 0);
#endif
  QK_SYNCBLOCK();

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
  const uint64_t rep_local = rep0 + grp;
  cx.g64 = nullptr;
  cx.g32 = nullptr;
  cx.stride = P.grp_stride64;
  cx.stride32 = P.grp_stride32;
  cx.base64 = P.table_bytes / 8 + grp;
  cx.base32 = (P.table_bytes + (P.n64 + P.grp_x64) * P.grp_stride64 * 8) / 4 + grp;
  const bool live = rep_local < P.n_reps;
  const uint64_t rep_global = P.rep_offset + rep_local;
  cx.rep_local = rep_local;
  cx.gs64 = P.gs64 + rep_local;
  cx.gs32 = P.gs32 + rep_local;
  cx.gc64 = P.gc64 + rep_local;
  cx.gc32 = P.gc32 + rep_local;
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
  cx.log_cnt = 0;
  cx.now = 0;
  const uint32_t n_comp = cx.hdr->n_comp, n_sched = cx.hdr->n_sched, n_ext = cx.hdr->n_ext;
  const uint32_t n_stale_words = cx.hdr->n_stale_words;
  const uint32_t n_sched_pad = (n_sched + 3u) & ~3u;

  if (P.do_init) {
    /* fresh replication: the group zeroes its state rows together, lane 0 sets up the components */
    for (uint32_t s = j; s < P.n64; s += G) S64(s) = 0;
    for (uint32_t s = j; s < P.n32; s += G) S32(s) = 0;
    for (uint32_t s = j; s < P.n64c; s += G) C64(s) = 0;
    for (uint32_t s = j; s < P.n32c; s += G) C32(s) = 0;
    for (uint32_t s = j; s < n_comp * ST_PER_COMP; s += G) {
      cx.gs64[(size_t)s * P.rep_pad] = 0;
      cx.gs32[(size_t)s * P.rep_pad] = 0;
    }
    QK_SYNCWARP();
    cx.now = P.t0;
    if (leader) {
      for (uint32_t i = 0; i < n_sched_pad; ++i) S64(G64_FIRE0 + i) = QK_TIME_NONE; /* incl. the padding slots */
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
            const uint32_t n_init = c->dist_time; /* with_initial_resource: the count rides in DevComp.dist_time */
            for (uint32_t k = 0; k < n_init; ++k) C32(c->ocold32 + k) = (63u << 26) | (initial_base + k);
            if (FT >= 2 && (c->flags & DC_CONTAINER)) {
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
            qk_vstore<SM>(cx, v->o64_res, v->dim, v->vinit);
            W64(v->o64_low) = QK_D2U(v->vlow);
            W64(v->o64_last) = P.t0;
          } else if (FT != 0 && c->kind == QK_VECTOR_PROCESS && cx.vecs[c->vec_idx].o32_qty != 0xFFFF) {
            W32(cx.vecs[c->vec_idx].o32_qty) = cx.vecs[c->vec_idx].dist_qty;
          }
        }
      }
      S64(G64_NEVENTS) = 0;
    }
  } else if (leader) {
    cx.now = S64(G64_NOW);
    cx.status = S32(G32_STATUS);
    cx.log_cnt = S32(G32_LOGCNT);
  }
  QK_SYNCWARP();
  if (leader)
    for (uint32_t w = 0; w < cx.n_pf_words; ++w) cx.npend += (uint32_t)QK_POPC(S32(cx.pf32 + w));

  /* ---- event loop: the leader's loop state is the replication's; the other lanes carry rem = 0, done = true and
   * take part in the collectives of every trip */
  constexpr bool EV32 = LOG;
  uint64_t events_done = 0;
  const uint64_t budget = (live && leader) ? P.event_budget : 0;
  const bool unlimited = P.event_budget == ~0ull;
  const uint32_t ev_initial = (live && leader) ? (unlimited ? 0xFFFFFFFFu : (uint32_t)P.event_budget) : 0u;
  uint32_t ev_left = ev_initial;
  const uint64_t t_until = P.t_until;
  uint32_t rem = (P.do_init && live && leader) ? cx.hdr->n_init : 0, ptr = cx.hdr->init_off;
  const uint32_t ident_off = cx.hdr->ident_off;
  uint64_t src = SRC_INIT;
  uint32_t ext_cursor = (FT != 0 && leader) ? S32(G32_EXT_CURSOR) : 0u;
  bool done = !leader || cx.status != 0;
  if (done) rem = 0;
  /* stand: the call at the head of the group's list is known to need the full handler (a component's own keyed event,
   * whose fire slot was just cleared so that "fire == NONE means idle" does not hold; or a call qk_quick_eval sent on) */
  bool stand = false;
  const uint16_t* pf_comp = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(cx.hdr) + cx.hdr->off_pf);
  for (;;) {
    QK_SYNCWARP();
    qk_scan_phase<LOG, G, FT>(cx, rem, ptr, src, stand, done, ev_left, events_done, ext_cursor, leader, lead, j, gshift, gmask,
                              live, unlimited, budget, t_until, n_sched, n_ext, n_stale_words, ident_off);
    QK_SYNCWARP();
    if (!QK_ANY_SYNC(!done || rem != 0)) break; /* uniform exit: every group of the warp has finished */

    /* -- REFILL (deferred variate prefetch): starts consume a pre-drawn duration.  When some group's next call may start
     * a process whose slot is still empty (a process that finishes and restarts in one call is the usual case), EVERY
     * group of the warp re-draws all its pending slots, one per lane: the sampler -- Philox, inverse CDF, from_secs_f64,
     * ~150 instructions -- then runs once for several lanes instead of once per lane.  (Measured alternatives: re-draw
     * every trip = more sampler passes, 1.52e9 -> 1.37e9 events/s on the haulage network; per-group threshold of G / 2
     * pending = 1.2 lanes per sampler instruction.) */
    {
#if QK_GROUP_REFILL_EAGER
      const bool go = QK_SHFL32((leader && cx.npend != 0) ? 1u : 0u, lead) != 0;
#else
      const bool must = leader && rem != 0 && cx.npend != 0 && qk_needs_refill_now<SM>(cx, cx.subs[ptr]);
      const bool go = QK_ANY_SYNC(must) && QK_SHFL32((leader && cx.npend != 0) ? 1u : 0u, lead) != 0;
#endif
      if (QK_ANY_SYNC(go)) {
        if (go) {
          uint32_t k = 0;
          for (uint32_t w = 0; w < cx.n_pf_words; ++w) {
            uint32_t m = S32(cx.pf32 + w);
            while (m) {
              const uint32_t i = (uint32_t)__ffs((int)m) - 1;
              m &= m - 1;
              if ((k++ % G) == j) {
                const DevHotB b = QK_HOT_B(pf_comp[w * 32 + i]);
                W64(b.o64_nextdur) = qk_sample_duration_code<SM>(cx, b.dist_time);
              }
            }
          }
        }
        QK_SYNCWARP();
        if (go && leader) {
          for (uint32_t w = 0; w < cx.n_pf_words; ++w) S32(cx.pf32 + w) = 0;
          cx.npend = 0;
        }
      }
    }
    QK_SYNCWARP();
    /* -- UPDATE: one full update_state per group that stands on a call */
    if (leader && rem != 0) {
      const uint32_t p = cx.subs[ptr];
      ptr += 1;
      rem -= 1;
      const bool fault = qk_update_state<LOG, SM, FT>(cx, p, src);
      if (fault) {
        rem = 0;
        done = true;
      }
    }
    stand = false;
  }
  /* step_until(t) leaves the clock at t */
  if (leader) {
    if (!cx.status && live && P.t_until != QK_TIME_NONE && P.t_until > cx.now) cx.now = P.t_until;
    if constexpr (LOG) qk_async_wait(); /* head-item copies still in flight land before the write-back */
    S64(G64_NOW) = cx.now;
    S64(G64_NEVENTS) += EV32 ? (uint64_t)(ev_initial - ev_left) : events_done;
    S32(G32_STATUS) = cx.status;
    if (FT != 0) S32(G32_EXT_CURSOR) = ext_cursor;
    S32(G32_LOGCNT) = cx.log_cnt;
  }
  /* ---- stage out: one bulk asynchronous store per slot row */
#ifndef QK_HOST_EMUL
  qk_fence_async_proxy();
  QK_SYNCBLOCK();
  bool stored = false;
  for (uint32_t s = tid; s < P.n64 + P.n32; s += nt) {
    if (s < P.n64)
      qk_bulk_store(P.g64 + (size_t)s * P.rep_pad + rep0, s64 + (size_t)s * P.grp_stride64, R * 8u);
    else
      qk_bulk_store(P.g32 + (size_t)(s - P.n64) * P.rep_pad + rep0, s32 + (size_t)(s - P.n64) * P.grp_stride32, R * 4u);
    stored = true;
  }
  if (stored) qk_bulk_store_wait();
#else
  QK_SYNCBLOCK();
  if (tid == 0) {
    for (uint32_t s = 0; s < P.n64; ++s)
      memcpy(P.g64 + (size_t)s * P.rep_pad + rep0, s64 + (size_t)s * P.grp_stride64, R * 8u);
    for (uint32_t s = 0; s < P.n32; ++s)
      memcpy(P.g32 + (size_t)s * P.rep_pad + rep0, s32 + (size_t)s * P.grp_stride32, R * 4u);
  }
#endif
}


/* ==================================== the PHASED variant ====================================
 * Same placement (a block's R replications in shared memory, G lanes per replication for the scheduler pop and the
 * subscriber broadcast), but the sequential part of an event -- the full update_state handler and the variate sampler,
 * two thirds of the lane-group kernel's warp instructions at 2-5 active lanes of 32 -- is not run by lane 0 of every group
 * (4 busy lanes per warp with G = 8): each trip of the loop has two block-wide phases,
 *     P  pop + quick    thread t works for replication t / G as lane t % G (qk_scan_phase, as in the lane-group kernel);
 *     U  refill + update thread t < R works for replication t: the handlers of 32 replications run in ONE warp, converged
 *                        the way the thread-per-replication kernel's UPDATE phase is,
 * and the loop state of a replication (call list, source id, flags, counters) lives in shared-memory rows next to its
 * state rows so that it can change hands between the phases.  Several smaller blocks per SM overlap one block's U phase
 * (one or two busy warps) with the others' P phases. */
enum { X64_SRC = 0, X64_EVD = 1, X64_COUNT = 2 };
enum { X32_REM = 0, X32_PTR = 1, X32_FLAGS = 2, X32_NPEND = 3, X32_EVLEFT = 4, X32_EXT = 5, X32_COUNT = 6 };
#define XF_STAND 1u
#define XF_DONE 2u
#define X64(k) S64(P.n64 + (k))
#define X32(k) S32(P.n32 + (k))

#ifndef QK_HOST_EMUL
#define QK_SYNCBLOCK_OR(p) (__syncthreads_or((p) ? 1 : 0) != 0)
#else
#define QK_SYNCBLOCK_OR(p) (qk_emul_ballot(p) != 0u)
#endif

template <bool LOG>
__device__ __forceinline__ void qk_bind_rep(Ctx& cx, const KParams& P, uint32_t slot, uint64_t rep0) {
  const uint64_t rep_local = rep0 + slot;
  const uint64_t rep_global = P.rep_offset + rep_local;
  cx.base64 = P.table_bytes / 8 + slot;
  cx.base32 = (P.table_bytes + (P.n64 + P.grp_x64) * P.grp_stride64 * 8) / 4 + slot;
  cx.rep_local = rep_local;
  cx.gs64 = P.gs64 + rep_local;
  cx.gs32 = P.gs32 + rep_local;
  cx.gc64 = P.gc64 + rep_local;
  cx.gc32 = P.gc32 + rep_local;
  cx.rep_lo = (uint32_t)rep_global;
  cx.rep_hi = (uint32_t)(rep_global >> 32);
  cx.log_base = LOG ? P.log + 2 * (size_t)rep_local * P.log_cap : nullptr;
}

template <bool LOG, int G, int FT>
__global__ void __launch_bounds__(qk_group_max_threads(G), 1) qk_phased_kernel(const __grid_constant__ KParams P) {
  constexpr int SM = QK_SMG;
  constexpr bool EV32 = LOG;
  uint8_t* const smem = qk_smem;
  const uint32_t tid = QK_TID, nt = QK_NT;
  const uint32_t R = P.grp_reps;
  const uint32_t grp = tid / G, j = tid % G;
  const uint32_t lane = tid % QK_WARP;
  const uint32_t lead = lane & ~(uint32_t)(G - 1);
  const bool leader = j == 0;
  const uint32_t gshift = lead;
  const uint32_t gmask = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
  const bool is_rep = tid < R;                          /* U phase: this thread works for replication `tid` */
  const bool rep_warp = (tid - lane) < R;               /* ... and this warp has such threads */
  const uint64_t rep0 = (uint64_t)(P.tile0 + QK_BID) * R;

  uint64_t* const s64 = reinterpret_cast<uint64_t*>(smem + P.table_bytes);
  uint32_t* const s32 = reinterpret_cast<uint32_t*>(smem + P.table_bytes + (size_t)(P.n64 + P.grp_x64) * P.grp_stride64 * 8);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + P.grp_bar_off);
  (void)bar;
#ifndef QK_HOST_EMUL
  if (!P.do_init) {
    if (tid == 0) {
      qk_mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    QK_SYNCBLOCK();
    if (tid == 0) qk_mbar_expect_tx(bar, (P.n64 * 8u + P.n32 * 4u) * R);
    for (uint32_t s = tid; s < P.n64 + P.n32; s += nt) {
      if (s < P.n64)
        qk_bulk_load(s64 + (size_t)s * P.grp_stride64, P.g64 + (size_t)s * P.rep_pad + rep0, R * 8u, bar);
      else
        qk_bulk_load(s32 + (size_t)(s - P.n64) * P.grp_stride32, P.g32 + (size_t)(s - P.n64) * P.rep_pad + rep0, R * 4u,
                     bar);
    }
  }
#else
  if (!P.do_init && tid == 0) {
    for (uint32_t s = 0; s < P.n64; ++s)
      memcpy(s64 + (size_t)s * P.grp_stride64, P.g64 + (size_t)s * P.rep_pad + rep0, R * 8u);
    for (uint32_t s = 0; s < P.n32; ++s)
      memcpy(s32 + (size_t)s * P.grp_stride32, P.g32 + (size_t)s * P.rep_pad + rep0, R * 4u);
  }
#endif
  {
    const uint4* srcp = reinterpret_cast<const uint4*>(P.tables);
    uint4* dstp = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = tid; i < P.table_bytes / 16; i += nt) dstp[i] = srcp[i];
  }
#ifndef QK_HOST_EMUL
  if (!P.do_init) qk_mbar_wait(bar, 0);
#endif
  QK_SYNCBLOCK();

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
  cx.g64 = nullptr;
  cx.g32 = nullptr;
  cx.stride = P.grp_stride64;
  cx.stride32 = P.grp_stride32;
  cx.rep_pad = (uint32_t)P.rep_pad;
  cx.key0 = (uint32_t)P.seed;
  cx.key1 = (uint32_t)(P.seed >> 32);
  cx.tapes = P.tapes;
  cx.tape_len = P.tape_len;
  cx.log_mask = P.log_cap ? P.log_cap - 1 : 0;
  cx.status = 0;
  cx.npend = 0;
  cx.log_cnt = 0;
  cx.now = 0;
  const uint32_t n_comp = cx.hdr->n_comp, n_sched = cx.hdr->n_sched, n_ext = cx.hdr->n_ext;
  const uint32_t n_stale_words = cx.hdr->n_stale_words;
  const uint32_t n_sched_pad = (n_sched + 3u) & ~3u;
  const uint32_t ident_off = cx.hdr->ident_off;
  const bool unlimited = P.event_budget == ~0ull;
  const uint64_t t_until = P.t_until;
  const uint16_t* pf_comp = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(cx.hdr) + cx.hdr->off_pf);

  /* ---- the replications' loop state, set up by their U-phase threads */
  if (P.do_init) {
    /* fresh replications: every thread zeroes a share of the block's rows (and of its replications' global planes) */
    for (uint32_t i = tid; i < (P.n64 + P.grp_x64) * P.grp_stride64; i += nt) s64[i] = 0;
    for (uint32_t i = tid; i < (P.n32 + X32_COUNT) * P.grp_stride32; i += nt) s32[i] = 0;
    for (uint32_t i = tid; i < P.n64c * R; i += nt) P.gc64[(size_t)(i / R) * P.rep_pad + rep0 + i % R] = 0;
    for (uint32_t i = tid; i < P.n32c * R; i += nt) P.gc32[(size_t)(i / R) * P.rep_pad + rep0 + i % R] = 0;
    for (uint32_t i = tid; i < n_comp * ST_PER_COMP * R; i += nt) {
      P.gs64[(size_t)(i / R) * P.rep_pad + rep0 + i % R] = 0;
      P.gs32[(size_t)(i / R) * P.rep_pad + rep0 + i % R] = 0;
    }
    QK_SYNCBLOCK();
  }
  if (is_rep) {
    qk_bind_rep<LOG>(cx, P, tid, rep0);
    const bool live = cx.rep_local < P.n_reps;
    uint32_t rem = 0;
    if (P.do_init) {
      cx.now = P.t0;
      for (uint32_t i = 0; i < n_sched_pad; ++i) S64(G64_FIRE0 + i) = QK_TIME_NONE; /* incl. the padding slots */
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
            const uint32_t n_init = c->dist_time; /* with_initial_resource: the count rides in DevComp.dist_time */
            for (uint32_t k = 0; k < n_init; ++k) C32(c->ocold32 + k) = (63u << 26) | (initial_base + k);
            if (FT >= 2 && (c->flags & DC_CONTAINER)) {
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
            qk_vstore<SM>(cx, v->o64_res, v->dim, v->vinit);
            W64(v->o64_low) = QK_D2U(v->vlow);
            W64(v->o64_last) = P.t0;
          } else if (FT != 0 && c->kind == QK_VECTOR_PROCESS && cx.vecs[c->vec_idx].o32_qty != 0xFFFF) {
            W32(cx.vecs[c->vec_idx].o32_qty) = cx.vecs[c->vec_idx].dist_qty;
          }
        }
        rem = cx.hdr->n_init;
      }
      S64(G64_NOW) = cx.now;
      S64(G64_NEVENTS) = 0;
      S32(G32_STATUS) = cx.status;
      S32(G32_LOGCNT) = 0;
    }
    uint32_t npend = 0;
    for (uint32_t w = 0; w < cx.n_pf_words; ++w) npend += (uint32_t)QK_POPC(S32(cx.pf32 + w));
    const bool frozen = !live || S32(G32_STATUS) != 0; /* a faulted replication stays frozen */
    X32(X32_REM) = frozen ? 0u : rem;
    X32(X32_PTR) = cx.hdr->init_off;
    X32(X32_FLAGS) = frozen ? XF_DONE : 0u;
    X32(X32_NPEND) = npend;
    X32(X32_EVLEFT) = !live ? 0u : (unlimited ? 0xFFFFFFFFu : (uint32_t)P.event_budget);
    X32(X32_EXT) = FT != 0 ? S32(G32_EXT_CURSOR) : 0u;
    X64(X64_SRC) = SRC_INIT;
    X64(X64_EVD) = 0;
  }
  QK_SYNCBLOCK();

  /* ---- event loop */
  for (;;) {
    /* -- phase P: pop + quick, a lane group per replication */
    {
      qk_bind_rep<LOG>(cx, P, grp, rep0);
      const bool live = cx.rep_local < P.n_reps;
      uint32_t rem = 0, ptr = 0, ev_left = 0, ext_cursor = 0;
      uint64_t src = 0, events_done = 0;
      bool stand = false, done = true;
      if (leader) {
        rem = X32(X32_REM);
        ptr = X32(X32_PTR);
        const uint32_t f = X32(X32_FLAGS);
        stand = (f & XF_STAND) != 0;
        done = (f & XF_DONE) != 0;
        ev_left = X32(X32_EVLEFT);
        ext_cursor = X32(X32_EXT);
        src = X64(X64_SRC);
        events_done = X64(X64_EVD);
        cx.now = S64(G64_NOW);
        cx.status = S32(G32_STATUS);
        cx.log_cnt = S32(G32_LOGCNT);
      }
      const uint64_t budget = (live && leader) ? P.event_budget : 0;
      QK_SYNCWARP();
      qk_scan_phase<LOG, G, FT>(cx, rem, ptr, src, stand, done, ev_left, events_done, ext_cursor, leader, lead, j, gshift,
                                gmask, live, unlimited, budget, t_until, n_sched, n_ext, n_stale_words, ident_off);
      if (leader) {
        if (cx.status) { /* a fault inside the pop (SetLowCapacity's add) */
          rem = 0;
          done = true;
        }
        X32(X32_REM) = rem;
        X32(X32_PTR) = ptr;
        X32(X32_FLAGS) = (stand ? XF_STAND : 0u) | (done ? XF_DONE : 0u);
        X32(X32_EVLEFT) = ev_left;
        X32(X32_EXT) = ext_cursor;
        X64(X64_SRC) = src;
        X64(X64_EVD) = events_done;
        S64(G64_NOW) = cx.now;
        S32(G32_STATUS) = cx.status;
        S32(G32_LOGCNT) = cx.log_cnt;
      }
      if (!QK_SYNCBLOCK_OR(leader && (!done || rem != 0))) break; /* every replication of the block has finished */
    }
    /* -- phase U: refill + one full update_state per replication that stands on a call, a thread per replication */
    if (rep_warp) {
      uint32_t rem = 0, ptr = 0;
      if (is_rep) {
        qk_bind_rep<LOG>(cx, P, tid, rep0);
        rem = X32(X32_REM);
        ptr = X32(X32_PTR);
        cx.npend = X32(X32_NPEND);
        cx.now = S64(G64_NOW);
        cx.status = S32(G32_STATUS);
        cx.log_cnt = S32(G32_LOGCNT);
      } else {
        cx.npend = 0;
      }
      {
        const bool must = rem != 0 && cx.npend != 0 && qk_needs_refill_now<SM>(cx, cx.subs[ptr]);
        if (QK_ANY_SYNC(must)) {
          if (cx.npend != 0) qk_refill<SM>(cx);
        }
      }
      QK_SYNCWARP();
      if (rem != 0) {
        const uint32_t p = cx.subs[ptr];
        const uint64_t src = X64(X64_SRC);
        ptr += 1;
        rem -= 1;
        const bool fault = qk_update_state<LOG, SM, FT>(cx, p, src);
        uint32_t f = X32(X32_FLAGS) & ~XF_STAND;
        if (fault) {
          rem = 0;
          f |= XF_DONE;
        }
        X32(X32_REM) = rem;
        X32(X32_PTR) = ptr;
        X32(X32_FLAGS) = f;
        S32(G32_STATUS) = cx.status;
        S32(G32_LOGCNT) = cx.log_cnt;
      }
      if (is_rep) X32(X32_NPEND) = cx.npend;
    }
    QK_SYNCBLOCK();
  }
  /* step_until(t) leaves the clock at t; the event counters go back to their state slots */
  if (is_rep) {
    qk_bind_rep<LOG>(cx, P, tid, rep0);
    const bool live = cx.rep_local < P.n_reps;
    uint64_t now = S64(G64_NOW);
    if (!S32(G32_STATUS) && live && P.t_until != QK_TIME_NONE && P.t_until > now) S64(G64_NOW) = P.t_until;
    if constexpr (LOG) qk_async_wait(); /* head-item copies still in flight land before the write-back */
    const uint32_t ev_initial = !live ? 0u : (unlimited ? 0xFFFFFFFFu : (uint32_t)P.event_budget);
    S64(G64_NEVENTS) += EV32 ? (uint64_t)(ev_initial - X32(X32_EVLEFT)) : X64(X64_EVD);
    if (FT != 0) S32(G32_EXT_CURSOR) = X32(X32_EXT);
  }
#ifndef QK_HOST_EMUL
  qk_fence_async_proxy();
  QK_SYNCBLOCK();
  bool stored = false;
  for (uint32_t s = tid; s < P.n64 + P.n32; s += nt) {
    if (s < P.n64)
      qk_bulk_store(P.g64 + (size_t)s * P.rep_pad + rep0, s64 + (size_t)s * P.grp_stride64, R * 8u);
    else
      qk_bulk_store(P.g32 + (size_t)(s - P.n64) * P.rep_pad + rep0, s32 + (size_t)(s - P.n64) * P.grp_stride32, R * 4u);
    stored = true;
  }
  if (stored) qk_bulk_store_wait();
#else
  QK_SYNCBLOCK();
  if (tid == 0) {
    for (uint32_t s = 0; s < P.n64; ++s)
      memcpy(P.g64 + (size_t)s * P.rep_pad + rep0, s64 + (size_t)s * P.grp_stride64, R * 8u);
    for (uint32_t s = 0; s < P.n32; ++s)
      memcpy(P.g32 + (size_t)s * P.rep_pad + rep0, s32 + (size_t)s * P.grp_stride32, R * 4u);
  }
#endif
}

#endif
