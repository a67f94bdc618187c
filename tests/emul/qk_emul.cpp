/*
 * qk_emul.cpp -- TEST INFRASTRUCTURE ONLY (see qk_host_emul.h).
 *
 * Implements the batch half of the C ABI (include/quokka_b200.h) on the CPU by compiling the REAL kernel
 * source quokkasim_b200/csrc/qk_kernels.cuh with g++ and running it one replication per "block" (nt == 1).
 * The model half (qk_model_*) is the product's own host code (qk_model.cpp), linked in unchanged.
 * Used by `pytest -m "not gpu"` to check kernel logic against the oracle without a GPU.  Not shipped, not
 * loaded by the quokkasim_b200 package.
 */
#define QK_HOST_EMUL 1
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <thread>

#include "qk_group.cuh"
#include "qk_model.h"

void qk_emul_fail(const char* what) {
  fprintf(stderr, "qk_emul: %s\n", what);
  abort();
}
std::atomic<uint64_t> qk_emul_counters[4];
extern "C" uint64_t qk_emul_counter(int k) { return qk_emul_counters[k].load(); }
thread_local uint8_t* qk_emul_smem = nullptr;
thread_local uint32_t qk_emul_bid = 0;
thread_local QkEmulTeam* qk_emul_team = nullptr;
thread_local uint32_t qk_emul_tid = 0, qk_emul_nt = 1;

struct qk_batch {
  qk_batch_config cfg;
  QkLowered low;
  bool log;
  uint32_t table_bytes_padded;
  uint64_t rep_pad;
  std::vector<uint64_t> g64;
  std::vector<uint32_t> g32;
  std::vector<uint64_t> gs64;
  std::vector<uint32_t> gs32;
  std::vector<uint64_t> gc64;
  std::vector<uint32_t> gc32;
  std::vector<uint8_t> tables;
  std::vector<uint4> logbuf;
  std::vector<std::vector<double> > tapes;
  std::vector<const double*> tape_ptr;
  std::vector<uint64_t> tape_len;
  bool any_tape;
  bool initialised;
  uint32_t lanes; /* 0: one "thread" per replication ; else the lane-group kernel with this group width */
  bool phased;    /* ... its phased variant */
  QkGroupGeom geo;
};

/* the lane-group kernel: one block = geo.reps replications x `lanes` lanes = one team of host threads */
template <bool L, int G>
static void group_block(const KParams& P, uint32_t ft, bool phased) {
  if (phased) {
    if (ft >= 2) qk_phased_kernel<L, G, 2>(P);
    else if (ft == 1) qk_phased_kernel<L, G, 1>(P);
    else qk_phased_kernel<L, G, 0>(P);
    return;
  }
  if (ft >= 2) qk_group_kernel<L, G, 2>(P);
  else if (ft == 1) qk_group_kernel<L, G, 1>(P);
  else qk_group_kernel<L, G, 0>(P);
}
static int run_group(qk_batch* b, KParams& P, uint32_t tile0, uint32_t n_tiles) {
  P.grp_reps = b->geo.reps;
  P.grp_stride64 = b->geo.stride64;
  P.grp_stride32 = b->geo.stride32;
  P.grp_bar_off = b->geo.bar_off;
  P.grp_x64 = b->phased ? 2u : 0u;
  const bool phased = b->phased;
  const uint32_t nt = b->geo.threads, ft = b->low.hdr.features, G = b->lanes;
  const uint32_t blocks = n_tiles ? n_tiles : (uint32_t)(b->rep_pad / b->geo.reps);
  std::vector<uint64_t> smem(b->geo.smem_bytes / 8 + 2);
  for (uint32_t blk = 0; blk < blocks; ++blk) {
    QkEmulTeam team;
    team.n = nt;
    team.count.store(0);
    team.gen.store(0);
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        qk_emul_smem = reinterpret_cast<uint8_t*>(smem.data());
        qk_emul_bid = blk;
        qk_emul_tid = t;
        qk_emul_nt = nt;
        qk_emul_team = &team;
        if (b->log) {
          if (G == 2) group_block<true, 2>(P, ft, phased);
          else if (G == 4) group_block<true, 4>(P, ft, phased);
          else group_block<true, 8>(P, ft, phased);
        } else {
          if (G == 2) group_block<false, 2>(P, ft, phased);
          else if (G == 4) group_block<false, 4>(P, ft, phased);
          else group_block<false, 8>(P, ft, phased);
        }
      });
    for (auto& x : th) x.join();
  }
  (void)tile0;
  return QK_OK;
}

static int run(qk_batch* b, uint32_t do_init, uint64_t t0, uint64_t t_until, uint64_t budget, uint32_t tile0 = 0,
               uint32_t n_tiles = 0) {
  KParams P;
  memset(&P, 0, sizeof(P));
  P.g64 = b->g64.data();
  P.g32 = b->g32.data();
  P.gs64 = b->gs64.data();
  P.gs32 = b->gs32.data();
  P.gc64 = b->gc64.data();
  P.gc32 = b->gc32.data();
  P.n64c = b->low.hdr.n64c;
  P.n32c = b->low.hdr.n32c;
  P.tables = b->tables.data();
  P.table_bytes = b->table_bytes_padded;
  P.n64 = b->low.hdr.n64;
  P.n32 = b->low.hdr.n32;
  P.n_reps = b->cfg.n_reps;
  P.rep_pad = b->rep_pad;
  P.rep_offset = b->cfg.rep_offset;
  P.seed = b->cfg.base_seed;
  P.log = b->log ? b->logbuf.data() : nullptr;
  P.log_cap = b->log ? b->cfg.log_capacity : 0;
  P.tapes = b->any_tape ? b->tape_ptr.data() : nullptr;
  P.tape_len = b->tape_len.data();
  P.t0 = t0;
  P.t_until = t_until;
  P.event_budget = budget;
  P.do_init = do_init;
  P.tile0 = tile0;
  if (b->lanes) return run_group(b, P, tile0, n_tiles);
  const size_t smem_bytes = b->table_bytes_padded + (size_t)P.n64 * 8 + (size_t)P.n32 * 4;
  std::vector<uint64_t> smem((smem_bytes + 7) / 8 + 2);
  qk_emul_smem = reinterpret_cast<uint8_t*>(smem.data());
  for (uint64_t r = 0; r < (n_tiles ? n_tiles : b->rep_pad); ++r) {
    qk_emul_bid = (uint32_t)r;
    const bool hbm = (b->cfg.flags & QK_BATCH_STATE_IN_HBM) != 0;
    const bool split = (b->cfg.flags & QK_BATCH_STATE_SPLIT) != 0;
    const bool qsplit = split && (b->cfg.flags & QK_BATCH_QUEUE_TWO_TIER) != 0;
    /* the product picks FT = 0 only for on-chip state; here both state placements run both feature sets so the
     * CPU suite covers the lean kernel on every simple model */
    const uint32_t ft = b->low.hdr.features;
#define QK_EMUL_RUN(L, S)                         \
  do {                                            \
    if (ft == 3) qk_step_kernel<L, S, 3>(P);      \
    else if (ft == 2) qk_step_kernel<L, S, 2>(P); \
    else if (ft == 1) qk_step_kernel<L, S, 1>(P); \
    else qk_step_kernel<L, S, 0>(P);              \
  } while (0)
    if (b->log) {
      if (hbm) QK_EMUL_RUN(true, 0);
      else if (qsplit) QK_EMUL_RUN(true, QK_SM_QSPLIT);
      else if (split) QK_EMUL_RUN(true, QK_SM_SPLIT);
      else QK_EMUL_RUN(true, -1);
    } else {
      if (hbm) QK_EMUL_RUN(false, 0);
      else if (qsplit) QK_EMUL_RUN(false, QK_SM_QSPLIT);
      else if (split) QK_EMUL_RUN(false, QK_SM_SPLIT);
      else QK_EMUL_RUN(false, -1);
    }
  }
  qk_emul_smem = nullptr;
  return QK_OK;
}

extern "C" {

int qk_batch_create(const qk_model* m, const qk_batch_config* cfg, qk_batch** out) {
  if (!m || !cfg || !out || cfg->n_reps == 0) return QK_ERR_INVALID_ARG;
  if (cfg->log_capacity && (cfg->log_capacity & (cfg->log_capacity - 1))) return QK_ERR_INVALID_ARG;
  qk_batch* b = new qk_batch();
  b->cfg = *cfg;
  b->log = cfg->log_capacity != 0;
  int rc = qk_lower(m, b->log,
                    (cfg->flags & QK_BATCH_STATE_SPLIT) ? ((cfg->flags & QK_BATCH_QUEUE_TWO_TIER) ? 2 : 1) : 0, &b->low);
  if (rc) {
    delete b;
    return rc;
  }
  const DevHeader& h = b->low.hdr;
  b->table_bytes_padded = (h.total_bytes + 15u) & ~15u;
  b->rep_pad = cfg->n_reps;
  b->lanes = 0;
  b->phased = (cfg->flags & QK_BATCH_STATE_LANE_GROUP) && (cfg->flags & QK_BATCH_GROUP_PHASED);
  if (cfg->flags & QK_BATCH_STATE_LANE_GROUP) {
    b->lanes = cfg->lanes_per_replication ? cfg->lanes_per_replication : 4;
    if (b->lanes != 2 && b->lanes != 4 && b->lanes != 8) {
      delete b;
      return QK_ERR_INVALID_ARG;
    }
    /* small blocks: `threads_per_block` (if given) = replications per block here; the whole block is one warp */
    const uint32_t cap = cfg->threads_per_block ? cfg->threads_per_block : 2;
    if (!qk_group_geometry(b->lanes, b->lanes, h.n64, h.n32, b->phased ? 2u : 0u, b->phased ? 6u : 0u, b->table_bytes_padded,
                           32u, 1u << 30, cap, 1u, &b->geo)) {
      delete b;
      return QK_ERR_CAPACITY;
    }
    b->rep_pad = (cfg->n_reps + b->geo.reps - 1) / b->geo.reps * b->geo.reps;
  }
  b->g64.assign((size_t)h.n64 * b->rep_pad, 0);
  b->g32.assign((size_t)h.n32 * b->rep_pad, 0);
  b->gs64.assign((size_t)h.n_comp * ST_PER_COMP * b->rep_pad, 0);
  b->gs32.assign((size_t)h.n_comp * ST_PER_COMP * b->rep_pad, 0);
  b->gc64.assign((size_t)(h.n64c ? h.n64c : 1) * b->rep_pad, 0);
  b->gc32.assign((size_t)(h.n32c ? h.n32c : 1) * b->rep_pad, 0);
  b->tables.assign(b->table_bytes_padded, 0);
  memcpy(b->tables.data(), b->low.blob.data(), b->low.blob.size());
  if (b->log) b->logbuf.assign((size_t)b->rep_pad * cfg->log_capacity * 2, uint4{0, 0, 0, 0});
  b->tapes.resize(h.n_dist);
  b->tape_ptr.assign(h.n_dist ? h.n_dist : 1, nullptr);
  b->tape_len.assign(h.n_dist ? h.n_dist : 1, 0);
  b->any_tape = false;
  b->initialised = false;
  *out = b;
  return QK_OK;
}
void qk_batch_destroy(qk_batch* b) { delete b; }

int qk_batch_info_get(qk_batch* b, qk_batch_info* out) {
  if (!b || !out) return QK_ERR_INVALID_ARG;
  memset(out, 0, sizeof(*out));
  out->threads_per_block = 1;
  out->grid_blocks = (uint32_t)b->rep_pad;
  out->table_bytes = b->table_bytes_padded;
  out->slots64 = b->low.hdr.n64;
  out->slots32 = b->low.hdr.n32;
  out->state_bytes_per_rep = b->low.hdr.n64 * 8 + b->low.hdr.n32 * 4;
  out->cold_bytes_per_rep = b->low.hdr.n64c * 8 + b->low.hdr.n32c * 4;
  out->shared_bytes_per_block = out->table_bytes + out->state_bytes_per_rep;
  out->replications_padded = b->rep_pad;
  out->state_bytes_total = (uint64_t)(out->state_bytes_per_rep + out->cold_bytes_per_rep) * b->rep_pad;
  out->log_bytes_total = b->log ? (uint64_t)b->rep_pad * b->cfg.log_capacity * 32 : 0;
  out->state_in_hbm = (b->cfg.flags & QK_BATCH_STATE_IN_HBM) ? 1u : 0u;
  out->state_split = (b->cfg.flags & QK_BATCH_STATE_SPLIT) ? 1u : 0u;
  out->lanes_per_replication = b->lanes ? b->lanes : 1;
  if (b->lanes) {
    out->threads_per_block = b->geo.threads;
    out->grid_blocks = (uint32_t)(b->rep_pad / b->geo.reps);
    out->shared_bytes_per_block = b->geo.smem_bytes;
  }
  out->last_step_chunks = 1;
  return QK_OK;
}

int qk_batch_set_tape(qk_batch* b, uint32_t dist_id, const double* values, uint64_t per_rep_len) {
  if (!b || dist_id >= b->low.hdr.n_dist) return QK_ERR_INVALID_ARG;
  b->tapes[dist_id].assign(values, values + per_rep_len * b->cfg.n_reps);
  if (b->tapes[dist_id].empty()) b->tapes[dist_id].push_back(0.0);
  b->tape_ptr[dist_id] = b->tapes[dist_id].data();
  b->tape_len[dist_id] = per_rep_len;
  b->any_tape = true;
  return QK_OK;
}
int qk_batch_init(qk_batch* b, uint64_t t0_ns) {
  if (!b) return QK_ERR_INVALID_ARG;
  b->initialised = true;
  return run(b, 1, t0_ns, QK_TIME_NONE, 0);
}
int qk_batch_step_until(qk_batch* b, uint64_t t_ns) {
  if (!b || !b->initialised) return QK_ERR_STATE;
  return run(b, 0, 0, t_ns, ~0ull);
}
int qk_batch_step_events(qk_batch* b, uint64_t n) {
  if (!b || !b->initialised) return QK_ERR_STATE;
  do { /* 32-bit event counter in the kernel: finite budgets go in launches of at most 2^32 - 2 events */
    const uint64_t k = n < 0xFFFFFFFEull ? n : 0xFFFFFFFEull;
    n -= k;
    const int rc = run(b, 0, 0, QK_TIME_NONE, k);
    if (rc) return rc;
  } while (n);
  return QK_OK;
}
int qk_batch_step_events_chunked(qk_batch* b, uint64_t n, uint64_t chunk) {
  if (!b || !b->initialised || chunk == 0) return QK_ERR_STATE;
  while (n) {
    const uint64_t k = n < chunk ? n : chunk;
    n -= k;
    const int rc = run(b, 0, 0, QK_TIME_NONE, k);
    if (rc) return rc;
  }
  return QK_OK;
}
int qk_batch_step_events_sliced(qk_batch* b, uint64_t n, uint32_t n_chunks) {
  /* the product's step_sliced: tiles in parts, each part chunk by chunk (here: one after the other, tile0 offsets) */
  if (!b || !b->initialised) return QK_ERR_STATE;
  if (n_chunks == 0) return qk_batch_step_events(b, n);
  const uint64_t chunk = (n + n_chunks - 1) / n_chunks;
  const uint64_t tiles = b->lanes ? b->rep_pad / b->geo.reps : b->rep_pad;
  const uint32_t parts = tiles >= 3 ? 3 : 1;
  for (uint64_t done = 0; done < n; done += chunk)
    for (uint32_t p = 0; p < parts; ++p) {
      const uint32_t t0 = (uint32_t)(tiles * p / parts), t1 = (uint32_t)(tiles * (p + 1) / parts);
      const int rc = run(b, 0, 0, QK_TIME_NONE, std::min<uint64_t>(chunk, n - done), t0, t1 - t0);
      if (rc) return rc;
    }
  return QK_OK;
}
int qk_batch_synchronize(qk_batch*) { return QK_OK; }
int qk_batch_last_step_ms(qk_batch*, float* ms, uint32_t* n) {
  if (ms) *ms = 0;
  if (n) *n = 1;
  return QK_OK;
}

int qk_batch_summary_get(qk_batch* b, qk_batch_summary* out) {
  if (!b || !out) return QK_ERR_INVALID_ARG;
  memset(out, 0, sizeof(*out));
  out->n_reps = b->cfg.n_reps;
  out->min_now_ns = ~0ull;
  for (uint64_t r = 0; r < b->cfg.n_reps; ++r) {
    out->n_events += b->g64[(size_t)G64_NEVENTS * b->rep_pad + r];
    out->n_faulted += b->g32[(size_t)G32_STATUS * b->rep_pad + r] ? 1 : 0;
    out->n_log_records += b->g32[(size_t)G32_LOGCNT * b->rep_pad + r];
    const uint64_t now = b->g64[(size_t)G64_NOW * b->rep_pad + r];
    out->min_now_ns = std::min(out->min_now_ns, now);
    out->max_now_ns = std::max(out->max_now_ns, now);
  }
  return QK_OK;
}

int qk_batch_read_status(qk_batch* b, uint64_t rep0, uint64_t n, uint32_t* status, uint64_t* now_ns,
                         uint64_t* n_events) {
  if (!b || rep0 + n > b->cfg.n_reps) return QK_ERR_INVALID_ARG;
  for (uint64_t i = 0; i < n; ++i) {
    if (status) status[i] = b->g32[(size_t)G32_STATUS * b->rep_pad + rep0 + i];
    if (now_ns) now_ns[i] = b->g64[(size_t)G64_NOW * b->rep_pad + rep0 + i];
    if (n_events) n_events[i] = b->g64[(size_t)G64_NEVENTS * b->rep_pad + rep0 + i];
  }
  return QK_OK;
}

int qk_batch_comp_stats(qk_batch* b, uint64_t rep0, uint64_t n, uint32_t comp, qk_comp_stats* out) {
  if (!b || rep0 + n > b->cfg.n_reps || comp >= b->low.hdr.n_comp) return QK_ERR_INVALID_ARG;
  for (uint64_t i = 0; i < n; ++i)
    qk_comp_stats_dev(&b->low.comps[comp], b->low.vecs.data(), comp, b->g64.data(), b->g32.data(),
                      b->low.hdr.split ? b->gc64.data() : b->g64.data(), b->low.hdr.split ? b->gc32.data() : b->g32.data(),
                      b->low.hdr.split ? b->low.hdr.n64c : 0u, b->low.hdr.split ? b->low.hdr.n32c : 0u,
                      b->gs64.data(), b->gs32.data(), b->rep_pad, rep0 + i, &out[i]);
  return QK_OK;
}

/* the packed partials (quokka_b200.h QP_*) stated the slow, obvious way: the CPU suite uses them to check the host
 * logic that combines the ranks (gloo) and qk_fill_comp_summary's 192-bit arithmetic */
int qk_batch_reduce_partials(qk_batch* b, uint64_t* part, uint32_t n_comp) {
  if (!b || !part || n_comp != b->low.hdr.n_comp) return QK_ERR_INVALID_ARG;
  for (uint32_t c = 0; c < n_comp; ++c) {
    uint64_t* w = part + (size_t)c * QK_PARTIAL_WORDS;
    for (int k = 0; k < QK_PARTIAL_WORDS; ++k) w[k] = (k >= QP_MIN_COUNT && k <= QP_NMAX_TIME) ? ~0ull : 0ull;
    std::vector<qk_comp_stats> st(b->cfg.n_reps);
    qk_batch_comp_stats(b, 0, b->cfg.n_reps, c, st.data());
    for (uint64_t r = 0; r < b->cfg.n_reps; ++r) {
      const uint64_t t = st[r].time_ns, lo = t & 0xFFFFFFFFull, hi = t >> 32, cn = st[r].count & 0xFFFFFFFFull;
      w[QP_COUNT] += st[r].count;
      w[QP_TIME_LO] += lo;
      w[QP_TIME_HI] += hi;
      w[QP_LEVEL] += st[r].level;
      w[QP_MIN_COUNT] = std::min(w[QP_MIN_COUNT], st[r].count);
      w[QP_MIN_TIME] = std::min(w[QP_MIN_TIME], t);
      w[QP_NMAX_COUNT] = std::min(w[QP_NMAX_COUNT], ~st[r].count);
      w[QP_NMAX_TIME] = std::min(w[QP_NMAX_TIME], ~t);
      w[QP_AA_LO] += (hi * hi) & 0xFFFFFFFFull;
      w[QP_AA_HI] += (hi * hi) >> 32;
      w[QP_AB_LO] += (hi * lo) & 0xFFFFFFFFull;
      w[QP_AB_HI] += (hi * lo) >> 32;
      w[QP_BB_LO] += (lo * lo) & 0xFFFFFFFFull;
      w[QP_BB_HI] += (lo * lo) >> 32;
      w[QP_CC_LO] += (cn * cn) & 0xFFFFFFFFull;
      w[QP_CC_HI] += (cn * cn) >> 32;
      w[QP_LEVEL_SQ] += st[r].level * st[r].level;
      if (c == 0) {
        w[QP_NEVENTS] += b->g64[(size_t)G64_NEVENTS * b->rep_pad + r];
        w[QP_NFAULTED] += b->g32[(size_t)G32_STATUS * b->rep_pad + r] ? 1 : 0;
        w[QP_NREPS] += 1;
      }
    }
  }
  return QK_OK;
}

int qk_batch_reduce_stats(qk_batch* b, qk_comp_summary* out, uint32_t n_comp) {
  if (!b || !out || n_comp != b->low.hdr.n_comp) return QK_ERR_INVALID_ARG;
  for (uint32_t c = 0; c < n_comp; ++c) {
    qk_comp_summary& o = out[c];
    memset(&o, 0, sizeof(o));
    o.min_count = o.min_time = ~0ull;
    unsigned __int128 tsum = 0, csq = 0;
    /* sum of time_ns^2 in 32-bit digits (independent of the product's half-sum scheme) */
    unsigned __int128 dig[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<qk_comp_stats> st(b->cfg.n_reps);
    qk_batch_comp_stats(b, 0, b->cfg.n_reps, c, st.data());
    for (uint64_t r = 0; r < b->cfg.n_reps; ++r) {
      o.sum_count += st[r].count;
      tsum += st[r].time_ns;
      o.sum_level += st[r].level;
      o.min_count = std::min(o.min_count, st[r].count);
      o.max_count = std::max(o.max_count, st[r].count);
      o.min_time = std::min(o.min_time, st[r].time_ns);
      o.max_time = std::max(o.max_time, st[r].time_ns);
      const unsigned __int128 sq = (unsigned __int128)st[r].time_ns * st[r].time_ns;
      for (int k = 0; k < 4; ++k) dig[k] += (uint64_t)(sq >> (32 * k)) & 0xFFFFFFFFull;
      csq += (unsigned __int128)st[r].count * st[r].count;
      o.sumsq_level += st[r].level * st[r].level;
    }
    for (int k = 0; k < 7; ++k) {
      dig[k + 1] += dig[k] >> 32;
      dig[k] &= 0xFFFFFFFFull;
    }
    for (int k = 0; k < 3; ++k) o.sumsq_time_ns2[k] = (uint64_t)dig[2 * k] | ((uint64_t)dig[2 * k + 1] << 32);
    o.sumsq_count_lo = (uint64_t)csq;
    o.sumsq_count_hi = (uint64_t)(csq >> 64);
    o.sumsq_time = ((long double)o.sumsq_time_ns2[2] * 18446744073709551616.0L * 18446744073709551616.0L +
                    (long double)o.sumsq_time_ns2[1] * 18446744073709551616.0L + (long double)o.sumsq_time_ns2[0]) *
                   1e-18L;
    o.sumsq_count = (double)((long double)o.sumsq_count_hi * 18446744073709551616.0L + (long double)o.sumsq_count_lo);
    o.sum_time_lo = (uint64_t)tsum;
    o.sum_time_hi = (uint64_t)(tsum >> 64);
    o.hist_lo_time = o.min_time;
    o.hist_hi_time = o.max_time;
    o.hist_lo_count = o.min_count;
    o.hist_hi_count = o.max_count;
    const uint64_t w_t = (o.max_time - o.min_time) / QK_HIST_BINS + 1, w_c = (o.max_count - o.min_count) / QK_HIST_BINS + 1;
    for (uint64_t r = 0; r < b->cfg.n_reps; ++r) {
      o.hist_time[(st[r].time_ns - o.min_time) / w_t] += 1;
      o.hist_count[(st[r].count - o.min_count) / w_c] += 1;
    }
  }
  return QK_OK;
}
int qk_batch_reduce_partials_device(qk_batch*, uint64_t*, uint32_t) { return QK_ERR_NO_DEVICE; }
int qk_batch_combine_partials_device(qk_batch*, const uint64_t*, uint32_t, uint64_t*, uint32_t) { return QK_ERR_NO_DEVICE; }
int qk_batch_histograms_device(qk_batch*, const uint64_t*, uint64_t*, uint32_t) { return QK_ERR_NO_DEVICE; }

int qk_batch_read_log(qk_batch* b, uint64_t rep, qk_log_record* buf, uint64_t cap, uint64_t* n_out,
                      uint64_t* n_total) {
  if (!b || rep >= b->cfg.n_reps || !b->log) return QK_ERR_INVALID_ARG;
  const uint32_t cnt = b->g32[(size_t)G32_LOGCNT * b->rep_pad + rep];
  const uint32_t lc = b->cfg.log_capacity;
  const uint64_t kept = std::min<uint64_t>(cnt, lc);
  if (n_total) *n_total = cnt;
  if (n_out) *n_out = std::min<uint64_t>(kept, cap);
  if (!buf || cap == 0) return QK_OK;
  if (cap < kept) return QK_ERR_BUFFER_TOO_SMALL;
  const qk_log_record* ring = reinterpret_cast<const qk_log_record*>(b->logbuf.data()) + (size_t)rep * lc;
  const uint64_t first = cnt - kept;
  for (uint64_t i = 0; i < kept; ++i) {
    buf[i] = ring[(first + i) & (lc - 1)];
    if (buf[i].kind != QK_LOG_VEC_DATA) buf[i].aux = 0; /* continuation records carry fp64 bits there */
  }
  return QK_OK;
}

int qk_batch_read_logs(qk_batch* b, uint64_t rep0, uint64_t n, qk_log_record* buf, uint32_t* kept, uint64_t* total) {
  if (!b || rep0 + n > b->cfg.n_reps || !b->log || !buf || !kept || !total) return QK_ERR_INVALID_ARG;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t k = 0, t = 0;
    const int rc = qk_batch_read_log(b, rep0 + i, buf + i * b->cfg.log_capacity, b->cfg.log_capacity, &k, &t);
    if (rc) return rc;
    kept[i] = (uint32_t)k;
    total[i] = t;
  }
  return QK_OK;
}

int qk_batch_read_vector(qk_batch* b, uint64_t rep, uint32_t comp, double* out3) {
  if (!b || rep >= b->cfg.n_reps || comp >= b->low.hdr.n_comp || b->low.comps[comp].kind < QK_VECTOR_STOCK ||
      b->low.comps[comp].kind > QK_VECTOR_SPLITTER)
    return QK_ERR_INVALID_ARG;
  const DevVec& v = b->low.vecs[b->low.comps[comp].vec_idx];
  out3[0] = out3[1] = out3[2] = 0.0;
  for (uint32_t k = 0; k < v.dim && k < 3; ++k) memcpy(&out3[k], b->low.hdr.split ? &b->gc64[(size_t)rep * b->low.hdr.n64c + v.o64_res + k]
                                                                             : &b->g64[(size_t)(v.o64_res + k) * b->rep_pad + rep], 8);
  return QK_OK;
}

int qk_batch_read_vector_n(qk_batch* b, uint64_t rep, uint32_t comp, double* out, uint32_t cap, uint32_t* dim_out) {
  if (!b || rep >= b->cfg.n_reps || comp >= b->low.hdr.n_comp || b->low.comps[comp].kind < QK_VECTOR_STOCK ||
      b->low.comps[comp].kind > QK_VECTOR_SPLITTER || !out)
    return QK_ERR_INVALID_ARG;
  const DevVec& v = b->low.vecs[b->low.comps[comp].vec_idx];
  if (dim_out) *dim_out = v.dim;
  if (cap < v.dim) return QK_ERR_BUFFER_TOO_SMALL;
  for (uint32_t k = 0; k < v.dim; ++k) memcpy(&out[k], b->low.hdr.split ? &b->gc64[(size_t)rep * b->low.hdr.n64c + v.o64_res + k]
                                                                            : &b->g64[(size_t)(v.o64_res + k) * b->rep_pad + rep], 8);
  return QK_OK;
}

int qk_batch_read_stock(qk_batch* b, uint64_t rep, uint32_t comp, uint32_t* items, uint64_t cap, uint64_t* n_out) {
  if (!b || rep >= b->cfg.n_reps || comp >= b->low.hdr.n_comp || b->low.comps[comp].kind != QK_DISCRETE_STOCK)
    return QK_ERR_INVALID_ARG;
  const DevComp& c = b->low.comps[comp];
  const uint32_t hc = b->g32[(size_t)(c.o32 + S32_HC) * b->rep_pad + rep];
  const uint32_t head = QK_HC_HEAD(hc), cnt = QK_HC_CNT(hc);
  if (n_out) *n_out = cnt;
  if (!items) return QK_OK;
  if (cap < cnt) return QK_ERR_BUFFER_TOO_SMALL;
  for (uint32_t i = 0; i < cnt; ++i)
    items[i] = b->low.hdr.split ? b->gc32[(size_t)rep * b->low.hdr.n32c + c.ocold32 + (head + i) % c.ring_cap]
                                : b->gc32[(size_t)(c.ocold32 + (head + i) % c.ring_cap) * b->rep_pad + rep];
  return QK_OK;
}

int qk_batch_read_containers(qk_batch* b, uint64_t rep, uint32_t comp, uint32_t* handles, uint64_t* resources,
                             uint64_t cap, uint64_t* n_out) {
  if (!b || rep >= b->cfg.n_reps || comp >= b->low.hdr.n_comp || !(b->low.comps[comp].flags & DC_CONTAINER))
    return QK_ERR_INVALID_ARG;
  const DevComp& c = b->low.comps[comp];
  auto h32 = [&](uint32_t slot) { return b->g32[(size_t)slot * b->rep_pad + rep]; };
  const bool aos = b->low.hdr.split != 0; /* cold planes [replication][slot] */
  auto c32 = [&](uint32_t slot) {
    return aos ? b->gc32[(size_t)rep * b->low.hdr.n32c + slot] : b->gc32[(size_t)slot * b->rep_pad + rep];
  };
  auto c64 = [&](uint32_t slot) {
    return aos ? b->gc64[(size_t)rep * b->low.hdr.n64c + slot] : b->gc64[(size_t)slot * b->rep_pad + rep];
  };
  auto w32 = [&](uint32_t slot) { return aos ? c32(slot) : b->g32[(size_t)slot * b->rep_pad + rep]; };
  std::vector<uint32_t> hs, ps;
  if (c.kind == QK_DISCRETE_STOCK) {
    const uint32_t hc = h32(c.o32 + S32_HC), head = QK_HC_HEAD(hc), cnt = QK_HC_CNT(hc);
    for (uint32_t i = 0; i < cnt; ++i) {
      const uint32_t pos = (head + i) % c.ring_cap;
      hs.push_back(c32(c.ocold32 + pos));
      ps.push_back(c.ocold64 + 3u * pos);
    }
  } else if (c.kind == QK_DISCRETE_PARALLEL_PROCESS || c.kind == QK_CONTAINER_LOAD_PROCESS ||
             c.kind == QK_CONTAINER_UNLOAD_PROCESS) {
    const uint32_t nprog = w32(c.o32 + PP32_NPROG), cq = w32(c.o32 + PP32_COMPLETE);
    const uint32_t ring = c.ring_cap, item0 = c.o32 + PP32_BASE_COUNT;
    for (uint32_t i = 0; i < nprog; ++i) {
      hs.push_back(w32(item0 + i));
      ps.push_back(c.ocold64 + 3u * i);
    }
    for (uint32_t i = 0; i < (cq >> 16); ++i) {
      const uint32_t pos = ((cq & 0xFFFFu) + i) % ring;
      hs.push_back(w32(item0 + ring + pos));
      ps.push_back(c.ocold64 + 3u * (ring + pos));
    }
  } else if (w32(c.o32 + P32_FLAGS) & PF_HAS_ITEM) {
    hs.push_back(w32(c.o32 + P32_ITEM));
    ps.push_back(c.ocold64);
  }
  if (n_out) *n_out = hs.size();
  if (!handles || !resources) return QK_OK;
  if (cap < hs.size()) return QK_ERR_BUFFER_TOO_SMALL;
  for (size_t i = 0; i < hs.size(); ++i) {
    handles[i] = hs[i];
    for (uint32_t k = 0; k < 3; ++k) resources[i * 3 + k] = c64(ps[i] + k);
  }
  return QK_OK;
}

/* the multi-GPU plane needs GPUs */
int qk_multi_create(const qk_model*, const qk_multi_config*, qk_multi**) { return QK_ERR_NO_DEVICE; }
void qk_multi_destroy(qk_multi*) {}
int qk_multi_n_shards(qk_multi*, uint32_t*) { return QK_ERR_NO_DEVICE; }
int qk_multi_shard(qk_multi*, uint32_t, qk_batch**, uint64_t*, uint64_t*, int32_t*) { return QK_ERR_NO_DEVICE; }
int qk_multi_set_tape(qk_multi*, uint32_t, const double*, uint64_t) { return QK_ERR_NO_DEVICE; }
int qk_multi_init(qk_multi*, uint64_t) { return QK_ERR_NO_DEVICE; }
int qk_multi_step_until(qk_multi*, uint64_t) { return QK_ERR_NO_DEVICE; }
int qk_multi_step_events(qk_multi*, uint64_t) { return QK_ERR_NO_DEVICE; }
int qk_multi_synchronize(qk_multi*) { return QK_ERR_NO_DEVICE; }
int qk_multi_last_step_ms(qk_multi*, float*) { return QK_ERR_NO_DEVICE; }
int qk_multi_stats(qk_multi*, qk_comp_summary*, uint32_t, qk_batch_summary*) { return QK_ERR_NO_DEVICE; }

} /* extern "C" */
