/*
 * qk_oracle.cpp -- CPU ORACLE for the QuokkaSim batched-replication hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see qk_oracle.h header comment).  One replication, one thread,
 * unbounded containers, written to follow the reference line by line; every function cites
 * the reference file:line (relative to /root/reference/quokkasim/src) it restates.
 * Deliberately shares NO code with quokkasim_b200/csrc/ so that agreement between the two
 * is evidence, not tautology.
 *
 * Scheduler semantics restate NeXosim 0.3.3 (un-vendored) under the canonical
 * order (time, origin_rank, insertion_seq); origin_rank = registration index + 1, external
 * scheduler events use rank 0.
 *
 * Build: see oracle/Makefile (g++ -O2 -ffp-contract=off).
 */
#include "qk_oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <deque>
#include <map>
#include <queue>
#include <thread>
#include <vector>

namespace {

typedef unsigned __int128 u128;

struct Fault {
  uint32_t code;
};

/* ------------------------------------------------------------------------------------------
 * Duration::from_secs_f64 (Rust std core::time, macro try_from_secs!, f64 instantiation:
 * mantissa_bits=52, exponent_bits=11, offset=44).  Call sites: discrete.rs:490,492,860,1107,1380;
 * delays.rs:102,132,163.  Round to nearest ns, ties to even, from the exact binary value.
 * ---------------------------------------------------------------------------------------- */
int from_secs_f64(double secs, uint64_t* out_ns) {
  if (secs < 0.0) return ORC_FAULT_BAD_DURATION; /* TryFromFloatSecsErrorKind::Negative */
  uint64_t bits;
  std::memcpy(&bits, &secs, 8);
  const uint64_t MANT_MASK = (1ull << 52) - 1;
  const uint64_t EXP_MASK = (1ull << 11) - 1;
  const int MIN_EXP = 1 - (1 << 11) / 2; /* -1023 */
  uint64_t mant = (bits & MANT_MASK) | (MANT_MASK + 1);
  int exp = (int)((bits >> 52) & EXP_MASK) + MIN_EXP;
  const uint64_t NANOS_PER_SEC = 1000000000ull;
  uint64_t s, n;
  if (exp < -31) {
    s = 0;
    n = 0; /* "the input represents less than 1ns and can not be rounded to it" */
  } else if (exp < 0) {
    u128 t = (u128)mant << (44 + exp);
    const int nanos_offset = 52 + 44;
    u128 nanos_tmp = (u128)NANOS_PER_SEC * t;
    uint64_t nanos = (uint64_t)(nanos_tmp >> nanos_offset);
    u128 rem_mask = ((u128)1 << nanos_offset) - 1;
    u128 rem_msb_mask = (u128)1 << (nanos_offset - 1);
    u128 rem = nanos_tmp & rem_mask;
    bool is_tie = rem == rem_msb_mask;
    bool is_even = (nanos & 1) == 0;
    bool rem_msb = (nanos_tmp & rem_msb_mask) == 0;
    bool add_ns = !(rem_msb || (is_even && is_tie));
    nanos += add_ns ? 1 : 0;
    if (nanos != NANOS_PER_SEC) {
      s = 0;
      n = nanos;
    } else {
      s = 1;
      n = 0;
    }
  } else if (exp < 52) {
    s = mant >> (52 - exp);
    u128 t = (u128)((mant << exp) & MANT_MASK);
    const int nanos_offset = 52;
    u128 nanos_tmp = (u128)NANOS_PER_SEC * t;
    uint64_t nanos = (uint64_t)(nanos_tmp >> nanos_offset);
    u128 rem_mask = ((u128)1 << nanos_offset) - 1;
    u128 rem_msb_mask = (u128)1 << (nanos_offset - 1);
    u128 rem = nanos_tmp & rem_mask;
    bool is_tie = rem == rem_msb_mask;
    bool is_even = (nanos & 1) == 0;
    bool rem_msb = (nanos_tmp & rem_msb_mask) == 0;
    bool add_ns = !(rem_msb || (is_even && is_tie));
    nanos += add_ns ? 1 : 0;
    if (nanos != NANOS_PER_SEC) {
      n = nanos;
    } else {
      s += 1;
      n = 0;
    }
  } else if (exp < 64) {
    s = mant << (exp - 52);
    n = 0;
  } else {
    return ORC_FAULT_BAD_DURATION; /* OverflowOrNan */
  }
  /* the executor's time base is u64 ns: durations of 2^34 s (544 years) and more are a fault */
  if (exp >= 34) return ORC_FAULT_BAD_DURATION;
  *out_ns = s * NANOS_PER_SEC + n;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11; Random123).  Native stream for batched replications:
 * key = (lo32(base_seed), hi32(base_seed)); counter = (lo32(rep), hi32(rep), stream, draw).
 * ---------------------------------------------------------------------------------------- */
void philox4x32_10(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
  uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
  uint32_t k0 = k_in[0], k1 = k_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

/* deterministic natural log built from + - * / only (identical operation order is specified
 * in DESIGN.md "Native sampling arithmetic" so that CPU and GPU agree bit for bit). */
double det_ln(double x) {
  int e;
  double m = std::frexp(x, &e); /* m in [0.5,1) */
  if (m < 0.70710678118654752440) {
    m = m * 2.0;
    e -= 1;
  }
  double s = (m - 1.0) / (m + 1.0);
  double z = s * s;
  double p = 1.0 / 27.0;
  p = p * z + 1.0 / 25.0;
  p = p * z + 1.0 / 23.0;
  p = p * z + 1.0 / 21.0;
  p = p * z + 1.0 / 19.0;
  p = p * z + 1.0 / 17.0;
  p = p * z + 1.0 / 15.0;
  p = p * z + 1.0 / 13.0;
  p = p * z + 1.0 / 11.0;
  p = p * z + 1.0 / 9.0;
  p = p * z + 1.0 / 7.0;
  p = p * z + 1.0 / 5.0;
  p = p * z + 1.0 / 3.0;
  p = p * z + 1.0;
  double r = 2.0 * s;
  r = r * p;
  double el = (double)e * 0.69314718055994530942;
  return r + el;
}

inline double u01_from(uint32_t lo, uint32_t hi) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 11) * (1.0 / 9007199254740992.0); /* 2^-53, as rand's Standard f64 */
}

/* Distribution::sample (common.rs:152-178) with the RNG replaced by Philox.
 * Triangular follows rand_distr 0.4.3 Triangular::sample; Uniform is low + u*(high-low);
 * Exponential is inverse-CDF; Normal is Marsaglia polar (rand_distr uses ziggurat: native mode
 * is distribution-equivalent, not stream-equivalent, to the reference -- that is what the tape is for). */
double sample_native(const orc_dist& d, uint64_t seed, uint64_t rep, uint32_t& draws) {
  if (d.kind == ORC_DIST_CONSTANT) return d.p[0];
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (;;) {
    uint32_t ctr[4] = {(uint32_t)rep, (uint32_t)(rep >> 32), d.stream, draws};
    uint32_t x[4];
    philox4x32_10(ctr, key, x);
    draws += 1;
    double u = u01_from(x[0], x[1]);
    switch (d.kind) {
      case ORC_DIST_UNIFORM: {
        double w = d.p[1] - d.p[0];
        double t = u * w;
        return d.p[0] + t;
      }
      case ORC_DIST_TRIANGULAR: {
        double mn = d.p[0], mx = d.p[1], mode = d.p[2];
        double diff_mode_min = mode - mn;
        double range = mx - mn;
        double f_range = u * range;
        if (f_range < diff_mode_min) {
          double a = f_range * diff_mode_min;
          return mn + std::sqrt(a);
        } else {
          double a = range - f_range;
          double b = mx - mode;
          double c = a * b;
          return mx - std::sqrt(c);
        }
      }
      case ORC_DIST_EXPONENTIAL: {
        double om = 1.0 - u;
        double l = det_ln(om);
        double nl = 0.0 - l;
        return nl * d.p[0];
      }
      case ORC_DIST_NORMAL:
      case ORC_DIST_TRUNCNORMAL: {
        double u2 = u01_from(x[2], x[3]);
        double v1 = 2.0 * u - 1.0;
        double v2 = 2.0 * u2 - 1.0;
        double a = v1 * v1;
        double b = v2 * v2;
        double s = a + b;
        if (s >= 1.0 || s == 0.0) continue;
        double l = det_ln(s);
        double m2 = -2.0 * l;
        double q = m2 / s;
        double f = std::sqrt(q);
        double zn = v1 * f;
        double sc = d.p[1] * zn;
        double xv = d.p[0] + sc;
        if (d.kind == ORC_DIST_TRUNCNORMAL) {
          if (!(xv >= d.p[2] && xv <= d.p[3])) continue; /* common.rs:166-173 rejection loop */
        }
        return xv;
      }
      default:
        return d.p[0];
    }
  }
}

/* ========================================================================================== */

struct EvId {
  uint16_t comp;
  uint32_t seq;
};
const EvId EV_INIT = {ORC_SRC_INIT, 0};
const EvId EV_SCH = {ORC_SRC_SCH, 0};

struct DelayModeDef {
  int until_delay; /* distribution instance ids */
  int until_fix;
};

struct CompDef {
  uint32_t kind;
  uint32_t low, max; /* discrete stock capacities (discrete.rs:50-51) */
  uint32_t n_initial;
  uint32_t initial_base; /* first handle index of the initial items */
  int dist_time, dist_qty;
  std::vector<DelayModeDef> delay_modes;
  int upstream, downstream; /* stock ids or -1 (single connection supported) */
  int env;                  /* environment id or -1 */
  std::vector<int> subscribers; /* stock.state_emitter / env.emit_change targets, connection order */
  int source_ordinal;
  /* vector */
  uint32_t dim;
  double vlow, vmax;
  double vinit[8];
  uint32_t tmask; /* fields counted by ResourceTotal (set_vector_n) */
  int ups[5], downs[5];
  uint32_t n_ports;
  double ratios[5];
  /* container family (vector_container.rs): factory of a container source; initial contents of a container stock */
  int dist_cqty; /* create_quantity_distr instance or -1 (None) */
  double csplit[3], ccapacity;
  std::vector<std::array<double, 3> > init_res;
  std::vector<uint8_t> init_has;
  std::vector<double> init_cap;
  CompDef()
      : kind(0), low(0), max(1), n_initial(0), initial_base(0), dist_time(-1), dist_qty(-1), upstream(-1),
        downstream(-1), env(-1), source_ordinal(-1), dim(0), vlow(0), vmax(0), n_ports(0), dist_cqty(-1),
        ccapacity(0) {
    csplit[0] = csplit[1] = csplit[2] = 0;
    for (int i = 0; i < 5; ++i) {
      ups[i] = downs[i] = -1;
      ratios[i] = 0;
    }
    for (int k = 0; k < 8; ++k) vinit[k] = 0;
    tmask = 7;
  }
};

struct ExtEvent {
  uint64_t t;
  uint32_t type; /* 0 = SetEnvironmentState, 1 = SetLowCapacity(F64Stock), 2 = SetProcessQuantity(F64Process): `state` =
                    the distribution instance df.create made at schedule time (core.rs:1388) */
  uint32_t target;
  uint32_t state;
  double value;
};

}  // namespace

struct orc_model {
  std::vector<CompDef> comps;
  std::vector<orc_dist> dists;
  std::vector<ExtEvent> ext;
  uint32_t n_sources;
  uint32_t n_initial_total;
  /* container capacities: immutable per container, so they are a function of the handle --
   * by source ordinal (factory capacity) or by initial-item index */
  double src_caps[64];
  std::vector<double> init_caps;
  orc_model() : n_sources(0), n_initial_total(0) {
    for (int i = 0; i < 64; ++i) src_caps[i] = 0;
  }
};

namespace {

inline bool is_process_like(uint32_t k) {
  return k == ORC_DISCRETE_PROCESS || k == ORC_DISCRETE_SOURCE || k == ORC_DISCRETE_SINK ||
         k == ORC_DISCRETE_PARALLEL_PROCESS || k == ORC_VECTOR_PROCESS || k == ORC_VECTOR_SOURCE ||
         k == ORC_VECTOR_SINK || k == ORC_VECTOR_COMBINER || k == ORC_VECTOR_SPLITTER ||
         (k >= ORC_CONTAINER_PROCESS && k <= ORC_CONTAINER_UNLOAD_PROCESS);
}
/* DiscreteStock<T> for T = String or Vector3Container (core.rs:541,549) */
inline bool is_item_stock(uint32_t k) { return k == ORC_DISCRETE_STOCK || k == ORC_CONTAINER_STOCK; }
inline bool is_container_kind(uint32_t k) { return k >= ORC_CONTAINER_STOCK && k <= ORC_CONTAINER_UNLOAD_PROCESS; }
inline bool is_multi_server(uint32_t k) {
  return k == ORC_DISCRETE_PARALLEL_PROCESS || k == ORC_CONTAINER_PARALLEL_PROCESS ||
         k == ORC_CONTAINER_LOAD_PROCESS || k == ORC_CONTAINER_UNLOAD_PROCESS;
}

enum { CAT_EMPTY = 0, CAT_NORMAL = 1, CAT_FULL = 2, CAT_NONE = 3 };

struct StockState {
  int cat;
  uint32_t occupied, empty;
};

struct DelayModeState {
  bool in_fix; /* DelayState::TimeUntilFix vs TimeUntilDelay (delays.rs:13-17) */
  uint64_t dur;
};

struct Comp {
  /* process-like runtime state (discrete.rs:345-354) */
  bool has_item;
  uint64_t left;
  uint32_t item;
  bool env_stopped;
  bool ttnpe_some;
  uint64_t ttnpe;
  bool sched_some;
  uint64_t sched_time;
  uint64_t sched_key;
  uint32_t next_event_index;
  uint64_t previous_check_time;
  uint64_t next_item_index; /* StringItemFactory.next_index (discrete.rs:666) */
  std::vector<DelayModeState> dm;
  /* parallel process (discrete.rs:1214-1216) */
  std::vector<std::pair<uint64_t, uint32_t> > in_progress;
  std::deque<uint32_t> complete;
  /* stock (discrete.rs:54) */
  std::deque<uint32_t> items;
  /* environment */
  uint32_t env_state;
  /* vector components (vector.rs): stock resource / process in-flight vector; combiner keeps M vectors */
  double vres[8];
  std::vector<std::array<double, 8> > parts;
  double vlow; /* low_capacity is mutable at run time (SetLowCapacity, core.rs:1382-1386) */
  int qty_dist; /* process_quantity_distr is too (SetProcessQuantity, core.rs:1387-1391): -1 = the model's */
  bool ttnde_some;
  uint64_t ttnde;
  /* stats */
  orc_comp_stats st;
  uint64_t area_last_t;
  Comp()
      : has_item(false), left(0), item(0), env_stopped(false), ttnpe_some(false), ttnpe(0), sched_some(false),
        sched_time(0), sched_key(0), next_event_index(0), previous_check_time(0), next_item_index(0), env_state(0),
        vlow(0), qty_dist(-1), ttnde_some(false), ttnde(0), area_last_t(0) {
    std::memset(&st, 0, sizeof(st));
    for (int k = 0; k < 8; ++k) vres[k] = 0;
  }
};

enum { ACT_KEYED = 0, ACT_EMIT = 1, ACT_EXT = 2 };

struct Action {
  uint64_t time;
  uint32_t rank;
  uint64_t seq;
  uint32_t type;
  uint32_t target;
  EvId src;
  uint64_t key; /* for keyed events */
  uint32_t ext_index;
  int pay_cat;    /* VectorStock::emit_change carries the state computed when it was scheduled (vector.rs:130,158) */
  double pay_occ;
};
struct ActionCmp {
  bool operator()(const Action& a, const Action& b) const {
    if (a.time != b.time) return a.time > b.time;
    if (a.rank != b.rank) return a.rank > b.rank;
    return a.seq > b.seq;
  }
};

}  // namespace

struct orc_sim {
  const orc_model* m;
  uint64_t seed, rep, t0, now;
  uint32_t status;
  uint64_t n_events;
  bool log_enabled;
  std::vector<Comp> c;
  std::vector<uint32_t> draws;
  std::vector<const double*> tape;
  std::vector<std::vector<double> > tape_store;
  std::priority_queue<Action, std::vector<Action>, ActionCmp> q;
  std::vector<uint8_t> cancelled; /* indexed by key */
  uint64_t next_seq;
  int pay_cat_next = 0;
  double pay_occ_next = 0;
  std::vector<orc_log_record> logs;
  /* container family: a container IS its handle; what it carries lives here (Vector3Container.resource) */
  struct CPay {
    bool has;
    double v[3];
  };
  std::map<uint32_t, CPay> cpay;
  double container_capacity(uint32_t h) const {
    return (h >> 26) == 63u ? m->init_caps[h & 0x3FFFFFFu] : m->src_caps[h >> 26];
  }
  void impl_load(uint32_t, EvId*);   /* oracle/qk_oracle_container.inc */
  void impl_unload(uint32_t, EvId*);
  void create_container(const CompDef&, uint32_t);

  /* ---- sampling -------------------------------------------------------------------------- */
  double sample(int dist_id) {
    const orc_dist& d = m->dists[dist_id];
    if (d.kind == ORC_DIST_CONSTANT) return d.p[0];
    if (!tape_store[dist_id].empty() || tape[dist_id]) {
      const std::vector<double>& t = tape_store[dist_id];
      if (draws[dist_id] >= t.size()) throw Fault{ORC_FAULT_TAPE_EXHAUSTED};
      return t[draws[dist_id]++];
    }
    return sample_native(d, seed, rep, draws[dist_id]);
  }
  uint64_t dur_from(double secs) {
    uint64_t ns;
    int rc = from_secs_f64(secs, &ns);
    if (rc) throw Fault{(uint32_t)rc};
    return ns;
  }

  /* ---- logging: `log()` of every component (discrete.rs:235-251, 566-582; core.rs:273-289) ---- */
  EvId log(uint32_t comp, EvId src, uint8_t kind, uint8_t reason, uint64_t payload) {
    Comp& s = c[comp];
    EvId id = {(uint16_t)comp, s.next_event_index};
    if (log_enabled) {
      orc_log_record r;
      r.time_ns = now;
      r.comp = (uint16_t)comp;
      r.kind = kind;
      r.reason = reason;
      r.seq = s.next_event_index;
      r.src_comp = src.comp;
      r.aux = 0;
      r.src_seq = src.seq;
      r.payload = payload;
      logs.push_back(r);
      /* records that serialise an item (discrete.rs:299-316, 612-631): a container's JSON carries its resource, so
       * its vector follows as one continuation record (format: qk_oracle_vector.inc header) */
      if (is_container_kind(m->comps[comp].kind) &&
          (kind == ORC_LOG_PROCESS_START || kind == ORC_LOG_PROCESS_FINISH || kind == ORC_LOG_STOCK_ADD ||
           (kind == ORC_LOG_STOCK_REMOVE && payload != ORC_ITEM_NONE))) {
        const CPay& cp = cpay[(uint32_t)payload];
        uint64_t w[3];
        for (int k = 0; k < 3; ++k) std::memcpy(&w[k], &cp.v[k], 8);
        if (!cp.has) {
          w[0] = ORC_CONTAINER_NONE_BITS;
          w[1] = w[2] = 0;
        }
        orc_log_record q;
        std::memset(&q, 0, sizeof(q));
        q.time_ns = w[0];
        q.comp = (uint16_t)comp;
        q.kind = ORC_LOG_VEC_DATA;
        q.reason = 0;
        q.seq = id.seq;
        std::memcpy(&q.src_comp, &w[1], 8);
        q.payload = w[2];
        logs.push_back(q);
      }
    }
    s.next_event_index += 1;
    return id;
  }

  /* ---- scheduler (NeXosim restated) ---------------------------------------------------------- */
  void schedule(uint64_t time, uint32_t rank, uint32_t type, uint32_t target, EvId src, uint64_t key,
                uint32_t ext_index) {
    Action a;
    a.time = time;
    a.rank = rank;
    a.seq = next_seq++;
    a.type = type;
    a.target = target;
    a.src = src;
    a.key = key;
    a.ext_index = ext_index;
    a.pay_cat = pay_cat_next;
    a.pay_occ = pay_occ_next;
    q.push(a);
  }
  uint64_t schedule_keyed(uint64_t time, uint32_t comp, EvId src) {
    uint64_t key = cancelled.size();
    cancelled.push_back(0);
    schedule(time, comp + 1, ACT_KEYED, comp, src, key, 0);
    return key;
  }

  /* ---- DiscreteStock (discrete.rs:158-252) --------------------------------------------------- */
  StockState stock_state(uint32_t sidx) { /* get_state discrete.rs:161-171 */
    const CompDef& d = m->comps[sidx];
    Comp& s = c[sidx];
    StockState r;
    r.occupied = (uint32_t)s.items.size();
    r.empty = d.max > r.occupied ? d.max - r.occupied : 0; /* saturating_sub */
    if (r.occupied <= d.low)
      r.cat = CAT_EMPTY;
    else if (r.occupied >= d.max)
      r.cat = CAT_FULL;
    else
      r.cat = CAT_NORMAL;
    return r;
  }
  void stock_area(uint32_t sidx) {
    Comp& s = c[sidx];
    s.st.t_a_ns += (uint64_t)s.items.size() * (now - s.area_last_t);
    s.area_last_t = now;
  }
  void stock_add(uint32_t sidx, uint32_t item, EvId src) { /* Stock::add core.rs:177-183 */
    Comp& s = c[sidx];
    StockState prev = stock_state(sidx); /* pre_add: set_previous_state core.rs:163-167 */
    stock_area(sidx);
    s.items.push_back(item); /* add_impl discrete.rs:181-186 */
    s.st.n_a += 1;
    if (s.items.size() > s.st.t_b_ns) s.st.t_b_ns = s.items.size();
    EvId id = log(sidx, src, ORC_LOG_STOCK_ADD, 0, item);
    StockState cur = stock_state(sidx); /* post_add discrete.rs:187-198 */
    if (prev.cat != cur.cat) schedule(now + 1, sidx + 1, ACT_EMIT, sidx, id, 0, 0);
  }
  bool stock_remove(uint32_t sidx, EvId src, uint32_t* item_out) { /* Stock::remove core.rs:197-204 */
    Comp& s = c[sidx];
    StockState prev = stock_state(sidx); /* pre_remove + remove_impl discrete.rs:202 */
    stock_area(sidx);
    bool some = !s.items.empty();
    uint32_t item = 0;
    if (some) {
      item = s.items.front();
      s.items.pop_front();
      s.st.n_b += 1;
    }
    EvId id = log(sidx, src, ORC_LOG_STOCK_REMOVE, 0, some ? (uint64_t)item : ORC_ITEM_NONE);
    StockState cur = stock_state(sidx); /* post_remove discrete.rs:210-225 */
    if (prev.cat != cur.cat) schedule(now + 1, sidx + 1, ACT_EMIT, sidx, id, 0, 0);
    *item_out = item;
    return some;
  }
  void stock_emit_change(uint32_t sidx, EvId src) { /* discrete.rs:227-233 */
    StockState st = stock_state(sidx);              /* state at emission time, not the payload */
    EvId nm = log(sidx, src, ORC_LOG_STOCK_STATE_CHANGE, (uint8_t)st.cat,
                  ((uint64_t)st.occupied << 32) | st.empty);
    const std::vector<int>& subs = m->comps[sidx].subscribers;
    for (size_t i = 0; i < subs.size(); ++i) {
      update_state(subs[i], nm);
      if (status) return;
    }
  }

  /* ---- DelayModes (delays.rs:67-175) --------------------------------------------------------- */
  int dm_active(const Comp& p) const { /* active_delay delays.rs:77-84 */
    for (size_t i = 0; i < p.dm.size(); ++i)
      if (p.dm[i].in_fix) return (int)i;
    return -1;
  }
  void dm_update_state(uint32_t pidx, uint64_t dt, int* from, int* to) { /* delays.rs:86-142 */
    Comp& p = c[pidx];
    const CompDef& d = m->comps[pidx];
    *from = -1;
    *to = -1;
    int a = dm_active(p);
    if (a >= 0) {
      uint64_t applied = std::min(dt, p.dm[a].dur);
      p.st.t_b_ns += applied;
      p.dm[a].dur -= applied; /* saturating_sub */
      if (p.dm[a].dur == 0) {
        double secs = sample(d.delay_modes[a].until_delay);
        p.dm[a].in_fix = false;
        p.dm[a].dur = dur_from(secs);
      }
      *from = a;
    } else {
      for (size_t i = 0; i < p.dm.size(); ++i)
        if (!p.dm[i].in_fix) p.dm[i].dur -= std::min(dt, p.dm[i].dur);
    }
    int a2 = dm_active(p);
    if (a2 >= 0) {
      *to = a2;
    } else {
      for (size_t i = 0; i < p.dm.size(); ++i) {
        if (!p.dm[i].in_fix && p.dm[i].dur == 0) {
          double secs = sample(d.delay_modes[i].until_fix);
          p.dm[i].in_fix = true;
          p.dm[i].dur = dur_from(secs);
          *to = (int)i;
          break;
        }
      }
    }
  }
  int dm_next_event(const Comp& p, uint64_t* dur) const { /* get_next_event delays.rs:144-155 */
    int a = dm_active(p);
    if (a >= 0) {
      *dur = p.dm[a].dur;
      return a;
    }
    int best = -1;
    for (size_t i = 0; i < p.dm.size(); ++i)
      if (!p.dm[i].in_fix && (best < 0 || p.dm[i].dur < p.dm[best].dur)) best = (int)i;
    if (best >= 0) *dur = p.dm[best].dur;
    return best;
  }

  /* the delay tick + DelayEnd/DelayStart logs shared by all process-like components
   * (discrete.rs:441-451, 816-826, 1062-1072, 1333-1343) */
  void tick_delays(uint32_t pidx, uint64_t dt, EvId* src) {
    int from, to;
    dm_update_state(pidx, dt, &from, &to);
    bool changed = (from != to); /* has_changed delays.rs:35-42 */
    if (changed) {
      if (from >= 0) *src = log(pidx, *src, ORC_LOG_DELAY_END, 0, (uint64_t)from);
      if (to >= 0) *src = log(pidx, *src, ORC_LOG_DELAY_START, 0, (uint64_t)to);
    }
  }
  /* "Update cached environment state" block (discrete.rs:455-471 and siblings) */
  void refresh_env(uint32_t pidx, EvId* src) {
    Comp& p = c[pidx];
    const CompDef& d = m->comps[pidx];
    bool new_stopped = d.env >= 0 ? (c[d.env].env_state == 1) : false;
    if (!p.env_stopped && new_stopped) {
      *src = log(pidx, *src, ORC_LOG_PROCESS_STOPPED, ORC_REASON_STOPPED_BY_ENV, 0);
      p.env_stopped = true;
    } else if (p.env_stopped && !new_stopped) {
      *src = log(pidx, *src, ORC_LOG_PROCESS_CONTINUE, ORC_REASON_RESUMED_BY_ENV, 0);
      p.env_stopped = false;
    }
  }

  /* ---- Process::update_state = pre -> impl -> post (core.rs:370-376) -------------------------- */
  void update_state(uint32_t pidx, EvId src) {
    const CompDef& d = m->comps[pidx];
    Comp& p = c[pidx];
    /* pre_update_state (discrete.rs:404-412): forget (do NOT cancel) a handle whose time has come */
    if (p.sched_some && p.sched_time <= now) p.sched_some = false;
    switch (d.kind) {
      /* Vector3ContainerProcess/Source/Sink/ParallelProcess ARE DiscreteProcess<.., Vector3Container> etc.
       * (core.rs:550-553): same code, the item is a container */
      case ORC_DISCRETE_PROCESS:
      case ORC_CONTAINER_PROCESS:
        impl_process(pidx, &src);
        break;
      case ORC_DISCRETE_SOURCE:
      case ORC_CONTAINER_SOURCE:
        impl_source(pidx, &src);
        break;
      case ORC_DISCRETE_SINK:
      case ORC_CONTAINER_SINK:
        impl_sink(pidx, &src);
        break;
      case ORC_DISCRETE_PARALLEL_PROCESS:
      case ORC_CONTAINER_PARALLEL_PROCESS:
        impl_parallel(pidx, &src);
        break;
      case ORC_CONTAINER_LOAD_PROCESS:
        impl_load(pidx, &src);
        break;
      case ORC_CONTAINER_UNLOAD_PROCESS:
        impl_unload(pidx, &src);
        break;
      default:
        impl_vector(pidx, &src);
        return; /* vector components have their own post rules */
    }
    post_discrete(pidx, src);
  }

  /* post_update_state (discrete.rs:535-564, 891-920, 1138-1167, 1429-1458) */
  void post_discrete(uint32_t pidx, EvId src) {
    Comp& p = c[pidx];
    if (p.ttnpe_some) {
      if (p.ttnpe == 0) throw Fault{ORC_FAULT_ZERO_DURATION};
      uint64_t next_time = now + p.ttnpe;
      if (p.sched_some) {
        if (next_time < p.sched_time) {
          cancelled[p.sched_key] = 1; /* action_key.cancel() */
          p.sched_key = schedule_keyed(next_time, pidx, src);
          p.sched_time = next_time;
        }
      } else {
        p.sched_key = schedule_keyed(next_time, pidx, src);
        p.sched_time = next_time;
        p.sched_some = true;
      }
    }
    p.previous_check_time = now;
  }

  /* DiscreteProcess::update_state_impl (discrete.rs:414-533) */
  void impl_process(uint32_t pidx, EvId* src) {
    const CompDef& d = m->comps[pidx];
    Comp& p = c[pidx];
    uint64_t dt = now - p.previous_check_time;
    {
      bool is_in_delay = dm_active(p) >= 0;
      bool is_in_process = p.has_item && !is_in_delay;
      bool is_env_blocked = p.env_stopped;
      if (!(is_in_delay || is_env_blocked)) {
        if (p.has_item) {
          uint64_t applied = std::min(dt, p.left);
          p.st.t_a_ns += applied;
          p.left -= applied;
          if (p.left == 0) {
            p.has_item = false;
            *src = log(pidx, *src, ORC_LOG_PROCESS_FINISH, 0, p.item);
            p.st.n_b += 1;
            if (d.downstream >= 0) stock_add(d.downstream, p.item, *src); /* no downstream check at finish */
          }
        }
      }
      if (!is_env_blocked && (is_in_delay || is_in_process)) tick_delays(pidx, dt, src);
    }
    refresh_env(pidx, src);
    bool is_env_stopped = p.env_stopped;
    bool has_active_delay = dm_active(p) >= 0 || is_env_stopped;
    if (!p.has_item && !has_active_delay) {
      bool us_some = d.upstream >= 0, ds_some = d.downstream >= 0;
      StockState us = {CAT_NONE, 0, 0}, ds = {CAT_NONE, 0, 0};
      if (us_some) us = stock_state(d.upstream);
      if (ds_some) ds = stock_state(d.downstream);
      if (us_some && (us.cat == CAT_NORMAL || us.cat == CAT_FULL) && ds_some &&
          (ds.cat == CAT_EMPTY || ds.cat == CAT_NORMAL)) {
        *src = log(pidx, *src, ORC_LOG_WITHDRAW_REQUEST, 0, 0);
        uint32_t item;
        bool got = stock_remove(d.upstream, *src, &item);
        if (got) {
          double secs = sample(d.dist_time);
          p.left = dur_from(secs);
          p.item = item;
          p.has_item = true;
          *src = log(pidx, *src, ORC_LOG_PROCESS_START, 0, item);
          p.st.n_a += 1;
          p.ttnpe_some = true;
          p.ttnpe = p.left;
        } else {
          *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_UPSTREAM_NO_RESOURCE, 0);
          p.ttnpe_some = false;
        }
      } else if (us_some && us.cat == CAT_EMPTY) {
        *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_UPSTREAM_EMPTY, 0);
        p.ttnpe_some = false;
      } else if (!us_some) {
        *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_UPSTREAM_NOT_CONNECTED, 0);
        p.ttnpe_some = false;
      } else if (ds_some && ds.cat == CAT_FULL) {
        *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_DOWNSTREAM_FULL, 0);
        p.ttnpe_some = false;
      } else { /* (_, None) */
        *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_DOWNSTREAM_NOT_CONNECTED, 0);
        p.ttnpe_some = false;
      }
    } else if (p.has_item && !has_active_delay) {
      p.ttnpe_some = true;
      p.ttnpe = p.left;
    } else { /* (_, true) */
      int a = dm_active(p);
      p.ttnpe_some = a >= 0;
      if (a >= 0) p.ttnpe = p.dm[a].dur;
    }
    /* time_to_next_delay_event (discrete.rs:527-531) is computed but never read by post (quirk Q1) */
  }

  /* DiscreteSource::update_state_impl (discrete.rs:789-889) */
  void impl_source(uint32_t pidx, EvId* src) {
    const CompDef& d = m->comps[pidx];
    Comp& p = c[pidx];
    uint64_t dt = now - p.previous_check_time;
    {
      bool is_in_delay = dm_active(p) >= 0;
      bool is_in_process = p.has_item && !is_in_delay;
      bool is_env_blocked = p.env_stopped;
      if (!(is_in_delay || is_env_blocked)) {
        if (p.has_item) {
          uint64_t applied = std::min(dt, p.left);
          p.st.t_a_ns += applied;
          p.left -= applied;
          if (p.left == 0) {
            p.has_item = false;
            *src = log(pidx, *src, ORC_LOG_PROCESS_FINISH, 0, p.item);
            p.st.n_b += 1;
            if (d.downstream >= 0) stock_add(d.downstream, p.item, *src);
          }
        }
      }
      if (!is_env_blocked && (is_in_delay || is_in_process)) tick_delays(pidx, dt, src);
    }
    refresh_env(pidx, src);
    bool is_env_stopped = p.env_stopped;
    bool has_active_delay = dm_active(p) >= 0 || is_env_stopped;
    if (!p.has_item && !has_active_delay) {
      if (d.downstream >= 0) {
        StockState ds = stock_state(d.downstream);
        if (ds.cat == CAT_EMPTY || ds.cat == CAT_NORMAL) {
          double secs = sample(d.dist_time);
          /* ItemFactory::create_item (discrete.rs:680-686): handle = (source ordinal, next_index) */
          uint32_t item = ((uint32_t)d.source_ordinal << 26) | (uint32_t)(p.next_item_index & 0x3FFFFFFu);
          p.next_item_index += 1;
          if (d.kind == ORC_CONTAINER_SOURCE) create_container(d, item); /* Vector3ContainerFactory::create_item */
          p.left = dur_from(secs);
          p.item = item;
          p.has_item = true;
          *src = log(pidx, *src, ORC_LOG_PROCESS_START, 0, item);
          p.st.n_a += 1;
          p.ttnpe_some = true;
          p.ttnpe = p.left;
        } else {
          *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_DOWNSTREAM_FULL, 0);
          p.ttnpe_some = false;
        }
      } else {
        *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_DOWNSTREAM_NOT_CONNECTED, 0);
        p.ttnpe_some = false;
      }
    } else if (p.has_item) { /* (Some, _) precedes (_, true): discrete.rs:874-876 */
      p.ttnpe_some = true;
      p.ttnpe = p.left;
    } else {
      int a = dm_active(p);
      p.ttnpe_some = a >= 0;
      if (a >= 0) p.ttnpe = p.dm[a].dur;
    }
  }

  /* DiscreteSink::update_state_impl (discrete.rs:1035-1136) */
  void impl_sink(uint32_t pidx, EvId* src) {
    const CompDef& d = m->comps[pidx];
    Comp& p = c[pidx];
    uint64_t dt = now - p.previous_check_time;
    {
      bool is_in_delay = dm_active(p) >= 0;
      bool is_in_process = p.has_item && !is_in_delay;
      bool is_env_blocked = p.env_stopped;
      if (!(is_in_delay || is_env_blocked)) {
        if (p.has_item) {
          uint64_t applied = std::min(dt, p.left);
          p.st.t_a_ns += applied;
          p.left -= applied;
          if (p.left == 0) {
            p.has_item = false;
            *src = log(pidx, *src, ORC_LOG_PROCESS_FINISH, 0, p.item); /* item leaves the model */
            p.st.n_b += 1;
          }
        }
      }
      if (!is_env_blocked && (is_in_delay || is_in_process)) tick_delays(pidx, dt, src);
    }
    refresh_env(pidx, src);
    bool is_env_stopped = p.env_stopped;
    bool has_active_delay = dm_active(p) >= 0 || is_env_stopped;
    if (!p.has_item && !has_active_delay) {
      if (d.upstream >= 0) {
        StockState us = stock_state(d.upstream);
        if (us.cat == CAT_NORMAL || us.cat == CAT_FULL) {
          *src = log(pidx, *src, ORC_LOG_WITHDRAW_REQUEST, 0, 0);
          uint32_t item;
          bool got = stock_remove(d.upstream, *src, &item);
          if (got) {
            double secs = sample(d.dist_time);
            p.left = dur_from(secs);
            p.item = item;
            p.has_item = true;
            *src = log(pidx, *src, ORC_LOG_PROCESS_START, 0, item);
            p.st.n_a += 1;
            p.ttnpe_some = true;
            p.ttnpe = p.left;
          } else {
            *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_UPSTREAM_NO_RESOURCE, 0);
            p.ttnpe_some = false;
          }
        } else {
          *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_UPSTREAM_EMPTY, 0);
          p.ttnpe_some = false;
        }
      } else {
        *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_UPSTREAM_NOT_CONNECTED, 0);
        p.ttnpe_some = false;
      }
    } else if (p.has_item) {
      /* discrete.rs:1127-1129 only refreshes time_to_next_delay_event: ttnpe keeps its old value */
    } else {
      p.ttnpe_some = false; /* discrete.rs:1130-1132 (quirk Q2) */
    }
  }

  /* DiscreteParallelProcess::update_state_impl (discrete.rs:1280-1417) */
  void impl_parallel(uint32_t pidx, EvId* src) {
    const CompDef& d = m->comps[pidx];
    Comp& p = c[pidx];
    uint64_t dt = now - p.previous_check_time;
    {
      bool is_in_delay = dm_active(p) >= 0;
      bool is_in_process = !p.in_progress.empty() && !is_in_delay;
      bool is_env_blocked = p.env_stopped;
      if (!(is_in_delay || is_env_blocked)) {
        /* retain_mut (discrete.rs:1298-1306) */
        std::vector<std::pair<uint64_t, uint32_t> > keep;
        for (size_t i = 0; i < p.in_progress.size(); ++i) {
          uint64_t applied = std::min(dt, p.in_progress[i].first);
          p.st.t_a_ns += applied;
          uint64_t left = p.in_progress[i].first - applied;
          if (left == 0)
            p.complete.push_back(p.in_progress[i].second);
          else
            keep.push_back(std::make_pair(left, p.in_progress[i].second));
        }
        p.in_progress.swap(keep);
        /* drain completed (discrete.rs:1311-1327); the popped item is dropped on `break` */
        while (!p.complete.empty()) {
          uint32_t item = p.complete.front();
          p.complete.pop_front();
          if (d.downstream >= 0) {
            StockState ds = stock_state(d.downstream);
            if (ds.cat == CAT_EMPTY || ds.cat == CAT_NORMAL) {
              *src = log(pidx, *src, ORC_LOG_PROCESS_FINISH, 0, item);
              p.st.n_b += 1;
              stock_add(d.downstream, item, *src);
            } else {
              *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_DOWNSTREAM_FULL, 0);
              break;
            }
          } else {
            *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_DOWNSTREAM_NOT_CONNECTED, 0);
            break;
          }
        }
      }
      if (!is_env_blocked && (is_in_delay || is_in_process)) tick_delays(pidx, dt, src);
    }
    refresh_env(pidx, src);
    if (p.env_stopped) {
      p.ttnpe_some = false;
    } else {
      for (;;) { /* discrete.rs:1373-1398 */
        if (d.upstream >= 0) {
          StockState us = stock_state(d.upstream);
          if (us.cat == CAT_EMPTY || us.cat == CAT_NORMAL) { /* quirk Q4: inverted test */
            *src = log(pidx, *src, ORC_LOG_WITHDRAW_REQUEST, 0, 0);
            uint32_t item;
            bool got = stock_remove(d.upstream, *src, &item);
            if (got) {
              double secs = sample(d.dist_time);
              uint64_t dur = dur_from(secs);
              p.in_progress.push_back(std::make_pair(dur, item));
              *src = log(pidx, *src, ORC_LOG_PROCESS_START, 0, item);
              p.st.n_a += 1;
            } else {
              break;
            }
          } else {
            *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_UPSTREAM_FULL, 0);
            break;
          }
        } else {
          *src = log(pidx, *src, ORC_LOG_PROCESS_NONSTART, ORC_REASON_UPSTREAM_NOT_CONNECTED, 0);
          break;
        }
      }
      if (p.in_progress.empty()) {
        p.ttnpe_some = false;
      } else {
        uint64_t mn = p.in_progress[0].first;
        for (size_t i = 1; i < p.in_progress.size(); ++i) mn = std::min(mn, p.in_progress[i].first);
        p.ttnpe_some = true;
        p.ttnpe = mn;
      }
    }
  }

  void impl_vector(uint32_t, EvId*); /* vector components: oracle/qk_oracle_vector.inc */

  /* ---- BasicEnvironment::set_state (core.rs:256-265) ------------------------------------------ */
  void env_set_state(uint32_t eidx, uint32_t state, EvId src) {
    Comp& e = c[eidx];
    if (e.env_state != state) {
      e.env_state = state;
      EvId id = log(eidx, src, ORC_LOG_ENV_STATE, (uint8_t)state, 0);
      const std::vector<int>& subs = m->comps[eidx].subscribers;
      for (size_t i = 0; i < subs.size(); ++i) {
        update_state(subs[i], id);
        if (status) return;
      }
    }
  }

  /* ---- Model::init for every component, registration order (SURVEY A.6 canonical rule 5) ------- */
  void init() {
    now = t0;
    try {
      /* DelayModes::modify(Add) samples until_delay at build time (delays.rs:159-164) */
      for (size_t i = 0; i < c.size(); ++i) {
        const CompDef& d = m->comps[i];
        c[i].dm.resize(d.delay_modes.size());
        for (size_t k = 0; k < d.delay_modes.size(); ++k) {
          double secs = sample(d.delay_modes[k].until_delay);
          c[i].dm[k].in_fix = false;
          c[i].dm[k].dur = dur_from(secs);
        }
        if (d.kind == ORC_VECTOR_STOCK) {
          for (int k = 0; k < 8; ++k) c[i].vres[k] = d.vinit[k];
          c[i].vlow = d.vlow;
          c[i].area_last_t = t0;
        }
        if (is_item_stock(d.kind)) {
          for (uint32_t k = 0; k < d.n_initial; ++k) {
            const uint32_t h = (63u << 26) | (d.initial_base + k);
            c[i].items.push_back(h);
            if (d.kind == ORC_CONTAINER_STOCK) {
              CPay cp;
              cp.has = d.init_has[k] != 0;
              for (int j = 0; j < 3; ++j) cp.v[j] = cp.has ? d.init_res[k][(size_t)j] : 0.0;
              cpay[h] = cp;
            }
          }
          c[i].area_last_t = t0;
          c[i].st.t_b_ns = d.n_initial;
        }
      }
      for (size_t i = 0; i < c.size(); ++i) {
        const CompDef& d = m->comps[i];
        if (d.kind == ORC_ENVIRONMENT) {
          log((uint32_t)i, EV_INIT, ORC_LOG_ENV_STATE, (uint8_t)c[i].env_state, 0); /* core.rs:229-234 */
        } else if (d.kind == ORC_VECTOR_COMBINER || d.kind == ORC_VECTOR_SPLITTER || d.kind == ORC_VECTOR_SOURCE ||
                   d.kind == ORC_VECTOR_SINK) {
          /* these init with EventId("{code}_{next_event_index:06}") instead of INIT_000000 (vector.rs:774,1111,1393,1619) */
          EvId own = {(uint16_t)i, c[i].next_event_index};
          update_state((uint32_t)i, own);
        } else if (is_process_like(d.kind)) {
          update_state((uint32_t)i, EV_INIT); /* discrete.rs:392-398, vector.rs:355-361 */
        }
      }
      /* externally scheduled events: origin rank 0, FIFO (core.rs:1377-1401) */
      for (size_t i = 0; i < m->ext.size(); ++i)
        schedule(m->ext[i].t, 0, ACT_EXT, m->ext[i].target, EV_SCH, 0, (uint32_t)i);
    } catch (Fault f) {
      status = f.code;
    }
  }

  bool step_one(uint64_t limit) {
    while (!q.empty()) {
      Action a = q.top();
      if (a.type == ACT_KEYED && cancelled[a.key]) {
        q.pop();
        continue;
      }
      if (a.time > limit) return false;
      q.pop();
      now = a.time;
      n_events += 1;
      try {
        switch (a.type) {
          case ACT_KEYED:
            update_state(a.target, a.src);
            break;
          case ACT_EMIT:
            if (is_item_stock(m->comps[a.target].kind))
              stock_emit_change(a.target, a.src);
            else
              vstock_emit_change(a.target, a.src, a.pay_cat, a.pay_occ);
            break;
          case ACT_EXT:
            ext_fire(a);
            break;
        }
      } catch (Fault f) {
        status = f.code;
      }
      return true;
    }
    return false;
  }
  void ext_fire(const Action& a) {
    const ExtEvent& e = m->ext[a.ext_index];
    if (e.type == 0) {
      env_set_state(e.target, e.state, a.src);
    } else if (e.type == 1) {
      ext_set_low_capacity(e);
    } else {
      /* with_process_quantity_distr_inplace(distr_instance), then update_state(SCH_000000): two scheduler actions of
       * origin 0 at the same time, nothing can come between them (core.rs:1389-1390) */
      c[e.target].qty_dist = (int)e.state;
      update_state(e.target, a.src);
    }
  }
  void vstock_emit_change(uint32_t, EvId, int, double);
  void ext_set_low_capacity(const ExtEvent&);
  void vstock_area(uint32_t);

  void finalize_stats() {
    for (size_t i = 0; i < c.size(); ++i)
      if (is_item_stock(m->comps[i].kind)) stock_area((uint32_t)i);
      else if (m->comps[i].kind == ORC_VECTOR_STOCK) vstock_area((uint32_t)i);
  }
};

#include "qk_oracle_vector.inc"
#include "qk_oracle_container.inc"

/* =========================================== C API ========================================== */
extern "C" {

orc_model* orc_model_new(void) { return new orc_model(); }
void orc_model_free(orc_model* m) { delete m; }

int orc_model_add_component(orc_model* m, uint32_t kind, uint32_t low, uint32_t max, uint32_t n_initial) {
  CompDef d;
  d.kind = kind;
  d.low = low;
  d.max = max;
  d.n_initial = n_initial;
  d.initial_base = m->n_initial_total;
  m->n_initial_total += n_initial;
  if (kind == ORC_DISCRETE_SOURCE || kind == ORC_CONTAINER_SOURCE) d.source_ordinal = (int)m->n_sources++;
  if (is_container_kind(kind)) d.dim = 3;
  if (is_process_like(kind)) {
    /* Default::default() for Distribution is Constant(1.) (common.rs:181-185) */
    orc_dist one;
    one.kind = ORC_DIST_CONSTANT;
    one.stream = 0;
    one.p[0] = 1.0;
    one.p[1] = one.p[2] = one.p[3] = 0;
    d.dist_time = (int)m->dists.size();
    m->dists.push_back(one);
    d.dist_qty = (int)m->dists.size();
    m->dists.push_back(one);
  }
  m->comps.push_back(d);
  return (int)m->comps.size() - 1;
}

int orc_model_set_dist(orc_model* m, uint32_t comp, uint32_t which, const orc_dist* d) {
  if (comp >= m->comps.size() || !is_process_like(m->comps[comp].kind)) return -1;
  int id = which == 0 ? m->comps[comp].dist_time : m->comps[comp].dist_qty;
  m->dists[id] = *d;
  return id;
}

int orc_model_add_delay_mode(orc_model* m, uint32_t comp, const orc_dist* ud, const orc_dist* uf, int* ud_id,
                             int* uf_id) {
  if (comp >= m->comps.size() || !is_process_like(m->comps[comp].kind)) return -1;
  DelayModeDef dm;
  dm.until_delay = (int)m->dists.size();
  m->dists.push_back(*ud);
  dm.until_fix = (int)m->dists.size();
  m->dists.push_back(*uf);
  m->comps[comp].delay_modes.push_back(dm);
  if (ud_id) *ud_id = dm.until_delay;
  if (uf_id) *uf_id = dm.until_fix;
  return (int)m->comps[comp].delay_modes.size() - 1;
}

/* connect_components (core.rs:565-1178): which ports each (a,b) pair wires, and subscriber order */
int orc_model_connect(orc_model* m, uint32_t a, uint32_t b, int32_t port_n) {
  if (a >= m->comps.size() || b >= m->comps.size()) return -1;
  CompDef& A = m->comps[a];
  CompDef& B = m->comps[b];
  (void)port_n;
  bool a_stock = A.kind == ORC_DISCRETE_STOCK, b_stock = B.kind == ORC_DISCRETE_STOCK;
  if (a_stock && (B.kind == ORC_DISCRETE_PROCESS || B.kind == ORC_DISCRETE_SINK ||
                  B.kind == ORC_DISCRETE_PARALLEL_PROCESS)) { /* core.rs:886-891,898-903,916-921 */
    if (B.upstream >= 0) return -2;
    A.subscribers.push_back((int)b);
    B.upstream = (int)a;
    return 0;
  }
  if (b_stock && (A.kind == ORC_DISCRETE_PROCESS || A.kind == ORC_DISCRETE_SOURCE ||
                  A.kind == ORC_DISCRETE_PARALLEL_PROCESS)) { /* core.rs:892-897,904-915 */
    if (A.downstream >= 0) return -2;
    B.subscribers.push_back((int)a);
    A.downstream = (int)b;
    return 0;
  }
  /* core.rs:1129-1172; there is no (BasicEnvironment, Vector3ContainerParallelProcess) arm */
  if (A.kind == ORC_ENVIRONMENT && is_process_like(B.kind) && B.kind != ORC_CONTAINER_PARALLEL_PROCESS) {
    if (B.env >= 0) return -2;
    A.subscribers.push_back((int)b);
    B.env = (int)a;
    return 0;
  }
  {
    int rc = orc_connect_vector(m, a, b, port_n);
    if (rc != -3) return rc;
    rc = orc_connect_container(m, a, b);
    if (rc != -3) return rc;
  }
  return -3; /* "No component connection defined from {} to {}" */
}

int orc_model_schedule_env(orc_model* m, uint64_t t_ns, uint32_t env, uint32_t state) {
  if (env >= m->comps.size() || m->comps[env].kind != ORC_ENVIRONMENT) return -1;
  ExtEvent e;
  e.t = t_ns;
  e.type = 0;
  e.target = env;
  e.state = state;
  e.value = 0;
  m->ext.push_back(e);
  return 0;
}
int orc_model_n_dists(const orc_model* m) { return (int)m->dists.size(); }

orc_sim* orc_sim_new(const orc_model* m, uint64_t base_seed, uint64_t rep, uint64_t t0_ns, int log_enabled) {
  orc_sim* s = new orc_sim();
  s->m = m;
  s->seed = base_seed;
  s->rep = rep;
  s->t0 = t0_ns;
  s->now = t0_ns;
  s->status = 0;
  s->n_events = 0;
  s->log_enabled = log_enabled != 0;
  s->c.resize(m->comps.size());
  s->draws.assign(m->dists.size(), 0);
  s->tape.assign(m->dists.size(), (const double*)0);
  s->tape_store.resize(m->dists.size());
  s->next_seq = 0;
  return s;
}
void orc_sim_free(orc_sim* s) { delete s; }
int orc_sim_set_tape(orc_sim* s, uint32_t dist_id, const double* v, uint64_t n) {
  if (dist_id >= s->tape_store.size()) return -1;
  s->tape_store[dist_id].assign(v, v + n);
  s->tape[dist_id] = s->tape_store[dist_id].data();
  if (n == 0) s->tape[dist_id] = (const double*)1; /* an empty tape is still a tape */
  return 0;
}
int orc_sim_init(orc_sim* s) {
  s->init();
  return (int)s->status;
}
int orc_sim_step_until(orc_sim* s, uint64_t t_ns) {
  while (!s->status && s->step_one(t_ns)) {
  }
  if (!s->status && t_ns > s->now) s->now = t_ns; /* step_until leaves now = t */
  return (int)s->status;
}
int orc_sim_step_events(orc_sim* s, uint64_t n) {
  for (uint64_t i = 0; i < n && !s->status; ++i)
    if (!s->step_one(~0ull)) break;
  return (int)s->status;
}
uint32_t orc_sim_status(const orc_sim* s) { return s->status; }
uint64_t orc_sim_now(const orc_sim* s) { return s->now; }
uint64_t orc_sim_n_events(const orc_sim* s) { return s->n_events; }
uint64_t orc_sim_log_count(const orc_sim* s) { return s->logs.size(); }
uint64_t orc_sim_read_log(const orc_sim* s, orc_log_record* buf, uint64_t cap) {
  uint64_t n = std::min<uint64_t>(cap, s->logs.size());
  if (n) std::memcpy(buf, s->logs.data(), n * sizeof(orc_log_record));
  return n;
}
int orc_sim_comp_stats(orc_sim* s, uint32_t comp, orc_comp_stats* out) {
  if (comp >= s->c.size()) return -1;
  s->finalize_stats();
  *out = s->c[comp].st;
  /* busy time is reported up to the CLOCK (not up to the component's last update_state): add the part of the
   * countdown that has elapsed since previous_check_time but has not been applied yet */
  const Comp& p = s->c[comp];
  const uint32_t kind = s->m->comps[comp].kind;
  if (is_process_like(kind) && !(s->dm_active(p) >= 0 || p.env_stopped)) {
    const uint64_t dt = s->now - p.previous_check_time;
    if (is_multi_server(kind)) {
      for (size_t i = 0; i < p.in_progress.size(); ++i) out->t_a_ns += std::min(dt, p.in_progress[i].first);
    } else if (p.has_item) {
      out->t_a_ns += std::min(dt, p.left);
    }
  }
  return 0;
}
uint64_t orc_sim_stock_items(const orc_sim* s, uint32_t comp, uint32_t* buf, uint64_t cap) {
  if (comp >= s->c.size()) return 0;
  const std::deque<uint32_t>& d = s->c[comp].items;
  uint64_t n = std::min<uint64_t>(cap, d.size());
  for (uint64_t i = 0; i < n; ++i) buf[i] = d[i];
  return d.size();
}
uint64_t orc_sim_draws(const orc_sim* s, uint32_t dist_id) {
  return dist_id < s->draws.size() ? s->draws[dist_id] : 0;
}

uint64_t orc_run_batch(const orc_model* m, uint64_t base_seed, uint64_t rep0, uint64_t n_reps, uint64_t t0_ns,
                       uint64_t n_events, uint64_t t_until_ns, int log_enabled, int n_threads,
                       orc_comp_stats* stats_out, uint32_t* status_out, uint64_t* n_events_out, uint64_t* now_out) {
  if (n_threads < 1) n_threads = 1;
  std::vector<uint64_t> totals((size_t)n_threads, 0);
  std::vector<std::thread> th;
  size_t nc = m->comps.size();
  for (int t = 0; t < n_threads; ++t) {
    th.push_back(std::thread([=, &totals]() {
      uint64_t tot = 0;
      for (uint64_t r = (uint64_t)t; r < n_reps; r += (uint64_t)n_threads) {
        orc_sim* s = orc_sim_new(m, base_seed, rep0 + r, t0_ns, log_enabled);
        orc_sim_init(s);
        if (n_events)
          orc_sim_step_events(s, n_events);
        else
          orc_sim_step_until(s, t_until_ns);
        tot += s->n_events;
        if (stats_out)
          for (size_t k = 0; k < nc; ++k) orc_sim_comp_stats(s, (uint32_t)k, &stats_out[r * nc + k]);
        if (status_out) status_out[r] = s->status;
        if (n_events_out) n_events_out[r] = s->n_events;
        if (now_out) now_out[r] = s->now;
        orc_sim_free(s);
      }
      totals[(size_t)t] = tot;
    }));
  }
  uint64_t total = 0;
  for (int t = 0; t < n_threads; ++t) {
    th[(size_t)t].join();
    total += totals[(size_t)t];
  }
  return total;
}

int orc_from_secs_f64(double secs, uint64_t* ns_out) { return from_secs_f64(secs, ns_out); }
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr, key, out);
}
double orc_ln(double x) { return det_ln(x); }
double orc_sample_native(const orc_dist* d, uint64_t base_seed, uint64_t rep, uint32_t* draw_counter) {
  return sample_native(*d, base_seed, rep, *draw_counter);
}

/* stand-alone DelayModes harness so the reference's two unit tests (delays.rs:182-242) can be
 * replayed against exactly the dm_* code the simulator uses */
struct orc_delay_modes {
  orc_model m;
  orc_sim* s;
};
orc_delay_modes* orc_dm_new(void) {
  orc_delay_modes* h = new orc_delay_modes();
  orc_model_add_component(&h->m, ORC_DISCRETE_PROCESS, 0, 0, 0);
  h->s = 0;
  return h;
}
void orc_dm_free(orc_delay_modes* h) {
  if (h->s) orc_sim_free(h->s);
  delete h;
}
int orc_dm_add(orc_delay_modes* h, double ud, double uf) {
  orc_dist a, b;
  std::memset(&a, 0, sizeof(a));
  std::memset(&b, 0, sizeof(b));
  a.kind = b.kind = ORC_DIST_CONSTANT;
  a.p[0] = ud;
  b.p[0] = uf;
  int k = orc_model_add_delay_mode(&h->m, 0, &a, &b, 0, 0);
  /* (re)build the sim state and sample the new mode's until_delay, as modify(Add) does */
  std::vector<DelayModeState> old;
  if (h->s) {
    old = h->s->c[0].dm;
    orc_sim_free(h->s);
  }
  h->s = orc_sim_new(&h->m, 0, 0, 0, 0);
  h->s->c[0].dm = old;
  DelayModeState st;
  st.in_fix = false;
  st.dur = h->s->dur_from(h->s->sample(h->m.comps[0].delay_modes[(size_t)k].until_delay));
  h->s->c[0].dm.push_back(st);
  return k;
}
void orc_dm_update_state(orc_delay_modes* h, uint64_t dt, int* from, int* to) {
  h->s->dm_update_state(0, dt, from, to);
}
int orc_dm_next_event(const orc_delay_modes* h, uint64_t* dur) { return h->s->dm_next_event(h->s->c[0], dur); }

} /* extern "C" */
