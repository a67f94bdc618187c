/*
 * qk_oracle.h -- C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a single-threaded, one-replication-at-a-time CPU restatement of the
 * QuokkaSim (jajetloh/quokkasim v0.2.2) component semantics on top of a restated
 * NeXosim 0.3.3 scheduler.  It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check (and time) the CUDA
 * path against it.  Nothing under quokkasim_b200/ may include, link or call it.
 *
 * PARITY PIN STATUS: the reference's engine (nexosim 0.3.3, Cargo.lock:517), RNG
 * (rand 0.8.5 / rand_distr 0.4.3) and time type (tai-time 0.3.3) are un-vendored
 * third-party crates and there is no Rust toolchain in this image, so the reference
 * cannot be run here.  The only value-pinning tests the reference holds for this path
 * are the two DelayModes unit tests (quokkasim/src/delays.rs:182-242); the oracle is
 * checked against both (tests/test_oracle_kat.py).  Everything else is pinned by the
 * hand-derived KAT-0 trace and, for the container family, the hand-derived KAT-1 trace
 * (tests/test_csv_export.py, tests/golden/kat1_*.csv), Rust std's documented
 * Duration::try_from_secs_f64 examples, and the Random123 Philox4x32-10 known answers.
 * The continuous (vector.rs) and container (vector_container.rs) restatements have no reference test or golden
 * file to pin them: "parity unpinned" against the real engine, anchored on the reference's call sites and checked by
 * conservation laws.
 * Scheduler tie-break order is "parity unpinned" against the real engine
 * (the reference itself leaves same-timestamp cross-model order unspecified).
 *
 * The numeric codes below restate the public ABI of include/quokka_b200.h; they are
 * duplicated on purpose so that the oracle does not depend on product headers.
 */
#ifndef QK_ORACLE_H
#define QK_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* component kinds */
enum {
  ORC_DISCRETE_STOCK = 1,
  ORC_DISCRETE_PROCESS = 2,
  ORC_DISCRETE_SOURCE = 3,
  ORC_DISCRETE_SINK = 4,
  ORC_DISCRETE_PARALLEL_PROCESS = 5,
  ORC_ENVIRONMENT = 6,
  ORC_VECTOR_STOCK = 16,
  ORC_VECTOR_PROCESS = 17,
  ORC_VECTOR_SOURCE = 18,
  ORC_VECTOR_SINK = 19,
  ORC_VECTOR_COMBINER = 20,
  ORC_VECTOR_SPLITTER = 21,
  /* container family (components/vector_container.rs; ComponentModel::Vector3Container*, core.rs:549-555):
   * discrete items that carry an Option<Vector3> resource and a capacity */
  ORC_CONTAINER_STOCK = 32,
  ORC_CONTAINER_PROCESS = 33,
  ORC_CONTAINER_SOURCE = 34,
  ORC_CONTAINER_SINK = 35,
  ORC_CONTAINER_PARALLEL_PROCESS = 36,
  ORC_CONTAINER_LOAD_PROCESS = 37,
  ORC_CONTAINER_UNLOAD_PROCESS = 38
};

/* distribution kinds (common.rs:37-44) */
enum {
  ORC_DIST_CONSTANT = 0,
  ORC_DIST_UNIFORM = 1,
  ORC_DIST_TRIANGULAR = 2,
  ORC_DIST_NORMAL = 3,
  ORC_DIST_TRUNCNORMAL = 4,
  ORC_DIST_EXPONENTIAL = 5
};

/* log record kinds */
enum {
  ORC_LOG_PROCESS_START = 1,
  ORC_LOG_PROCESS_CONTINUE = 2,
  ORC_LOG_PROCESS_FINISH = 3,
  ORC_LOG_PROCESS_NONSTART = 4,
  ORC_LOG_PROCESS_STOPPED = 5,
  ORC_LOG_WITHDRAW_REQUEST = 6,
  ORC_LOG_DELAY_START = 7,
  ORC_LOG_DELAY_END = 8,
  ORC_LOG_STOCK_ADD = 16,
  ORC_LOG_STOCK_REMOVE = 17,
  ORC_LOG_STOCK_STATE_CHANGE = 18,
  ORC_LOG_ENV_STATE = 24,
  /* vector (continuous) components: records carry fp64 payloads, see DESIGN.md */
  ORC_LOG_VSTOCK_ADD = 32,
  ORC_LOG_VSTOCK_REMOVE = 33,
  ORC_LOG_VSTOCK_STATE_CHANGE = 34,
  ORC_LOG_VPROC_START = 40,
  ORC_LOG_VPROC_SUCCESS = 41,
  ORC_LOG_VPROC_FAILURE = 42,
  ORC_LOG_VPROC_COMBINE_START = 43,
  ORC_LOG_VPROC_COMBINE_SUCCESS = 44,
  ORC_LOG_VPROC_SPLIT_START = 45,
  ORC_LOG_VPROC_SPLIT_SUCCESS = 46,
  ORC_LOG_VEC_DATA = 63 /* continuation record carrying 3 doubles */
};

/* reason codes (the &'static str reasons of DiscreteProcessLogType) */
enum {
  ORC_REASON_NONE = 0,
  ORC_REASON_UPSTREAM_EMPTY = 1,         /* "Upstream is empty" */
  ORC_REASON_UPSTREAM_NOT_CONNECTED = 2, /* "Upstream is not connected" */
  ORC_REASON_DOWNSTREAM_FULL = 3,        /* "Downstream is full" */
  ORC_REASON_DOWNSTREAM_NOT_CONNECTED = 4,
  ORC_REASON_UPSTREAM_NO_RESOURCE = 5,   /* "Upstream did not provide resource" */
  ORC_REASON_STOPPED_BY_ENV = 6,         /* "Stopped by environment" */
  ORC_REASON_RESUMED_BY_ENV = 7,         /* "Resumed by environment" */
  ORC_REASON_UPSTREAM_FULL = 8,          /* "Upstream is full" (ParallelProcess quirk Q4) */
  ORC_REASON_AN_UPSTREAM_EMPTY = 9,      /* "At least one upstream is empty" */
  ORC_REASON_NO_UPSTREAMS = 10,          /* "No upstreams are connected" */
  ORC_REASON_A_DOWNSTREAM_FULL = 11,     /* "At least one downstream is full" */
  ORC_REASON_NO_DOWNSTREAMS = 12,        /* "No downstreams are connected" */
  ORC_REASON_NO_PROCESS_SLOTS = 13,      /* "No remaining process slots"          vector_container.rs:369,676 */
  ORC_REASON_DS_CONTAINERS_FULL = 14,    /* "Downstream container stock is full"  vector_container.rs:601 */
  ORC_REASON_DS_RESOURCE_FULL = 15       /* "Downstream resource stock is full"   vector_container.rs:605 */
};

/* per-replication status */
enum {
  ORC_OK = 0,
  ORC_FAULT_ZERO_DURATION = 1, /* panic!("Time until next event is zero!") discrete.rs:541 */
  ORC_FAULT_BAD_DURATION = 2,  /* Duration::from_secs_f64 panics: negative / NaN / overflow */
  ORC_FAULT_TAPE_EXHAUSTED = 3,
  ORC_FAULT_SOURCE_VECTOR = 8, /* panic!("Source vector has total 0 or negative") vector.rs:1490 */
  ORC_FAULT_NOT_CONNECTED = 9  /* `.next().unwrap()` on a Requestor nobody connected (vector_container.rs:358) */
};

#define ORC_SRC_INIT 0xFFFFu /* EventId::from_init()      common.rs:14 */
#define ORC_SRC_SCH 0xFFFEu  /* EventId::from_scheduler() common.rs:18 */
#define ORC_ITEM_NONE 0xFFFFFFFFFFFFFFFFull
/* a container whose `resource` is None: first double of its vector carries this (signalling-NaN) bit pattern */
#define ORC_CONTAINER_NONE_BITS 0x7FF40000000000C0ull

typedef struct {
  uint64_t time_ns;
  uint16_t comp;
  uint8_t kind;
  uint8_t reason;
  uint32_t seq;
  uint16_t src_comp;
  uint16_t aux;
  uint32_t src_seq;
  uint64_t payload;
} orc_log_record; /* 32 bytes */

typedef struct {
  uint32_t kind;
  uint32_t stream; /* seed the DistributionFactory would have assigned (common.rs:73-148) */
  double p[4];
} orc_dist;

/* per-component summary statistics of one replication */
typedef struct {
  uint64_t n_a;    /* process-like: ProcessStart count ; stock: Add count  */
  uint64_t n_b;    /* process-like: ProcessFinish count; stock: Remove count */
  uint64_t t_a_ns; /* process-like: busy ns ; stock: time-integral of occupancy (items*ns) */
  uint64_t t_b_ns; /* process-like: ns spent in an active delay ; stock: max occupancy */
  double fvalue;   /* vector stock: current total ; vector process-like: sum of success quantities */
  double faux;     /* vector stock: integral of total dt (unit*s) ; vector process-like: in-flight total */
} orc_comp_stats;

typedef struct orc_model orc_model;
typedef struct orc_sim orc_sim;

orc_model* orc_model_new(void);
void orc_model_free(orc_model*);
/* returns component id (= registration rank) or <0 */
int orc_model_add_component(orc_model*, uint32_t kind, uint32_t low_capacity, uint32_t max_capacity,
                            uint32_t n_initial_items);
/* which: 0 = process_time_distr, 1 = process_quantity_distr. returns distribution instance id */
int orc_model_set_dist(orc_model*, uint32_t comp, uint32_t which, const orc_dist*);
/* returns delay-mode index within the component; sets the two distribution instance ids */
int orc_model_add_delay_mode(orc_model*, uint32_t comp, const orc_dist* until_delay, const orc_dist* until_fix,
                             int* until_delay_id, int* until_fix_id);
int orc_model_connect(orc_model*, uint32_t a, uint32_t b, int32_t port_n);
int orc_model_schedule_env(orc_model*, uint64_t t_ns, uint32_t env, uint32_t state);
int orc_model_n_dists(const orc_model*);
/* vector components (fp64) */
int orc_model_set_vector(orc_model*, uint32_t comp, uint32_t dim, double low, double max, const double* init);
/* user-defined resource (quokkasim_examples/src/bin/iron_ore.rs:21-96): dim <= 8 fields, ResourceTotal sums the fields
 * named by total_mask in field order, ResourceRemove is proportional over every field */
int orc_model_set_vector_n(orc_model*, uint32_t comp, uint32_t dim, uint32_t total_mask, double low, double max,
                           const double* init);
int orc_model_set_splits(orc_model*, uint32_t comp, uint32_t n, const double* ratios);
int orc_model_schedule_low_capacity(orc_model*, uint64_t t_ns, uint32_t stock, double value);
int orc_model_schedule_process_quantity(orc_model*, uint64_t t_ns, uint32_t process, const orc_dist* d);
/* container family. Vector3ContainerFactory (vector_container.rs:126-156): create_quantity_distr (NULL = None),
 * create_component_split, capacity; *dist_id_out = distribution instance id or -1 */
int orc_model_set_container_factory(orc_model*, uint32_t source, const orc_dist* create_quantity, const double* split3,
                                    double capacity, int* dist_id_out);
/* with_initial_resource(ItemDeque<Vector3Container>) of a container stock: resources [n][3], has_resource [n],
 * capacities [n] */
int orc_model_set_initial_containers(orc_model*, uint32_t stock, uint32_t n, const double* resources,
                                     const uint8_t* has_resource, const double* capacities);

orc_sim* orc_sim_new(const orc_model*, uint64_t base_seed, uint64_t rep, uint64_t t0_ns, int log_enabled);
void orc_sim_free(orc_sim*);
int orc_sim_set_tape(orc_sim*, uint32_t dist_id, const double* values, uint64_t n);
int orc_sim_init(orc_sim*);
int orc_sim_step_until(orc_sim*, uint64_t t_ns);
int orc_sim_step_events(orc_sim*, uint64_t n_events);
uint32_t orc_sim_status(const orc_sim*);
uint64_t orc_sim_now(const orc_sim*);
uint64_t orc_sim_n_events(const orc_sim*);
uint64_t orc_sim_log_count(const orc_sim*);
uint64_t orc_sim_read_log(const orc_sim*, orc_log_record* buf, uint64_t cap);
int orc_sim_comp_stats(orc_sim*, uint32_t comp, orc_comp_stats* out);
/* discrete stock contents (item handles, FIFO order); returns count */
uint64_t orc_sim_stock_items(const orc_sim*, uint32_t comp, uint32_t* buf, uint64_t cap);
/* vector stock contents */
int orc_sim_vector_resource(const orc_sim*, uint32_t comp, double* out3);
int orc_sim_vector_resource_n(const orc_sim*, uint32_t comp, double* out, uint32_t cap); /* returns dim */
uint64_t orc_sim_draws(const orc_sim*, uint32_t dist_id);
/* containers held by a container stock (FIFO order) or a container process-like component (in progress, then
 * complete): handles [cap], resources [cap][3] (first double = ORC_CONTAINER_NONE_BITS for None); returns count */
uint64_t orc_sim_containers(const orc_sim*, uint32_t comp, uint32_t* handles, double* resources, uint64_t cap);

/* batch runner (CPU baseline): runs reps [rep0, rep0+n) with native Philox streams on n_threads host threads.
 * stats_out: [n_reps][n_comp] orc_comp_stats or NULL; status_out: [n_reps] or NULL. returns total events executed. */
uint64_t orc_run_batch(const orc_model*, uint64_t base_seed, uint64_t rep0, uint64_t n_reps, uint64_t t0_ns,
                       uint64_t n_events, uint64_t t_until_ns, int log_enabled, int n_threads,
                       orc_comp_stats* stats_out, uint32_t* status_out, uint64_t* n_events_out,
                       uint64_t* now_out);

/* primitives exposed for known-answer tests */
int orc_from_secs_f64(double secs, uint64_t* ns_out); /* 0 ok, else fault code */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_ln(double x);
double orc_sample_native(const orc_dist*, uint64_t base_seed, uint64_t rep, uint32_t* draw_counter);

/* DelayModes stand-alone harness (for the two reference unit tests, delays.rs:182-242) */
typedef struct orc_delay_modes orc_delay_modes;
orc_delay_modes* orc_dm_new(void);
void orc_dm_free(orc_delay_modes*);
int orc_dm_add(orc_delay_modes*, double until_delay_const, double until_fix_const);
/* returns from/to as mode index or -1 */
void orc_dm_update_state(orc_delay_modes*, uint64_t dt_ns, int* from, int* to);
int orc_dm_next_event(const orc_delay_modes*, uint64_t* dur_ns); /* returns mode index or -1 */

#ifdef __cplusplus
}
#endif
#endif
