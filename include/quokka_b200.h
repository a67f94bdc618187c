/*
 * quokka_b200.h -- C ABI of the B200 batched-replication executor for QuokkaSim models.
 *
 * This is the drop-in boundary for ONE hot path of jajetloh/quokkasim v0.2.2: advancing many
 * independent replications of one stock-and-flow model in lockstep.  The reference has no FFI of
 * its own; what its generated code binds is the NeXosim engine API.  Each entry point below names
 * the reference call it replaces (paths relative to /root/reference/quokkasim/src unless noted):
 *
 *   qk_model_create / qk_model_add_component   <- ComponentModel::register_component ->
 *                                                 SimInit::add_model         core.rs:1180-1196, 419-423
 *   qk_model_set_dist                          <- with_process_time_distr / with_process_quantity_distr
 *                                                 (quokkasim_derive_macros/src/lib.rs:73-78) with a
 *                                                 Distribution built by DistributionFactory::create
 *                                                                                      common.rs:73-148
 *   qk_model_add_delay_mode                    <- with_delay_mode(DelayModeChange::Add) delays.rs:157-164
 *   qk_model_connect                           <- connect_components!(&mut a,&mut b[,n]) core.rs:565-1178
 *   qk_model_set_logged                        <- connect_logger!(&mut logger,&mut c)   core.rs:1303-1338
 *   qk_model_schedule_env_state                <- create_scheduled_event!(SetEnvironmentState)
 *                                                                                      core.rs:1393-1396
 *   qk_model_schedule_low_capacity             <- create_scheduled_event!(SetLowCapacity)  core.rs:1382-1386
 *   qk_model_schedule_process_quantity         <- create_scheduled_event!(SetProcessQuantity) core.rs:1387-1391
 *   qk_batch_create + qk_batch_init            <- SimInit::init(t0) (one Simulation per replication)
 *                                                 quokkasim_examples/src/bin/discrete_queue.rs:109
 *   qk_batch_step_until / qk_batch_step_events <- Simulation::step_until / step   discrete_queue.rs:111
 *   qk_batch_read_log                          <- Logger::get_buffer().into_reader()    core.rs:388-401
 *   qk_batch_comp_stats / qk_batch_reduce_stats<- (no counterpart: the reference has no summaries;
 *                                                 defined from the log records, DESIGN.md "Statistics")
 *
 * Conventions
 *   - every function returns QK_OK (0) or a negative qk_status; qk_last_error() gives the
 *     thread-local message of the last failure on the calling thread
 *   - plain pointers and sizes only; host buffers are caller-owned and copied; the library owns
 *     all device memory of a batch; *_destroy frees
 *   - a qk_batch is single-owner (not thread-safe), like `&mut Simulation`
 *   - there is NO CPU fallback: qk_batch_create fails with QK_ERR_NO_DEVICE without a CUDA device
 *   - per-replication faults (the reference's panics, e.g. "Time until next event is zero!"
 *     discrete.rs:541) never abort the batch; they freeze that replication and are reported in the
 *     per-replication status array (qk_batch_read_status)
 */
#ifndef QUOKKA_B200_H
#define QUOKKA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QK_ABI_VERSION 2

typedef enum {
  QK_OK = 0,
  QK_ERR_INVALID_ARG = -1,
  QK_ERR_UNSUPPORTED_CONNECTION = -2, /* "No component connection defined from {} to {}" */
  QK_ERR_PORT_IN_USE = -3,            /* second connection on a single-reply port */
  QK_ERR_NO_DEVICE = -4,
  QK_ERR_CUDA = -5,
  QK_ERR_STATE = -6, /* call out of order (e.g. step before init) */
  QK_ERR_CAPACITY = -7, /* model exceeds a compiled-in or shared-memory limit */
  QK_ERR_BUFFER_TOO_SMALL = -8
} qk_status;

/* component kinds (ComponentModel variants, core.rs:505-557; the item/vector element type is not
 * part of the kind: items are opaque 32-bit handles on the device) */
typedef enum {
  QK_DISCRETE_STOCK = 1,            /* DiscreteStock            discrete.rs:38-59   */
  QK_DISCRETE_PROCESS = 2,          /* DiscreteProcess          discrete.rs:320-355 */
  QK_DISCRETE_SOURCE = 3,           /* DiscreteSource           discrete.rs:689-723 */
  QK_DISCRETE_SINK = 4,             /* DiscreteSink             discrete.rs:942-974 */
  QK_DISCRETE_PARALLEL_PROCESS = 5, /* DiscreteParallelProcess  discrete.rs:1188-1224 */
  QK_ENVIRONMENT = 6,               /* BasicEnvironment         core.rs:218-226 */
  /* continuous components (components/vector.rs); resource f64 or Vector3, see qk_model_set_vector */
  QK_VECTOR_STOCK = 16,    /* VectorStock    vector.rs:42-62   */
  QK_VECTOR_PROCESS = 17,  /* VectorProcess  vector.rs:278-320 */
  QK_VECTOR_SOURCE = 18,   /* VectorSource   vector.rs:1324-1358 */
  QK_VECTOR_SINK = 19,     /* VectorSink     vector.rs:1578-1612 */
  QK_VECTOR_COMBINER = 20, /* VectorCombiner vector.rs:726-765 */
  QK_VECTOR_SPLITTER = 21, /* VectorSplitter vector.rs:1026-1065 */
  /* container family (components/vector_container.rs; ComponentModel::Vector3Container*, core.rs:549-555): discrete
   * items that carry an Option<Vector3> resource and a capacity.  Stock/Process/ParallelProcess/Source/Sink are the
   * discrete components instantiated with Vector3Container; Load/Unload move resource between a Vector3Stock and the
   * containers.  For Load/Unload `max_capacity` of the descriptor is max_process_count (default 1). */
  QK_CONTAINER_STOCK = 32,            /* Vector3ContainerStock            core.rs:549 */
  QK_CONTAINER_PROCESS = 33,          /* Vector3ContainerProcess          core.rs:550 */
  QK_CONTAINER_SOURCE = 34,           /* Vector3ContainerSource           core.rs:552 */
  QK_CONTAINER_SINK = 35,             /* Vector3ContainerSink             core.rs:553 */
  QK_CONTAINER_PARALLEL_PROCESS = 36, /* Vector3ContainerParallelProcess  core.rs:551 */
  QK_CONTAINER_LOAD_PROCESS = 37,     /* ContainerLoadingProcess          vector_container.rs:158-196 */
  QK_CONTAINER_UNLOAD_PROCESS = 38    /* ContainerUnloadingProcess        vector_container.rs:445-483 */
} qk_component_kind;

/* DistributionConfig (common.rs:37-44) */
typedef enum {
  QK_DIST_CONSTANT = 0,    /* p0 = value */
  QK_DIST_UNIFORM = 1,     /* p0 = min, p1 = max */
  QK_DIST_TRIANGULAR = 2,  /* p0 = min, p1 = max, p2 = mode */
  QK_DIST_NORMAL = 3,      /* p0 = mean, p1 = std */
  QK_DIST_TRUNCNORMAL = 4, /* p0 = mean, p1 = std, p2 = min, p3 = max */
  QK_DIST_EXPONENTIAL = 5  /* p0 = mean */
} qk_dist_kind;

typedef struct {
  uint32_t kind;   /* qk_dist_kind */
  uint32_t stream; /* the seed DistributionFactory::create would have given this instance
                      (common.rs:76,83,97,120,133): instances with equal `stream` draw identical
                      sequences, which reproduces the factory's seed sharing */
  double p[4];
} qk_dist_desc;

typedef struct {
  uint32_t kind;            /* qk_component_kind */
  uint32_t low_capacity;    /* DiscreteStock.low_capacity (default 0)  discrete.rs:70 */
  uint32_t max_capacity;    /* DiscreteStock.max_capacity (default 1)  discrete.rs:71 */
  uint32_t n_initial_items; /* with_initial_resource(ItemDeque) length; handles are assigned by the library */
  uint32_t capacity_hint;   /* 0 = derive. Stock: item-ring slots; ParallelProcess: in-flight slots */
  uint32_t initial_env_state; /* BasicEnvironment::with_state: 0 Normal, 1 Stopped */
} qk_component_desc;

/* log records: one per `self.log(..)` call of a component (discrete.rs:235-251,566-582; core.rs:273-289) */
typedef enum {
  QK_LOG_PROCESS_START = 1,    /* payload = item handle */
  QK_LOG_PROCESS_CONTINUE = 2, /* reason */
  QK_LOG_PROCESS_FINISH = 3,   /* payload = item handle */
  QK_LOG_PROCESS_NONSTART = 4, /* reason */
  QK_LOG_PROCESS_STOPPED = 5,  /* reason */
  QK_LOG_WITHDRAW_REQUEST = 6,
  QK_LOG_DELAY_START = 7, /* payload = delay-mode index */
  QK_LOG_DELAY_END = 8,
  QK_LOG_STOCK_ADD = 16,          /* payload = item handle */
  QK_LOG_STOCK_REMOVE = 17,       /* payload = item handle or QK_ITEM_NONE */
  QK_LOG_STOCK_STATE_CHANGE = 18, /* reason = 0 Empty / 1 Normal / 2 Full; payload = occupied<<32 | empty */
  QK_LOG_ENV_STATE = 24,          /* reason = 0 Normal / 1 Stopped */
  /* vector components: `payload` holds the bits of an fp64 scalar; each vector of the record follows as one
   * QK_LOG_VEC_DATA continuation record (reason = index, seq = seq of the record it continues) whose three doubles
   * overlay time_ns, (src_comp|aux|src_seq) and payload */
  QK_LOG_VSTOCK_ADD = 32,          /* payload = balance after ; 1 vector = what was added */
  QK_LOG_VSTOCK_REMOVE = 33,       /* payload = balance after ; 1 vector = what was removed */
  QK_LOG_VSTOCK_STATE_CHANGE = 34, /* reason = state ; payload = occupied ; vector[0] = empty */
  QK_LOG_VPROC_START = 40,         /* payload = quantity ; 1 vector */
  QK_LOG_VPROC_SUCCESS = 41,       /* payload = quantity ; 1 vector */
  QK_LOG_VPROC_FAILURE = 42,       /* reason */
  QK_LOG_VPROC_COMBINE_START = 43, /* payload = quantity ; M vectors */
  QK_LOG_VPROC_COMBINE_SUCCESS = 44,
  QK_LOG_VPROC_SPLIT_START = 45,
  QK_LOG_VPROC_SPLIT_SUCCESS = 46, /* payload = total ; N vectors */
  QK_LOG_VEC_DATA = 63
} qk_log_kind;

typedef enum {
  QK_REASON_NONE = 0,
  QK_REASON_UPSTREAM_EMPTY = 1,           /* "Upstream is empty" */
  QK_REASON_UPSTREAM_NOT_CONNECTED = 2,   /* "Upstream is not connected" */
  QK_REASON_DOWNSTREAM_FULL = 3,          /* "Downstream is full" */
  QK_REASON_DOWNSTREAM_NOT_CONNECTED = 4, /* "Downstream is not connected" */
  QK_REASON_UPSTREAM_NO_RESOURCE = 5,     /* "Upstream did not provide resource" */
  QK_REASON_STOPPED_BY_ENV = 6,           /* "Stopped by environment" */
  QK_REASON_RESUMED_BY_ENV = 7,           /* "Resumed by environment" */
  QK_REASON_UPSTREAM_FULL = 8,            /* "Upstream is full" (DiscreteParallelProcess, discrete.rs:1390) */
  QK_REASON_AN_UPSTREAM_EMPTY = 9,        /* "At least one upstream is empty"   vector.rs:938 */
  QK_REASON_NO_UPSTREAMS = 10,            /* "No upstreams are connected"       vector.rs:942 */
  QK_REASON_A_DOWNSTREAM_FULL = 11,       /* "At least one downstream is full"  vector.rs:1231 */
  QK_REASON_NO_DOWNSTREAMS = 12,          /* "No downstreams are connected"     vector.rs:1235 */
  QK_REASON_NO_PROCESS_SLOTS = 13,        /* "No remaining process slots"          vector_container.rs:369,676 */
  QK_REASON_DS_CONTAINERS_FULL = 14,      /* "Downstream container stock is full"  vector_container.rs:601 */
  QK_REASON_DS_RESOURCE_FULL = 15         /* "Downstream resource stock is full"   vector_container.rs:605 */
} qk_reason;

#define QK_SRC_INIT 0xFFFFu /* source_event_id "INIT_000000" (common.rs:14) */
#define QK_SRC_SCH 0xFFFEu  /* source_event_id "SCH_000000"  (common.rs:18) */
#define QK_ITEM_NONE 0xFFFFFFFFFFFFFFFFull
/* item handle = (source ordinal << 26) | index; ordinal 63 = initial stock contents */
#define QK_ITEM_SOURCE(h) ((uint32_t)(h) >> 26)
#define QK_ITEM_INDEX(h) ((uint32_t)(h)&0x3FFFFFFu)
/* a container whose `resource` is None: the first double of its vector carries this (signalling-NaN) bit pattern.
 * Records that serialise a container (ProcessStart / ProcessFinish / Add / Remove(Some)) of a container component are
 * followed by ONE QK_LOG_VEC_DATA continuation record holding the container's resource. */
#define QK_CONTAINER_NONE_BITS 0x7FF40000000000C0ull

typedef struct {
  uint64_t time_ns;  /* absolute ns since MonotonicTime::EPOCH */
  uint16_t comp;     /* component id (registration order) */
  uint8_t kind;      /* qk_log_kind */
  uint8_t reason;    /* qk_reason or state category */
  uint32_t seq;      /* next_event_index at the time of the call: EventId = "{code}_{seq:06}" */
  uint16_t src_comp; /* source_event_id component, or QK_SRC_INIT / QK_SRC_SCH */
  uint16_t aux;      /* 0 (device-internal lane tag, cleared on read) */
  uint32_t src_seq;
  uint64_t payload;
} qk_log_record; /* 32 bytes */

/* per-replication fault codes */
typedef enum {
  QK_REP_OK = 0,
  QK_REP_ZERO_DURATION = 1,  /* panic!("Time until next event is zero!") */
  QK_REP_BAD_DURATION = 2,   /* Duration::from_secs_f64 panic (negative / NaN / overflow) */
  QK_REP_TAPE_EXHAUSTED = 3, /* variate tape shorter than the number of sample() calls */
  QK_REP_STOCK_OVERFLOW = 4, /* item ring full (stock capacity is advisory in the reference) */
  QK_REP_EMIT_OVERFLOW = 5,  /* too many pending emit_change events on one stock */
  QK_REP_PARALLEL_OVERFLOW = 6,
  QK_REP_ITEM_INDEX_OVERFLOW = 7,
  QK_REP_SOURCE_VECTOR = 8, /* panic!("Source vector has total 0 or negative") vector.rs:1490 */
  QK_REP_NOT_CONNECTED = 9  /* `.next().unwrap()` on a port nobody connected (vector_container.rs:358) */
} qk_rep_status;

/* per-component, per-replication summary (exact integers) */
typedef struct {
  uint64_t count;   /* process-like: ProcessFinish records ; stock: Add records */
  uint64_t time_ns; /* process-like: busy ns (countdown actually consumed) ; stock: integral of occupancy, item*ns */
  uint64_t aux;     /* process-like: ns spent in an active delay ; stock: maximum occupancy seen */
  uint64_t level;   /* process-like: items in progress now ; stock: occupancy now (vector stock: state category) */
  double fvalue;    /* vector stock: current total ; vector process-like: sum of success quantities ; else 0 */
  double faux;      /* vector stock: integral of total dt (unit*s) ; vector process-like: in-flight total ; else 0 */
} qk_comp_stats;

/* whole-batch reduction of qk_comp_stats over the replications of THIS batch (one GPU).
 * sums are exact (128-bit as hi/lo); multi-GPU callers add the fields across ranks
 * (NCCL sum over the u64 arrays, min/max on the extrema) -- see qk_batch_reduce_stats_device. */
#define QK_HIST_BINS 32
typedef struct {
  uint64_t sum_count;
  uint64_t sum_time_lo, sum_time_hi; /* 128-bit sum of time_ns = hi * 2^64 + lo */
  uint64_t sum_level;
  uint64_t min_count, max_count;
  uint64_t min_time, max_time;
  double sumsq_time; /* sum over replications of (time_ns * 1e-9)^2: the exact integer below, rounded once */
  double sumsq_count;
  uint64_t hist_time[QK_HIST_BINS]; /* replications binned on time_ns over [hist_lo_time, hist_hi_time] */
  uint64_t hist_count[QK_HIST_BINS];
  uint64_t hist_lo_time, hist_hi_time, hist_lo_count, hist_hi_count;
  /* exact second moments (ABI 2): independent of how the replications are spread over blocks or GPUs */
  uint64_t sumsq_time_ns2[3];              /* sum of time_ns^2 as a 192-bit integer, least significant limb first */
  uint64_t sumsq_count_lo, sumsq_count_hi; /* sum of count^2, 128-bit */
  uint64_t sumsq_level;                    /* sum of level^2 (stock occupancy now / items in progress now) */
} qk_comp_summary;

/* Packed partial sums of one component over the replications of one batch: QK_PARTIAL_WORDS u64 words.  This is what
 * crosses GPUs -- ONE all-gather of [n_comp][QK_PARTIAL_WORDS] per rank, then qk_batch_combine_partials_device -- and
 * every word is an exact integer: the four extrema words combine with unsigned MIN (maxima are stored complemented),
 * all others with a wrapping SUM.  Sums that can exceed 64 bits travel as the sums of their low and high 32-bit halves:
 * time_ns = a * 2^32 + b gives sum(time_ns) = TIME_HI * 2^32 + TIME_LO and sum(time_ns^2) = AA * 2^64 + 2 * AB * 2^32 +
 * BB with AA = AA_HI * 2^32 + AA_LO the sum of a^2 (AB, BB, CC = sum of count^2 alike). */
#define QK_PARTIAL_WORDS 20
enum {
  QP_COUNT = 0, QP_TIME_LO = 1, QP_TIME_HI = 2, QP_LEVEL = 3,
  QP_NEVENTS = 4, QP_NFAULTED = 5, /* whole batch, stored with component 0 only */
  QP_MIN_COUNT = 6, QP_MIN_TIME = 7, QP_NMAX_COUNT = 8, QP_NMAX_TIME = 9, /* ~max so that MIN serves both */
  QP_AA_LO = 10, QP_AA_HI = 11, QP_AB_LO = 12, QP_AB_HI = 13, QP_BB_LO = 14, QP_BB_HI = 15,
  QP_CC_LO = 16, QP_CC_HI = 17, QP_LEVEL_SQ = 18,
  QP_NREPS = 19 /* replications reduced, stored with component 0 only */
};

typedef struct {
  uint64_t n_reps;
  uint64_t n_events;   /* executed scheduler actions, all replications */
  uint64_t n_faulted;  /* replications with status != QK_REP_OK */
  uint64_t n_log_records;
  uint64_t min_now_ns, max_now_ns;
} qk_batch_summary;

typedef struct qk_model qk_model;
typedef struct qk_batch qk_batch;

/* ---- model description (host only) ---------------------------------------------------------- */
int qk_abi_version(void);
/* identifies the sources and compiler flags this library was built from (quokkasim_b200/build.py: source_id) */
const char* qk_build_id(void);
int qk_model_create(qk_model** out);
void qk_model_destroy(qk_model*);
/* registration order = component id = scheduler origin rank - 1 */
int qk_model_add_component(qk_model*, const qk_component_desc*, uint32_t* out_id);
/* which: 0 = process_time_distr, 1 = process_quantity_distr; out_dist_id identifies the instance for tapes */
int qk_model_set_dist(qk_model*, uint32_t comp, uint32_t which, const qk_dist_desc*, uint32_t* out_dist_id);
int qk_model_add_delay_mode(qk_model*, uint32_t comp, const qk_dist_desc* until_delay, const qk_dist_desc* until_fix,
                            uint32_t* out_until_delay_id, uint32_t* out_until_fix_id);
/* port_n = -1 for None */
int qk_model_connect(qk_model*, uint32_t a, uint32_t b, int32_t port_n);
int qk_model_set_logged(qk_model*, uint32_t comp, int enabled);
/* vector components. dim = 1 (f64) or 3 (Vector3). VectorStock: low/max capacity + initial resource (with_low_capacity,
 * with_max_capacity, with_initial_resource); VectorSource: init = source_vector; other kinds: dim only */
int qk_model_set_vector(qk_model*, uint32_t comp, uint32_t dim, double low_capacity, double max_capacity,
                        const double* init);
/* user-defined resource types (quokkasim_examples/src/bin/iron_ore.rs:21-96: `IronOre`, five f64 fields,
 * `total() = fe + other_elements`): dim in 1..QK_VEC_MAX_DIM fields; ResourceTotal sums the fields named by
 * total_mask (bit k = field k) in field order; ResourceAdd is field-wise; ResourceRemove(q) takes the proportion
 * q / total() of EVERY field.  init[dim].  (dim 1 = f64, dim 3 with mask 7 = Vector3: the call above.)  Components
 * connect only when dim and mask agree.  In the log a vector of more than three fields takes ceil(dim / 3)
 * continuation records: QK_LOG_VEC_DATA with index byte = vector index + 8 * chunk. */
#define QK_VEC_MAX_DIM 8
int qk_model_set_vector_n(qk_model*, uint32_t comp, uint32_t dim, uint32_t total_mask, double low_capacity,
                          double max_capacity, const double* init);
/* VectorCombiner<M> / VectorSplitter<N>: number of ports (1..5) and split_ratios (vector.rs:756, 1056) */
int qk_model_set_ports(qk_model*, uint32_t comp, uint32_t n_ports, const double* split_ratios);
/* create_scheduled_event!(SetLowCapacity(v)) on an F64Stock (core.rs:1382-1386): setter + add(0.) at t */
int qk_model_schedule_low_capacity(qk_model*, uint64_t t_ns, uint32_t stock, double value);
int qk_model_schedule_env_state(qk_model*, uint64_t t_ns, uint32_t env, uint32_t state);
/* create_scheduled_event!(SetProcessQuantity(cfg)) on an F64Process (core.rs:1387-1391): `dist` is the instance
 * DistributionFactory::create made when the event was scheduled (its seed in `stream`); at t it replaces
 * process_quantity_distr and the process gets update_state(SCH_000000).  *out_dist_id identifies it for tapes. */
int qk_model_schedule_process_quantity(qk_model*, uint64_t t_ns, uint32_t process, const qk_dist_desc* dist,
                                       uint32_t* out_dist_id);
/* Vector3ContainerFactory of a QK_CONTAINER_SOURCE (vector_container.rs:126-156): create_quantity_distr (NULL = None:
 * containers are created empty), create_component_split[3], capacity.  *out_dist_id = the distribution instance
 * (for tapes) or 0xFFFFFFFF */
int qk_model_set_container_factory(qk_model*, uint32_t source, const qk_dist_desc* create_quantity, const double* split3,
                                   double capacity, uint32_t* out_dist_id);
/* with_initial_resource(ItemDeque<Vector3Container>) of a QK_CONTAINER_STOCK: n containers in FIFO order with
 * resources[n][3] (ignored where has_resource[k] == 0), capacities[n].  Replaces n_initial_items of the descriptor. */
int qk_model_set_initial_containers(qk_model*, uint32_t stock, uint32_t n, const double* resources,
                                    const uint8_t* has_resource, const double* capacities);
int qk_model_n_components(const qk_model*, uint32_t* out);
int qk_model_n_dists(const qk_model*, uint32_t* out);
/* bytes of on-chip state one replication needs (for capacity planning / roofline arithmetic) */
int qk_model_state_bytes(const qk_model*, int with_log, uint32_t* out_bytes);

/* ---- batch of replications on one GPU -------------------------------------------------------- */
/* State placement. Default: per-replication state is staged in shared memory for the whole launch, one thread per
 * replication, when at least 12 warps of replications fit per SM that way; larger models get a lane group per
 * replication (state still in shared memory), and models too large even for that operate directly on the HBM state
 * planes (through L1/L2).  The flags force one placement. */
#define QK_BATCH_STATE_IN_HBM 1u
#define QK_BATCH_STATE_ON_CHIP 2u
/* large models: a group of `lanes_per_replication` lanes (4, 8 or 16; 0 = choose) owns one replication, whose state
 * stays in shared memory: the scheduler pop and the emit_change broadcast are done by the group's lanes side by side */
#define QK_BATCH_STATE_LANE_GROUP 4u
/* with QK_BATCH_STATE_LANE_GROUP: the phased variant -- the lane groups do the scheduler pop and the subscriber broadcast,
 * then one thread per replication runs the full handlers of the whole block side by side (qk_group.cuh) */
#define QK_BATCH_GROUP_PHASED 8u
/* large models, one thread per replication: only the scheduler queue (the fire slots), the stocks' state words and the
 * event-id counters are staged in shared memory; the component records (countdowns, flags, items in process, pre-drawn
 * durations, draw counters, resources) stay in global memory and are read and written in place, through L2, by the
 * handlers alone -- the scheduler pop and the emit_change broadcast never leave the chip */
#define QK_BATCH_STATE_SPLIT 16u
/* with QK_BATCH_STATE_SPLIT: a two-tier scheduler queue.  Only the stocks' pending emit_change events (due one nanosecond
 * after they are scheduled) are kept in shared memory; the keyed events of the process-like components live in a global
 * array by rank, summarised on chip by a busy mask and a cached minimum that is re-established by a scan of that array
 * only after the minimum itself has fired.  A third of the shared memory per replication: three times the warps. */
#define QK_BATCH_QUEUE_TWO_TIER 32u
typedef struct {
  uint64_t n_reps;
  uint64_t rep_offset;   /* global index of this batch's first replication (Philox counter words 0,1) */
  uint64_t base_seed;    /* Philox key */
  int32_t device;        /* CUDA device ordinal */
  uint32_t log_capacity; /* records kept per replication (ring, power of two); 0 = logging off */
  uint32_t threads_per_block; /* 0 = choose */
  uint32_t flags;        /* QK_BATCH_* bits */
  void* stream;          /* cudaStream_t to launch on, NULL = the library creates its own */
  uint32_t lanes_per_replication; /* ABI 2: with QK_BATCH_STATE_LANE_GROUP: 4, 8 or 16 ; 0 = choose */
  uint32_t reserved;
} qk_batch_config;

/* launch geometry and memory footprint chosen for a batch (for roofline arithmetic and capacity planning) */
typedef struct {
  uint32_t threads_per_block, grid_blocks, shared_bytes_per_block, table_bytes;
  uint32_t slots64, slots32;         /* per-replication state: slots64 * 8 + slots32 * 4 bytes */
  uint32_t state_bytes_per_rep;      /* HOT state: loaded from / stored to HBM once per launch */
  uint32_t resident_blocks_per_sm;   /* occupancy of the step kernel as reported by the CUDA runtime */
  uint32_t state_in_hbm;             /* 1 if the batch operates on the HBM planes in place */
  uint32_t cold_bytes_per_rep;       /* COLD state (item rings, source ids of pending events): global memory only */
  uint64_t replications_padded;
  uint64_t state_bytes_total, log_bytes_total;
  /* ABI 2 */
  uint32_t lanes_per_replication; /* 1: one thread per replication ; 4/8/16: a lane group per replication */
  uint32_t wave_blocks;           /* blocks of the step kernel resident on the whole GPU at once */
  uint32_t last_step_chunks;      /* state residencies of the last step call: 1, or the chunks of a sliced step */
  uint32_t state_split;           /* 1: QK_BATCH_STATE_SPLIT placement (state_bytes_per_rep = the staged part) */
} qk_batch_info;

int qk_batch_create(const qk_model*, const qk_batch_config*, qk_batch** out);
int qk_batch_info_get(qk_batch*, qk_batch_info* out);
void qk_batch_destroy(qk_batch*);
/* pre-drawn variate tape for one distribution instance: values[rep * per_rep_len + i], host pointer, copied */
int qk_batch_set_tape(qk_batch*, uint32_t dist_id, const double* values, uint64_t per_rep_len);
/* Model::init of every component at t0 (registration order) */
int qk_batch_init(qk_batch*, uint64_t t0_ns);
int qk_batch_step_until(qk_batch*, uint64_t t_ns);
int qk_batch_step_events(qk_batch*, uint64_t n_events);
/* same result as qk_batch_step_events, cut into launches of `chunk` events (state round-trips HBM between them);
 * for measuring the kernel at a small number of events per state residency */
int qk_batch_step_events_chunked(qk_batch*, uint64_t n_events, uint64_t chunk);
/* same result again, as the sliced schedule qk_batch_step_events chooses by itself when a launch would end in a
 * mostly empty last wave: the step is cut into n_chunks chunks of events and the blocks of replications into
 * parts, each part advanced chunk by chunk on its own stream so that the launches overlap and the GPU stays full
 * (n_chunks = 0: let the library decide) */
int qk_batch_step_events_sliced(qk_batch*, uint64_t n_events, uint32_t n_chunks);
int qk_batch_synchronize(qk_batch*);

/* ---- results ------------------------------------------------------------------------------------ */
int qk_batch_summary_get(qk_batch*, qk_batch_summary* out);
int qk_batch_read_status(qk_batch*, uint64_t rep0, uint64_t n, uint32_t* status, uint64_t* now_ns, uint64_t* n_events);
int qk_batch_comp_stats(qk_batch*, uint64_t rep0, uint64_t n, uint32_t comp, qk_comp_stats* out);
int qk_batch_reduce_stats(qk_batch*, qk_comp_summary* out, uint32_t n_comp);
/* the packed partials of this batch on the host: partials[n_comp][QK_PARTIAL_WORDS] */
int qk_batch_reduce_partials(qk_batch*, uint64_t* partials, uint32_t n_comp);
/* device-resident variants for multi-GPU callers that own the collective (quokkasim_b200/parallel.py; qk_multi_* below
 * does the same inside the library): fill a caller-provided DEVICE array with this batch's packed partials, enqueued on
 * the batch's stream ... */
int qk_batch_reduce_partials_device(qk_batch*, uint64_t* d_partials, uint32_t n_comp);
/* ... combine n_parts such arrays laid end to end (the result of an all-gather) into d_out (may alias none of them) */
int qk_batch_combine_partials_device(qk_batch*, const uint64_t* d_parts, uint32_t n_parts, uint64_t* d_out,
                                     uint32_t n_comp);
/* ... and bin this batch's replications over the GLOBAL ranges found in the combined partials:
 * d_hist [n_comp][2][QK_HIST_BINS] (time_ns, count); the histograms of the ranks then add */
int qk_batch_histograms_device(qk_batch*, const uint64_t* d_partials, uint64_t* d_hist, uint32_t n_comp);
/* records of one replication, oldest first (at most log_capacity newest); *n_total = records ever written */
int qk_batch_read_log(qk_batch*, uint64_t rep, qk_log_record* buf, uint64_t cap, uint64_t* n_out, uint64_t* n_total);
/* the same for replications [rep0, rep0 + n) with ONE gather kernel and one copy: buf[n][log_capacity] (the first
 * kept[i] records of row i are valid, oldest first), kept[n], total[n] */
int qk_batch_read_logs(qk_batch*, uint64_t rep0, uint64_t n, qk_log_record* buf, uint32_t* kept, uint64_t* total);
/* resource of a vector stock / in-flight vector of a vector process-like component (3 doubles) */
int qk_batch_read_vector(qk_batch*, uint64_t rep, uint32_t comp, double* out3);
/* ... of any resource type: out[cap >= dim], *dim_out = the number of fields (qk_model_set_vector_n) */
int qk_batch_read_vector_n(qk_batch*, uint64_t rep, uint32_t comp, double* out, uint32_t cap, uint32_t* dim_out);
/* discrete stock contents in FIFO order */
int qk_batch_read_stock(qk_batch*, uint64_t rep, uint32_t comp, uint32_t* items, uint64_t cap, uint64_t* n_out);
/* containers held by a container stock (FIFO order) or by a container process-like component (in progress first, then
 * completed-but-not-pushed): handles[cap], resources[cap][3] as raw fp64 bits (first word QK_CONTAINER_NONE_BITS for a
 * container without resource) */
int qk_batch_read_containers(qk_batch*, uint64_t rep, uint32_t comp, uint32_t* handles, uint64_t* resources,
                             uint64_t cap, uint64_t* n_out);
/* timing of the last step call measured with CUDA events on the batch's stream (ms) and kernels launched */
int qk_batch_last_step_ms(qk_batch*, float* ms, uint32_t* n_launches);

/* ---- one job on several GPUs of one box (`devices, n_dev`, statistics "after internal NCCL reduce") ----------
 * The replications [0, n_reps) of the job are sharded over the devices by contiguous range (device d of D owns
 * [n_reps*d/D, n_reps*(d+1)/D), = its Philox counter range); each shard is a qk_batch on its own stream, driven from
 * the calling thread (every step call only enqueues).  Replaces running one `Simulation` per replication
 * (quokkasim_examples/src/bin/discrete_queue.rs:109-111) on a box with several GPUs; there is no data-path collective.
 * qk_multi_stats is the one exchange: packed partials -> ncclAllGather -> combine -> global-range histograms ->
 * ncclAllReduce(sum); it returns the same qk_comp_summary a single batch holding all the replications would.
 * NCCL (libnccl.so.2) is loaded at run time, and only when a job has more than one device. */
typedef struct qk_multi qk_multi;
typedef struct {
  uint64_t n_reps;        /* replications of the whole job */
  uint64_t rep_offset;    /* global index of the job's first replication */
  uint64_t base_seed;
  const int32_t* devices; /* CUDA ordinals, n_devices of them; NULL / 0 = every visible device */
  uint32_t n_devices;
  uint32_t log_capacity, threads_per_block, flags, lanes_per_replication; /* as qk_batch_config, for every shard */
} qk_multi_config;
int qk_multi_create(const qk_model*, const qk_multi_config*, qk_multi** out);
void qk_multi_destroy(qk_multi*);
int qk_multi_n_shards(qk_multi*, uint32_t* n);
/* shard i: its batch (for the per-replication read-back calls qk_batch_read_*), first replication, count, device */
int qk_multi_shard(qk_multi*, uint32_t i, qk_batch** batch, uint64_t* rep0, uint64_t* n_reps, int32_t* device);
/* values[rep * per_rep_len + i] for the replications of the whole job */
int qk_multi_set_tape(qk_multi*, uint32_t dist_id, const double* values, uint64_t per_rep_len);
int qk_multi_init(qk_multi*, uint64_t t0_ns);
int qk_multi_step_until(qk_multi*, uint64_t t_ns);
int qk_multi_step_events(qk_multi*, uint64_t n_events);
int qk_multi_synchronize(qk_multi*);
/* device-timed duration of the last step call: the maximum over the GPUs */
int qk_multi_last_step_ms(qk_multi*, float* ms_max);
/* globally reduced summary: out[n_comp]; *total (optional) = whole-job counts and clock range */
int qk_multi_stats(qk_multi*, qk_comp_summary* out, uint32_t n_comp, qk_batch_summary* total);

const char* qk_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* QUOKKA_B200_H */
