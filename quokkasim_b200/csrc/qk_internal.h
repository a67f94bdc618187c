/*
 * qk_internal.h -- lowered model tables and per-replication state layout shared by the host
 * lowering (qk_model.cpp) and the sm_100a kernels (qk_kernels.cu).  Not part of the public ABI.
 *
 * Per-replication state is two planes of slots:
 *     plane64: u64 slots   (times, countdowns, integrals)
 *     plane32: u32 slots   (flags, counters, item rings)
 * On chip a block of NT replications keeps them in shared memory replication-fastest
 * (slot * NT + tid): whatever slot a lane touches, lane i hits bank i, so data-dependent
 * component indices never cause bank conflicts.  In HBM the same planes are stored
 * [slot][replication] so block load/store is fully coalesced.
 */
#ifndef QK_INTERNAL_H
#define QK_INTERNAL_H
#include <stdint.h>

#include "../../include/quokka_b200.h"

#define QK_TIME_NONE 0xFFFFFFFFFFFFFFFFull
#define QK_MAX_DELAY_MODES 8
#define QK_MAX_COMPONENTS 1024

/* ---- global (per replication) slots ------------------------------------------------------------ */
enum { G64_NOW = 0, G64_NEVENTS = 1, G64_FIRE0 = 2 }; /* fire[i] for i in [0, n_sched) follows */
enum { G32_STATUS = 0, G32_EXT_CURSOR = 1, G32_LOGCNT = 2, G32_STALE0 = 3 }; /* stale mask words follow, then
                                                                              the prefetch-pending mask words */

/* pre-drawn process durations (deferred variate prefetch): a u64 slot per prefetchable distribution holds the
 * NEXT duration in ns, already converted with Duration::from_secs_f64, or one of these codes */
#define QK_DUR_EMPTY 0xFFFFFFFFFFFFFFFFull     /* consumed, waiting for the refill phase */
#define QK_DUR_FAULT_BASE 0xFFFFFFFFFFFFFF00ull /* QK_DUR_FAULT_BASE + qk_rep_status: raise when consumed */

/* ---- process-like slots, relative to DevComp.o64 / o32 ----------------------------------------- */
enum { P64_LEFT = 0, P64_PREV = 1, P64_BASE_COUNT = 2 }; /* optional: ttnpe, nextdur, dm durations */
enum { P32_FLAGS = 0, P32_ITEM = 1, P32_BASE_COUNT = 2 }; /* optional: next item index */
/* flags */
#define PF_HAS_ITEM 1u
#define PF_ENV_STOPPED 2u
#define PF_FIX_SHIFT 8 /* bits 8..15: delay mode k is in TimeUntilFix */
/* LOG-only slots: the seq counter at hot slot olog32 ; the source ids of the pending / stale keyed event at COLD64
 * slots olog64 + PL64_* */
enum { PL64_SCHED_SRC = 0, PL64_STALE_SRC = 1, PL64_COUNT = 2 };
enum { L32_SEQ = 0 };

/* ---- cold planes ----------------------------------------------------------------------------------
 * State that is touched once or twice per item and never decides control flow -- stock item rings, the source
 * ids carried by pending events -- lives only in global memory ([slot][replication], like the hot planes in HBM)
 * and is read and written in place through L1/L2.  Keeping it out of shared memory is what lets 20 warps of
 * replications stay resident per SM instead of 12-16 (the step kernel's throughput is proportional to resident
 * warps: profiles/README.md "occupancy experiment"). */

/* ---- discrete stock slots ------------------------------------------------------------------------ */
enum { S64_COUNT = 0 }; /* the fire slot holds emit_t0; nothing else is 64-bit */
enum { S32_HC = 0, S32_EMIT = 1, S32_COUNT = 2, S32_EMITSRC = 2, S32_HEADITEM = 3 };
/* S32_HC = ring head (15 bits) | item count (15 bits) | state category (2 bits: 0 Empty, 1 Normal, 2 Full).  The
 * category is a function of the count (get_state, discrete.rs:161-171) and is cached in the word so that the many
 * "what state is my neighbour in" questions of the event loop are ONE state load -- DevHotB.up_hc / down_hc name this
 * word directly -- instead of a table quad, a state load and two compares. */
#define QK_HC_HEAD(hc) ((hc)&0x7FFFu)
#define QK_HC_CNT(hc) (((hc) >> 15) & 0x7FFFu)
#define QK_HC_CAT(hc) ((hc) >> 30)
#define QK_HC_PACK(head, cnt, cat) ((head) | ((cnt) << 15) | ((cat) << 30))
#define QK_HC_NONE 0xFFFFu /* up_hc / down_hc of a port nobody connected */
/* LOG only, discrete stocks: S32_EMITSRC = source id of the oldest pending emit_change ; S32_HEADITEM = copy of the
 * ring's head item (on-chip kernels; refilled by an asynchronous global->shared copy, see qk_stock_remove) */
/* S32_EMIT = n | split << 8 | head << 16.
 * COLD state (global memory only, never staged in shared memory -- see "cold planes" below): the item ring
 * [ring_cap] at cold32 slot `ocold32`, followed (LOG) by the emit source-seq ring [emit_depth] */

/* ---- parallel process slots ---------------------------------------------------------------------- */
enum { PP64_PREV = 0, PP64_BASE_COUNT = 1 }; /* then dm durs, in-progress durations[cap] */
enum { PP32_FLAGS = 0, PP32_NPROG = 1, PP32_COMPLETE = 2, PP32_BASE_COUNT = 3 };
/* then in-progress items[cap], complete ring[cap] */

/* ---- statistics planes ----------------------------------------------------------------------------
 * Accumulators are NOT part of the on-chip state: they live only in HBM ([comp * 2 + k][replication]) and are
 * updated with fire-and-forget RED atomics (the working set of the resident blocks stays in L2).
 *   stat64[0]: process-like: sum of started durations (busy ns = this - remaining) ; stock: occupancy integral
 *   stat64[1]: process-like: ns spent in an active delay
 *   stat32[0]: process-like: ProcessFinish count ; stock: Add count
 *   stat32[1]: stock: maximum occupancy */
enum { ST64_TIME = 0, ST64_DELAY = 1, ST64_FQ = 2, ST32_COUNT = 0, ST32_MAXOCC = 1, ST_PER_COMP = 3 };
/* vector components: stat64[ST64_TIME] of a VectorStock holds the fp64 integral of its total (unit*s);
 * stat64[ST64_FQ] of a vector process-like component holds the fp64 sum of its success quantities */

/* ---- environment --------------------------------------------------------------------------------- */
enum { E32_STATE = 0, E32_COUNT = 1 };

#define DC_LOGGED 1u
#define DC_SIMPLE 2u /* single-server discrete component without delay modes or environment (quick path) */
#define DC_CONTAINER 4u /* the items of this component are Vector3Containers: every slot that holds an item handle has
                           three cold64 slots (from DevComp.ocold64) holding what the container carries */

/* Container family (vector_container.rs).  Vector3ContainerStock / Process / Source / Sink / ParallelProcess are
 * lowered to the DISCRETE kinds with DC_CONTAINER set (they are the discrete components instantiated with another item
 * type, core.rs:549-553); ContainerLoadingProcess / ContainerUnloadingProcess keep their own kinds and share the
 * multi-server layout of the parallel process (slots below), with max_process_count in DevComp.max / ring_cap.
 * Payload layout at cold64 slot ocold64: stock -> [ring_cap][3]; single-server process-like -> [3];
 * multi-server -> in-progress [ring_cap][3], then the complete ring [ring_cap][3]. */

typedef struct {
  uint8_t kind;    /* qk_component_kind */
  uint8_t n_dm;    /* delay modes */
  uint8_t n_subs;  /* subscribers of state_emitter / emit_change */
  uint8_t src_ord; /* source ordinal for item handles */
  int16_t up, down, env;
  uint16_t subs_off;
  uint32_t low, max;
  uint16_t o64, o32;
  uint16_t olog64, olog32;
  uint16_t fire_idx;  /* index into fire[] / stale mask, 0xFFFF if not schedulable */
  uint16_t ring_cap;  /* stock: item ring slots ; parallel: in-flight slots */
  uint16_t dist_time; /* distribution instance of process_time_distr */
  uint16_t dm_off;    /* first DevDelayMode */
  uint16_t o64_dm;    /* first delay-mode duration slot (absolute) */
  uint16_t o64_ttnpe; /* absolute slot, 0xFFFF if none (only DiscreteSink keeps ttnpe across calls) */
  uint16_t vec_idx; /* index into DevVec[] for vector components */
  uint16_t o32_next_item; /* absolute slot (sources) */
  uint16_t emit_depth;
  uint16_t flags;
  uint16_t o64_nextdur; /* absolute slot of the pre-drawn next process duration, 0xFFFF if none */
  uint16_t pf_idx;      /* bit index in the prefetch-pending mask */
  uint32_t const_dur_lo, const_dur_hi; /* process_time_distr is Constant: its duration in ns (or QK_DUR_FAULT code) */
  uint16_t ocold32; /* cold32 slot: discrete stock item ring (+ emit source ring) ; vector stock emit source ring */
  uint16_t ocold64; /* DC_CONTAINER: first cold64 payload slot */
} DevComp; /* 68 bytes = 17 words: an odd word stride keeps per-lane table reads of different components off the
              same shared-memory bank */

/* ---- hot component rows --------------------------------------------------------------------------
 * What the event loop reads of a component on every call, packed into three 16-byte quads so that one LDS.128 per
 * quad brings it into registers (per-field LDS.U16 reads of DevComp cost one shared-memory wavefront each AND had to
 * be repeated after every state store, because tables and state share one shared-memory array).  48-byte stride:
 * quad index 3c mod 8 is a permutation for 8 consecutive components, so lanes that work on different components
 * read different bank groups. */
typedef struct {
  uint8_t kind, n_dm, n_subs, src_ord;
  uint16_t o32, o64;
  uint16_t fire_idx, flags; /* DC_LOGGED */
  uint16_t olog32;
  uint16_t olog64; /* process-like, vector stock: cold64 slot (LOG) ; discrete stock: cold32 slot of the item ring */
} DevHotA; /* every kind */
typedef struct {
  uint16_t up_hc, down_hc; /* single-server discrete components: plane32 slot of the S32_HC word of the upstream /
                              downstream stock (QK_HC_NONE if not connected); the stock ids are in DevComp.up / down */
  int16_t env;
  uint16_t dist_time;
  uint16_t o64_dm, dm_off;
  uint16_t o64_nextdur, pf_idx;
} DevHotB; /* process-like */
typedef struct {
  uint32_t low, max;
  uint16_t ring_cap, emit_depth;
  uint16_t subs_off, o32;
} DevHotSB; /* stocks and BasicEnvironment (which only uses subs_off / o32) */
typedef struct {
  uint16_t o64_ttnpe, o32_next_item;
  uint32_t const_dur_lo, const_dur_hi;
  uint16_t vec_idx, ring_cap;
} DevHotC; /* process-like, second quad */
typedef struct {
  DevHotA a;
  union {
    DevHotB b;
    DevHotSB sb;
  };
  DevHotC c;
} DevHot; /* 48 bytes */

/* One per entry of subs[] (same index): what the quick path of the event loop needs to know about the callee, in one
 * 8-byte load -- where its fire slot is and which state words tell whether it could start -- instead of walking
 * subs[] -> hot row A -> fire slot -> hot row B -> neighbour words one dependent shared-memory access after the other. */
typedef struct {
  uint16_t fire_slot;        /* absolute plane64 slot of the callee's fire time */
  uint16_t up_hc, down_hc;   /* plane32 slots of the S32_HC words of its upstream / downstream stock (0 if forced) */
  uint16_t flags;            /* SQ_* */
} DevSubQ;
#define SQ_QUICK 1u      /* the callee is a SIMPLE component: the quick path applies */
#define SQ_US_FORCED 2u  /* upstream category is the constant in bits 4-5 (1: a source needs none ; 3: not connected) */
#define SQ_DS_FORCED 4u  /* downstream category is the constant in bits 6-7 */
#define SQ_US_SHIFT 4
#define SQ_DS_SHIFT 6
#define SQ_DM 512u       /* the callee is a single-server discrete component WITH delay modes (no environment, no
                            container items): not SIMPLE, but two of its calls still change nothing that needs the full
                            handler -- idle and still blocked (as for SIMPLE components), and busy with neither the
                            process countdown nor any until-delay countdown running out (qk_dm_quick) */
#define SQ_DUP 256u      /* the callee appears more than once in this call list (a process whose upstream and downstream
                            are the same stock): the lane-group kernel, which lets the lanes of a group write the
                            ProcessNonStart records of a list side by side, takes such entries one at a time */

typedef struct {
  uint32_t kind, stream;
  double p[4];
  uint32_t draw_slot; /* absolute plane32 slot of the draw counter */
  uint32_t pad;
} DevDist; /* 48 bytes */

typedef struct {
  uint16_t until_delay, until_fix; /* distribution instance ids */
} DevDelayMode;

typedef struct {
  uint64_t t;
  uint32_t type; /* 0 SetEnvironmentState, 1 SetLowCapacity (F64Stock), 2 SetProcessQuantity (F64Process): `state` = the
                    distribution instance created at schedule time */
  uint32_t target;
  uint32_t state;
  uint32_t pad;
  double value;
} DevExt; /* 32 bytes */

/* extra description of a vector component (vector.rs); resource type f64 (dim 1), Vector3 (dim 3) or a user-defined
 * struct of up to QK_VEC_MAX_DIM fields (feature set 3) */
typedef struct {
  double vmax;      /* VectorStock.max_capacity */
  double vinit[QK_VEC_MAX_DIM]; /* VectorStock initial resource / VectorSource.source_vector */
  double ratios[5]; /* VectorSplitter.split_ratios */
  int16_t ups[5];   /* VectorCombiner upstream stocks by port */
  int16_t downs[5]; /* VectorSplitter downstream stocks by port */
  uint16_t dim, n_ports;
  uint16_t dist_qty; /* process_quantity_distr */
  uint16_t o64_res;  /* stock: resource[dim] ; process-like: in-flight vector[dim] (+1 quantity slot for combiners) */
  uint16_t o64_low;  /* stock: run-time low_capacity (f64 bits) */
  uint16_t o64_last; /* stock: time of the last level change (for the integral) */
  double vlow;       /* VectorStock.low_capacity at t0 */
  uint16_t o32_qty;  /* F64Process that a SetProcessQuantity event targets: plane32 slot holding the distribution instance
                        process_quantity_distr currently is (0xFFFF: it never changes, dist_qty) */
  uint16_t tmask;    /* fields ResourceTotal counts (bit k = field k; Vector3: 7) */
  uint16_t pad[2];
} DevVec; /* 160 bytes */

/* header of the table blob (all offsets in bytes from the blob start, 8-byte aligned) */
typedef struct {
  uint32_t n_comp, n_sched, n_dist, n_dm, n_subs, n_ext;
  uint32_t n64, n32;         /* hot slots per replication (for the chosen LOG mode) */
  uint32_t n64c, n32c;       /* cold slots per replication */
  uint32_t n_stale_words;
  uint32_t ident_off;        /* index in subs[] of the identity list (subs[ident_off + c] == c) */
  uint32_t init_off, n_init; /* subs[init_off ..): components whose Model::init does something, registration order */
  uint32_t n_pf, n_pf_words, pf32; /* prefetchable distributions; pf32 = first pending-mask slot (plane32) */
  uint32_t off_pf;           /* u16 pf_comp[n_pf]: prefetch index -> component */
  uint32_t n_vec, off_vecs;  /* DevVec[] */
  uint32_t off_comps, off_dists, off_dms, off_subs, off_sched, off_ext;
  uint32_t total_bytes;
  uint32_t log_enabled;
  uint32_t off_hot; /* DevHot[n_comp], 16-byte aligned */
  uint32_t off_subq;  /* DevSubQ[n_subs], 8-byte aligned */
  uint32_t off_ccap;  /* double ccap[64 + n_initial]: container capacity by source ordinal, then by initial-item index */
  uint32_t off_cinit; /* uint64_t cinit[n_initial][3]: resource bits of the initial containers (NONE sentinel) */
  uint32_t features; /* 0 = every process-like component is SIMPLE and nothing is scheduled externally ; 1 = general ;
                        2 = general + container family (vector_container.rs) ; 3 = general + user-defined vector
                        resources of up to QK_VEC_MAX_DIM fields (no containers) */
  uint32_t n_hot, off_hot_rank; /* placement 2: entries of the hot fire array (stocks) ; u16 hot_rank[n_hot]: index -> rank */
  uint32_t gshift, n_groups, gmin64, gidx32, gdirty32; /* placement 2: group minima of the cold fire array (qk_lower) */
  uint32_t emask32;             /* placement 2: first hot32 slot of the pending-emit mask (bit = index in the hot fire array) */
  uint32_t cfire0, busy32;      /* placement 2: first cold64 slot of the process-like fire array (by rank, padded to a
                                   multiple of four) ; first hot32 slot of the busy mask (bit = rank) */
  uint32_t split;    /* placement. 1: component records (everything but the fire slots, the stocks' hot words and the seq counters)
                        were allocated in the COLD planes: DevComp.o64 / o32, o64_dm, o64_nextdur, o64_ttnpe,
                        o32_next_item, DevDist.draw_slot, DevVec.o64_res / o64_low / o64_last / o32_qty are cold-plane slot
                        numbers (the kernels' W64 / W32 accessors).  2: as 1, and the scheduler queue is two-tier (qk_lower) */
} DevHeader;

/* ---- lane-group kernel geometry (qk_group.cuh): R replications per block, G lanes each ----------------------------
 * Shared memory holds the block's state slot-major: 64-bit rows of stride64 elements, then 32-bit rows of stride32.
 * Rows are moved by bulk asynchronous copies, so R * 4 bytes must be a multiple of 16 and the row starts 16-byte aligned
 * (R a multiple of 4; stride64 even; stride32 a multiple of 4).  Bank spread: the G lanes of a group read G consecutive
 * slots of ONE replication, i.e. addresses one row stride apart -- stride64 = 2 (mod 4) puts 8 consecutive 64-bit
 * rows on 8 different bank pairs, stride32 = 4 (mod 8) 8 consecutive 32-bit rows on 8 different banks. */
typedef struct {
  uint32_t reps, stride64, stride32, bar_off, smem_bytes, threads;
} QkGroupGeom;
/* x64 / x32: extra rows after the state rows (the phased kernel keeps each replication's loop state there) */
static inline int qk_group_geometry(uint32_t G, uint32_t warp, uint32_t n64, uint32_t n32, uint32_t x64, uint32_t x32,
                                    uint32_t table_bytes, uint32_t max_threads, uint32_t max_smem, uint32_t reps_cap,
                                    uint32_t rep_multiple, QkGroupGeom* out) {
  uint32_t unit = G >= warp ? 1u : warp / G; /* whole warps */
  if (unit < rep_multiple) unit = rep_multiple;
  int found = 0;
  for (uint32_t R = unit; R * G <= max_threads && R <= reps_cap; R += unit) {
    uint32_t s64 = R, s32 = R;
    while ((s64 & 3u) != 2u) ++s64;
    while ((s32 & 7u) != 4u) ++s32;
    const uint32_t bar = (table_bytes + (n64 + x64) * s64 * 8u + (n32 + x32) * s32 * 4u + 7u) & ~7u;
    if (bar + 16u > max_smem) break;
    found = 1;
    out->reps = R;
    out->stride64 = s64;
    out->stride32 = s32;
    out->bar_off = bar;
    out->smem_bytes = bar + 16u;
    out->threads = R * G;
  }
  return found;
}

#endif
