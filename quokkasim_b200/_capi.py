"""ctypes binding of include/quokka_b200.h (the C-ABI shared library with the sm_100a kernels).

There is no CPU fallback: if libquokka_b200.so is missing this module raises at load time, and every
batch call fails loudly when no CUDA device is present (QK_ERR_NO_DEVICE).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libquokka_b200.so")

# qk_component_kind
DISCRETE_STOCK, DISCRETE_PROCESS, DISCRETE_SOURCE, DISCRETE_SINK, DISCRETE_PARALLEL_PROCESS, ENVIRONMENT = 1, 2, 3, 4, 5, 6
VECTOR_STOCK, VECTOR_PROCESS, VECTOR_SOURCE, VECTOR_SINK, VECTOR_COMBINER, VECTOR_SPLITTER = 16, 17, 18, 19, 20, 21
CONTAINER_STOCK, CONTAINER_PROCESS, CONTAINER_SOURCE, CONTAINER_SINK = 32, 33, 34, 35
CONTAINER_PARALLEL_PROCESS, CONTAINER_LOAD_PROCESS, CONTAINER_UNLOAD_PROCESS = 36, 37, 38
CONTAINER_NONE_BITS = 0x7FF40000000000C0
# qk_dist_kind
DIST_CONSTANT, DIST_UNIFORM, DIST_TRIANGULAR, DIST_NORMAL, DIST_TRUNCNORMAL, DIST_EXPONENTIAL = 0, 1, 2, 3, 4, 5
# qk_log_kind
LOG_PROCESS_START, LOG_PROCESS_CONTINUE, LOG_PROCESS_FINISH, LOG_PROCESS_NONSTART = 1, 2, 3, 4
LOG_PROCESS_STOPPED, LOG_WITHDRAW_REQUEST, LOG_DELAY_START, LOG_DELAY_END = 5, 6, 7, 8
LOG_STOCK_ADD, LOG_STOCK_REMOVE, LOG_STOCK_STATE_CHANGE, LOG_ENV_STATE = 16, 17, 18, 24
LOG_VSTOCK_ADD, LOG_VSTOCK_REMOVE, LOG_VSTOCK_STATE_CHANGE = 32, 33, 34
LOG_VPROC_START, LOG_VPROC_SUCCESS, LOG_VPROC_FAILURE = 40, 41, 42
LOG_VPROC_COMBINE_START, LOG_VPROC_COMBINE_SUCCESS, LOG_VPROC_SPLIT_START, LOG_VPROC_SPLIT_SUCCESS = 43, 44, 45, 46
LOG_VEC_DATA = 63
SRC_INIT, SRC_SCH = 0xFFFF, 0xFFFE
ITEM_NONE = 0xFFFFFFFFFFFFFFFF
HIST_BINS = 32
TIME_NONE = 0xFFFFFFFFFFFFFFFF

STATUS_NAMES = {
    0: "QK_OK", -1: "QK_ERR_INVALID_ARG", -2: "QK_ERR_UNSUPPORTED_CONNECTION", -3: "QK_ERR_PORT_IN_USE",
    -4: "QK_ERR_NO_DEVICE", -5: "QK_ERR_CUDA", -6: "QK_ERR_STATE", -7: "QK_ERR_CAPACITY",
    -8: "QK_ERR_BUFFER_TOO_SMALL",
}

# every symbol include/quokka_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "qk_abi_version", "qk_build_id", "qk_model_create", "qk_model_destroy", "qk_model_add_component", "qk_model_set_dist",
    "qk_model_add_delay_mode", "qk_model_connect", "qk_model_set_logged", "qk_model_schedule_env_state",
    "qk_model_set_vector", "qk_model_set_vector_n", "qk_batch_read_vector_n", "qk_model_set_ports", "qk_model_schedule_low_capacity", "qk_model_schedule_process_quantity", "qk_batch_read_vector",
    "qk_model_n_components", "qk_model_n_dists", "qk_model_state_bytes", "qk_batch_create", "qk_batch_destroy",
    "qk_batch_info_get",
    "qk_batch_set_tape", "qk_batch_init", "qk_batch_step_until", "qk_batch_step_events", "qk_batch_step_events_chunked", "qk_batch_step_events_sliced",
    "qk_batch_synchronize",
    "qk_batch_summary_get", "qk_batch_read_status", "qk_batch_comp_stats", "qk_batch_reduce_stats",
    "qk_batch_reduce_partials", "qk_batch_reduce_partials_device", "qk_batch_combine_partials_device",
    "qk_batch_histograms_device", "qk_batch_read_log", "qk_batch_read_logs", "qk_batch_read_stock",
    "qk_batch_last_step_ms", "qk_last_error",
    "qk_model_set_container_factory", "qk_model_set_initial_containers", "qk_batch_read_containers",
    "qk_multi_create", "qk_multi_destroy", "qk_multi_n_shards", "qk_multi_shard", "qk_multi_set_tape", "qk_multi_init",
    "qk_multi_step_until", "qk_multi_step_events", "qk_multi_synchronize", "qk_multi_last_step_ms", "qk_multi_stats",
]


class QkError(RuntimeError):
    """A failed C-ABI call; .status is the qk_status, the message is qk_last_error()."""

    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), msg))
        self.status = status
        self.detail = msg


class DistDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("stream", C.c_uint32), ("p", C.c_double * 4)]


class ComponentDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("low_capacity", C.c_uint32), ("max_capacity", C.c_uint32),
        ("n_initial_items", C.c_uint32), ("capacity_hint", C.c_uint32), ("initial_env_state", C.c_uint32),
    ]


class BatchConfig(C.Structure):
    _fields_ = [
        ("n_reps", C.c_uint64), ("rep_offset", C.c_uint64), ("base_seed", C.c_uint64), ("device", C.c_int32),
        ("log_capacity", C.c_uint32), ("threads_per_block", C.c_uint32), ("flags", C.c_uint32), ("stream", C.c_void_p),
        ("lanes_per_replication", C.c_uint32), ("reserved", C.c_uint32),
    ]


class MultiConfig(C.Structure):
    _fields_ = [
        ("n_reps", C.c_uint64), ("rep_offset", C.c_uint64), ("base_seed", C.c_uint64), ("devices", C.POINTER(C.c_int32)),
        ("n_devices", C.c_uint32), ("log_capacity", C.c_uint32), ("threads_per_block", C.c_uint32), ("flags", C.c_uint32),
        ("lanes_per_replication", C.c_uint32),
    ]


class BatchSummary(C.Structure):
    _fields_ = [
        ("n_reps", C.c_uint64), ("n_events", C.c_uint64), ("n_faulted", C.c_uint64), ("n_log_records", C.c_uint64),
        ("min_now_ns", C.c_uint64), ("max_now_ns", C.c_uint64),
    ]


class BatchInfo(C.Structure):
    _fields_ = [
        ("threads_per_block", C.c_uint32), ("grid_blocks", C.c_uint32), ("shared_bytes_per_block", C.c_uint32),
        ("table_bytes", C.c_uint32), ("slots64", C.c_uint32), ("slots32", C.c_uint32),
        ("state_bytes_per_rep", C.c_uint32), ("resident_blocks_per_sm", C.c_uint32),
        ("state_in_hbm", C.c_uint32), ("cold_bytes_per_rep", C.c_uint32),
        ("replications_padded", C.c_uint64), ("state_bytes_total", C.c_uint64), ("log_bytes_total", C.c_uint64),
        ("lanes_per_replication", C.c_uint32), ("wave_blocks", C.c_uint32), ("last_step_chunks", C.c_uint32),
        ("state_split", C.c_uint32),
    ]


LOG_DTYPE = np.dtype(
    [("time_ns", "<u8"), ("comp", "<u2"), ("kind", "u1"), ("reason", "u1"), ("seq", "<u4"), ("src_comp", "<u2"),
     ("aux", "<u2"), ("src_seq", "<u4"), ("payload", "<u8")]
)
COMP_STATS_DTYPE = np.dtype([("count", "<u8"), ("time_ns", "<u8"), ("aux", "<u8"), ("level", "<u8"), ("fvalue", "<f8"),
                             ("faux", "<f8")])
COMP_SUMMARY_DTYPE = np.dtype(
    [("sum_count", "<u8"), ("sum_time_lo", "<u8"), ("sum_time_hi", "<u8"), ("sum_level", "<u8"),
     ("min_count", "<u8"), ("max_count", "<u8"), ("min_time", "<u8"), ("max_time", "<u8"),
     ("sumsq_time", "<f8"), ("sumsq_count", "<f8"), ("hist_time", "<u8", (HIST_BINS,)),
     ("hist_count", "<u8", (HIST_BINS,)), ("hist_lo_time", "<u8"), ("hist_hi_time", "<u8"),
     ("hist_lo_count", "<u8"), ("hist_hi_count", "<u8"), ("sumsq_time_ns2", "<u8", (3,)),
     ("sumsq_count_lo", "<u8"), ("sumsq_count_hi", "<u8"), ("sumsq_level", "<u8")]
)
# packed partials (quokka_b200.h QK_PARTIAL_WORDS / QP_*)
PARTIAL_WORDS = 20
(QP_COUNT, QP_TIME_LO, QP_TIME_HI, QP_LEVEL, QP_NEVENTS, QP_NFAULTED, QP_MIN_COUNT, QP_MIN_TIME, QP_NMAX_COUNT,
 QP_NMAX_TIME, QP_AA_LO, QP_AA_HI, QP_AB_LO, QP_AB_HI, QP_BB_LO, QP_BB_HI, QP_CC_LO, QP_CC_HI, QP_LEVEL_SQ,
 QP_NREPS) = range(20)
assert LOG_DTYPE.itemsize == 32 and COMP_STATS_DTYPE.itemsize == 48

_LIBS = {}


def load_library(path=None):
    """Load the C-ABI library (default: the in-tree libquokka_b200.so) and declare its signatures."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise ImportError(
            "%s not found: build it with `python -m quokkasim_b200.build` (nvcc, sm_100a). "
            "quokkasim_b200 has no CPU fallback." % path
        )
    L = C.CDLL(path)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    P = C.POINTER
    sig = {
        "qk_abi_version": (C.c_int, []),
        "qk_build_id": (C.c_char_p, []),
        "qk_model_create": (C.c_int, [P(vp)]),
        "qk_model_destroy": (None, [vp]),
        "qk_model_add_component": (C.c_int, [vp, P(ComponentDesc), P(u32)]),
        "qk_model_set_dist": (C.c_int, [vp, u32, u32, P(DistDesc), P(u32)]),
        "qk_model_add_delay_mode": (C.c_int, [vp, u32, P(DistDesc), P(DistDesc), P(u32), P(u32)]),
        "qk_model_connect": (C.c_int, [vp, u32, u32, i32]),
        "qk_model_set_logged": (C.c_int, [vp, u32, C.c_int]),
        "qk_model_schedule_env_state": (C.c_int, [vp, u64, u32, u32]),
        "qk_model_set_vector": (C.c_int, [vp, u32, u32, C.c_double, C.c_double, P(C.c_double)]),
        "qk_model_set_ports": (C.c_int, [vp, u32, u32, P(C.c_double)]),
        "qk_model_schedule_low_capacity": (C.c_int, [vp, u64, u32, C.c_double]),
        "qk_model_schedule_process_quantity": (C.c_int, [vp, u64, u32, P(DistDesc), P(u32)]),
        "qk_batch_read_vector": (C.c_int, [vp, u64, u32, P(C.c_double)]),
        "qk_model_set_vector_n": (C.c_int, [vp, u32, u32, u32, C.c_double, C.c_double, P(C.c_double)]),
        "qk_batch_read_vector_n": (C.c_int, [vp, u64, u32, P(C.c_double), u32, P(u32)]),
        "qk_model_n_components": (C.c_int, [vp, P(u32)]),
        "qk_model_n_dists": (C.c_int, [vp, P(u32)]),
        "qk_model_state_bytes": (C.c_int, [vp, C.c_int, P(u32)]),
        "qk_batch_create": (C.c_int, [vp, P(BatchConfig), P(vp)]),
        "qk_batch_destroy": (None, [vp]),
        "qk_batch_info_get": (C.c_int, [vp, P(BatchInfo)]),
        "qk_batch_set_tape": (C.c_int, [vp, u32, vp, u64]),
        "qk_batch_init": (C.c_int, [vp, u64]),
        "qk_batch_step_until": (C.c_int, [vp, u64]),
        "qk_batch_step_events": (C.c_int, [vp, u64]),
        "qk_batch_step_events_chunked": (C.c_int, [vp, u64, u64]),
        "qk_batch_step_events_sliced": (C.c_int, [vp, u64, u32]),
        "qk_batch_synchronize": (C.c_int, [vp]),
        "qk_batch_summary_get": (C.c_int, [vp, P(BatchSummary)]),
        "qk_batch_read_status": (C.c_int, [vp, u64, u64, vp, vp, vp]),
        "qk_batch_comp_stats": (C.c_int, [vp, u64, u64, u32, vp]),
        "qk_batch_reduce_stats": (C.c_int, [vp, vp, u32]),
        "qk_batch_reduce_partials": (C.c_int, [vp, vp, u32]),
        "qk_batch_reduce_partials_device": (C.c_int, [vp, vp, u32]),
        "qk_batch_combine_partials_device": (C.c_int, [vp, vp, u32, vp, u32]),
        "qk_batch_histograms_device": (C.c_int, [vp, vp, vp, u32]),
        "qk_batch_read_logs": (C.c_int, [vp, u64, u64, vp, vp, vp]),
        "qk_batch_read_log": (C.c_int, [vp, u64, vp, u64, P(u64), P(u64)]),
        "qk_batch_read_stock": (C.c_int, [vp, u64, u32, vp, u64, P(u64)]),
        "qk_batch_last_step_ms": (C.c_int, [vp, P(C.c_float), P(u32)]),
        "qk_last_error": (C.c_char_p, []),
        "qk_model_set_container_factory": (C.c_int, [vp, u32, P(DistDesc), P(C.c_double), C.c_double, P(u32)]),
        "qk_model_set_initial_containers": (C.c_int, [vp, u32, u32, vp, vp, vp]),
        "qk_batch_read_containers": (C.c_int, [vp, u64, u32, vp, vp, u64, P(u64)]),
        "qk_multi_create": (C.c_int, [vp, P(MultiConfig), P(vp)]),
        "qk_multi_destroy": (None, [vp]),
        "qk_multi_n_shards": (C.c_int, [vp, P(u32)]),
        "qk_multi_shard": (C.c_int, [vp, u32, P(vp), P(u64), P(u64), P(i32)]),
        "qk_multi_set_tape": (C.c_int, [vp, u32, vp, u64]),
        "qk_multi_init": (C.c_int, [vp, u64]),
        "qk_multi_step_until": (C.c_int, [vp, u64]),
        "qk_multi_step_events": (C.c_int, [vp, u64]),
        "qk_multi_synchronize": (C.c_int, [vp]),
        "qk_multi_last_step_ms": (C.c_int, [vp, P(C.c_float)]),
        "qk_multi_stats": (C.c_int, [vp, vp, u32, P(BatchSummary)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIBS[path] = L
    return L


def check(L, rc):
    if rc != 0:
        raise QkError(rc, (L.qk_last_error() or b"").decode("utf-8", "replace"))
