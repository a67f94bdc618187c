"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE -- see oracle/qk_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (quokkasim_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# numeric codes (restated from qk_oracle.h)
DISCRETE_STOCK, DISCRETE_PROCESS, DISCRETE_SOURCE, DISCRETE_SINK, DISCRETE_PARALLEL_PROCESS, ENVIRONMENT = 1, 2, 3, 4, 5, 6
VECTOR_STOCK, VECTOR_PROCESS, VECTOR_SOURCE, VECTOR_SINK, VECTOR_COMBINER, VECTOR_SPLITTER = 16, 17, 18, 19, 20, 21
CONTAINER_STOCK, CONTAINER_PROCESS, CONTAINER_SOURCE, CONTAINER_SINK = 32, 33, 34, 35
CONTAINER_PARALLEL_PROCESS, CONTAINER_LOAD_PROCESS, CONTAINER_UNLOAD_PROCESS = 36, 37, 38
CONTAINER_NONE_BITS = 0x7FF40000000000C0
DIST_CONSTANT, DIST_UNIFORM, DIST_TRIANGULAR, DIST_NORMAL, DIST_TRUNCNORMAL, DIST_EXPONENTIAL = 0, 1, 2, 3, 4, 5

LOG_DTYPE = np.dtype(
    [
        ("time_ns", "<u8"),
        ("comp", "<u2"),
        ("kind", "u1"),
        ("reason", "u1"),
        ("seq", "<u4"),
        ("src_comp", "<u2"),
        ("aux", "<u2"),
        ("src_seq", "<u4"),
        ("payload", "<u8"),
    ]
)
assert LOG_DTYPE.itemsize == 32
STATS_DTYPE = np.dtype([("n_a", "<u8"), ("n_b", "<u8"), ("t_a_ns", "<u8"), ("t_b_ns", "<u8"), ("fvalue", "<f8"),
                        ("faux", "<f8")])


class Dist(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("stream", C.c_uint32), ("p", C.c_double * 4)]


def make_dist(kind, stream=0, p=()):
    d = Dist()
    d.kind = kind
    d.stream = stream & 0xFFFFFFFF
    for i, v in enumerate(p):
        d.p[i] = float(v)
    return d


def build(force=False):
    """Compile oracle/libqk_oracle.so with the committed Makefile (building the checker is not using it)."""
    so = os.path.join(_HERE, "libqk_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h", ".inc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = build()
    L = C.CDLL(so)
    vp, u32, u64, i32, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_double
    P = C.POINTER
    sig = {
        "orc_model_new": (vp, []),
        "orc_model_free": (None, [vp]),
        "orc_model_add_component": (C.c_int, [vp, u32, u32, u32, u32]),
        "orc_model_set_dist": (C.c_int, [vp, u32, u32, P(Dist)]),
        "orc_model_add_delay_mode": (C.c_int, [vp, u32, P(Dist), P(Dist), P(C.c_int), P(C.c_int)]),
        "orc_model_connect": (C.c_int, [vp, u32, u32, i32]),
        "orc_model_schedule_env": (C.c_int, [vp, u64, u32, u32]),
        "orc_model_n_dists": (C.c_int, [vp]),
        "orc_model_set_vector": (C.c_int, [vp, u32, u32, dbl, dbl, P(dbl)]),
        "orc_model_set_vector_n": (C.c_int, [vp, u32, u32, u32, dbl, dbl, P(dbl)]),
        "orc_sim_vector_resource_n": (C.c_int, [vp, u32, P(dbl), u32]),
        "orc_model_set_splits": (C.c_int, [vp, u32, u32, P(dbl)]),
        "orc_model_schedule_low_capacity": (C.c_int, [vp, u64, u32, dbl]),
        "orc_model_schedule_process_quantity": (C.c_int, [vp, u64, u32, vp]),
        "orc_sim_new": (vp, [vp, u64, u64, u64, C.c_int]),
        "orc_sim_free": (None, [vp]),
        "orc_sim_set_tape": (C.c_int, [vp, u32, P(dbl), u64]),
        "orc_sim_init": (C.c_int, [vp]),
        "orc_sim_step_until": (C.c_int, [vp, u64]),
        "orc_sim_step_events": (C.c_int, [vp, u64]),
        "orc_sim_status": (u32, [vp]),
        "orc_sim_now": (u64, [vp]),
        "orc_sim_n_events": (u64, [vp]),
        "orc_sim_log_count": (u64, [vp]),
        "orc_sim_read_log": (u64, [vp, vp, u64]),
        "orc_sim_comp_stats": (C.c_int, [vp, u32, vp]),
        "orc_sim_stock_items": (u64, [vp, u32, P(u32), u64]),
        "orc_sim_vector_resource": (C.c_int, [vp, u32, P(dbl)]),
        "orc_sim_draws": (u64, [vp, u32]),
        "orc_model_set_container_factory": (C.c_int, [vp, u32, P(Dist), P(dbl), dbl, P(C.c_int)]),
        "orc_model_set_initial_containers": (C.c_int, [vp, u32, u32, vp, vp, vp]),
        "orc_sim_containers": (u64, [vp, u32, vp, vp, u64]),
        "orc_run_batch": (u64, [vp, u64, u64, u64, u64, u64, u64, C.c_int, C.c_int, vp, vp, vp, vp]),
        "orc_from_secs_f64": (C.c_int, [dbl, P(u64)]),
        "orc_philox4x32_10": (None, [P(u32), P(u32), P(u32)]),
        "orc_ln": (dbl, [dbl]),
        "orc_sample_native": (dbl, [P(Dist), u64, u64, P(u32)]),
        "orc_dm_new": (vp, []),
        "orc_dm_free": (None, [vp]),
        "orc_dm_add": (C.c_int, [vp, dbl, dbl]),
        "orc_dm_update_state": (None, [vp, u64, P(C.c_int), P(C.c_int)]),
        "orc_dm_next_event": (C.c_int, [vp, P(u64)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


class OracleModel:
    """Thin object wrapper over orc_model_*."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.orc_model_new()
        self.n_comp = 0

    def __del__(self):
        try:
            self.L.orc_model_free(self.h)
        except Exception:
            pass

    def add_component(self, kind, low=0, maxc=1, n_initial=0):
        i = self.L.orc_model_add_component(self.h, kind, low, maxc, n_initial)
        if i < 0:
            raise RuntimeError("orc_model_add_component failed")
        self.n_comp = i + 1
        return i

    def set_dist(self, comp, which, dist):
        i = self.L.orc_model_set_dist(self.h, comp, which, C.byref(dist))
        if i < 0:
            raise RuntimeError("orc_model_set_dist failed")
        return i

    def add_delay_mode(self, comp, until_delay, until_fix):
        a, b = C.c_int(), C.c_int()
        k = self.L.orc_model_add_delay_mode(self.h, comp, C.byref(until_delay), C.byref(until_fix), C.byref(a), C.byref(b))
        if k < 0:
            raise RuntimeError("orc_model_add_delay_mode failed")
        return k, a.value, b.value

    def connect(self, a, b, n=-1):
        rc = self.L.orc_model_connect(self.h, a, b, n)
        if rc != 0:
            raise RuntimeError("No component connection defined from %d to %d (rc=%d)" % (a, b, rc))

    def schedule_env(self, t_ns, env, state):
        if self.L.orc_model_schedule_env(self.h, t_ns, env, state) != 0:
            raise RuntimeError("orc_model_schedule_env failed")

    def n_dists(self):
        return self.L.orc_model_n_dists(self.h)

    def set_vector(self, comp, dim, low, maxc, init, total_mask=None):
        arr = (C.c_double * 8)(*([float(x) for x in init] + [0.0] * (8 - len(init))))
        if total_mask is None:
            rc = self.L.orc_model_set_vector(self.h, comp, dim, low, maxc, arr)
        else:
            rc = self.L.orc_model_set_vector_n(self.h, comp, dim, total_mask, low, maxc, arr)
        if rc != 0:
            raise RuntimeError("orc_model_set_vector failed")

    def set_splits(self, comp, ratios):
        arr = (C.c_double * len(ratios))(*ratios)
        if self.L.orc_model_set_splits(self.h, comp, len(ratios), arr) != 0:
            raise RuntimeError("orc_model_set_splits failed")

    def set_container_factory(self, source, create_quantity, split, capacity):
        did = C.c_int(-1)
        arr = (C.c_double * 3)(*[float(x) for x in split])
        rc = self.L.orc_model_set_container_factory(self.h, source, C.byref(create_quantity) if create_quantity else None,
                                                    arr, capacity, C.byref(did))
        if rc != 0:
            raise RuntimeError("orc_model_set_container_factory failed")
        return did.value

    def set_initial_containers(self, stock, resources, has, capacities):
        r = np.ascontiguousarray(resources, dtype=np.float64).reshape(-1, 3)
        h = np.ascontiguousarray(has, dtype=np.uint8)
        c = np.ascontiguousarray(capacities, dtype=np.float64)
        if self.L.orc_model_set_initial_containers(self.h, stock, len(h), r.ctypes.data, h.ctypes.data, c.ctypes.data) != 0:
            raise RuntimeError("orc_model_set_initial_containers failed")

    def schedule_low_capacity(self, t_ns, stock, value):
        if self.L.orc_model_schedule_low_capacity(self.h, t_ns, stock, value) != 0:
            raise RuntimeError("orc_model_schedule_low_capacity failed")

    def schedule_process_quantity(self, t_ns, process, dist):
        did = self.L.orc_model_schedule_process_quantity(self.h, t_ns, process, C.byref(dist))
        if did < 0:
            raise RuntimeError("orc_model_schedule_process_quantity failed")
        return did

    def run_batch(self, base_seed, rep0, n_reps, t0_ns=0, n_events=0, t_until_ns=0, log=False, threads=1,
                  want_stats=True):
        stats = np.zeros((n_reps, self.n_comp), dtype=STATS_DTYPE) if want_stats else None
        status = np.zeros(n_reps, dtype=np.uint32)
        nev = np.zeros(n_reps, dtype=np.uint64)
        now = np.zeros(n_reps, dtype=np.uint64)
        total = self.L.orc_run_batch(
            self.h, base_seed, rep0, n_reps, t0_ns, n_events, t_until_ns, int(log), threads,
            stats.ctypes.data if stats is not None else None, status.ctypes.data, nev.ctypes.data, now.ctypes.data,
        )
        return total, stats, status, nev, now


class OracleSim:
    def __init__(self, model, base_seed=0, rep=0, t0_ns=0, log=True):
        self.L = lib()
        self.model = model
        self.h = self.L.orc_sim_new(model.h, base_seed, rep, t0_ns, int(log))

    def __del__(self):
        try:
            self.L.orc_sim_free(self.h)
        except Exception:
            pass

    def set_tape(self, dist_id, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.L.orc_sim_set_tape(self.h, dist_id, v.ctypes.data_as(C.POINTER(C.c_double)), v.size)

    def init(self):
        return self.L.orc_sim_init(self.h)

    def step_until(self, t_ns):
        return self.L.orc_sim_step_until(self.h, t_ns)

    def step_events(self, n):
        return self.L.orc_sim_step_events(self.h, n)

    @property
    def status(self):
        return self.L.orc_sim_status(self.h)

    @property
    def now(self):
        return self.L.orc_sim_now(self.h)

    @property
    def n_events(self):
        return self.L.orc_sim_n_events(self.h)

    def log(self):
        n = self.L.orc_sim_log_count(self.h)
        buf = np.zeros(n, dtype=LOG_DTYPE)
        if n:
            self.L.orc_sim_read_log(self.h, buf.ctypes.data, n)
        return buf

    def comp_stats(self, comp):
        out = np.zeros(1, dtype=STATS_DTYPE)
        self.L.orc_sim_comp_stats(self.h, comp, out.ctypes.data)
        return out[0]

    def stock_items(self, comp):
        buf = (C.c_uint32 * 4096)()
        n = self.L.orc_sim_stock_items(self.h, comp, buf, 4096)
        return list(buf[: min(n, 4096)])

    def vector_resource(self, comp):
        out = (C.c_double * 3)()
        self.L.orc_sim_vector_resource(self.h, comp, out)
        return list(out)

    def vector_resource_n(self, comp):
        out = (C.c_double * 8)()
        n = self.L.orc_sim_vector_resource_n(self.h, comp, out, 8)
        return list(out[:n])

    def containers(self, comp):
        h = np.zeros(4096, dtype=np.uint32)
        r = np.zeros((4096, 3), dtype=np.uint64)
        n = self.L.orc_sim_containers(self.h, comp, h.ctypes.data, r.ctypes.data, 4096)
        return h[:n], r[:n]

    def draws(self, dist_id):
        return self.L.orc_sim_draws(self.h, dist_id)
