"""One job on several GPUs of one box from a single process: the Python face of qk_multi_* (include/quokka_b200.h).

The library shards the replications over the devices (contiguous ranges = Philox counter ranges), drives one batch per
GPU from the calling thread, and does the one exchange of the path -- the final reduction of summary statistics -- with
NCCL inside the library (all-gather of the packed partials, sum all-reduce of the histograms).  `parallel.py` is the
other way to use several GPUs: one process per GPU under torchrun, with torch.distributed carrying the same exchange.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from .api import BatchedSimulation


class MultiGpuSimulation:
    def __init__(self, model, n_reps, devices=None, t0=0, base_seed=0, rep_offset=0, log_capacity=0, flags=0, lib=None,
                 threads_per_block=0):
        """`model`: a workloads.Model (or anything with .components / .ext / .ext_low like it)."""
        # the model rows are built by the same code as for a single batch: a BatchedSimulation that never creates its
        # own device batch
        self._host = model.sim(lib, 1, t0=t0, base_seed=base_seed, log_capacity=log_capacity, flags=flags)
        self._host._build_model()
        self.L = self._host.L
        self.n_comp = self._host.n_comp
        self.n_reps = n_reps
        self.t0 = t0
        self.log_capacity = log_capacity
        cfg = capi.MultiConfig()
        cfg.n_reps = n_reps
        cfg.rep_offset = rep_offset
        cfg.base_seed = base_seed
        self._devs = None
        if devices:
            self._devs = (C.c_int32 * len(devices))(*devices)
            cfg.devices = C.cast(self._devs, C.POINTER(C.c_int32))
            cfg.n_devices = len(devices)
        cfg.log_capacity = log_capacity
        cfg.threads_per_block = threads_per_block
        cfg.flags = flags & 0xFF
        cfg.lanes_per_replication = (flags >> 8) & 0xFF
        self._multi = C.c_void_p()
        capi.check(self.L, self.L.qk_multi_create(self._host._model, C.byref(cfg), C.byref(self._multi)))
        capi.check(self.L, self.L.qk_multi_init(self._multi, int(t0)))

    def close(self):
        if self._multi:
            self.L.qk_multi_destroy(self._multi)
            self._multi = C.c_void_p()
        self._host.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_shards(self):
        n = C.c_uint32()
        capi.check(self.L, self.L.qk_multi_n_shards(self._multi, C.byref(n)))
        return n.value

    def shard(self, i):
        """(BatchedSimulation view of shard i for the per-replication read-back calls, first replication, count, device)"""
        b, r0, n, dev = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_int32()
        capi.check(self.L, self.L.qk_multi_shard(self._multi, i, C.byref(b), C.byref(r0), C.byref(n), C.byref(dev)))
        view = BatchedSimulation.__new__(BatchedSimulation)
        view.__dict__.update(self._host.__dict__)
        view._batch, view._model, view._owns = b, C.c_void_p(), False
        view.n_reps, view.rep_offset, view.device = n.value, r0.value, dev.value
        return view, r0.value, n.value, dev.value

    def init(self, t0=None):
        capi.check(self.L, self.L.qk_multi_init(self._multi, int(self.t0 if t0 is None else t0)))

    def step_events(self, n_events):
        capi.check(self.L, self.L.qk_multi_step_events(self._multi, int(n_events)))

    def step_until(self, t_ns):
        capi.check(self.L, self.L.qk_multi_step_until(self._multi, int(t_ns)))

    def synchronize(self):
        capi.check(self.L, self.L.qk_multi_synchronize(self._multi))

    def last_step_ms(self):
        ms = C.c_float()
        capi.check(self.L, self.L.qk_multi_last_step_ms(self._multi, C.byref(ms)))
        return ms.value

    def stats(self):
        """(per-component summaries over ALL replications, whole-job totals) -- NCCL reduction inside the library."""
        out = np.zeros(self.n_comp, dtype=capi.COMP_SUMMARY_DTYPE)
        tot = capi.BatchSummary()
        capi.check(self.L, self.L.qk_multi_stats(self._multi, out.ctypes.data, self.n_comp, C.byref(tot)))
        return out, {f: getattr(tot, f) for f, _ in tot._fields_}
