"""Host-side mirror of the QuokkaSim model-builder API for the batched B200 executor.

The reference is Rust and there is no Rust toolchain in this image, so the host side above the C ABI is
restated here with the reference's names, argument meaning and error behaviour:

  reference (quokkasim v0.2.2)                                    here
  --------------------------------------------------------------  ---------------------------------------------
  DistributionConfig / DistributionFactory::create  common.rs:37-148   DistributionConfig / DistributionFactory.create
  Distribution::Constant(x)                          common.rs:28       Distribution.Constant(x)
  #[derive(WithMethods)] builders  quokkasim_derive_macros/src/lib.rs:22-94
                                                                   .new() .with_name() .with_code() .with_type()
                                                                   .with_low_capacity() .with_max_capacity()
                                                                   .with_initial_resource() .with_process_time_distr()
                                                                   .with_delay_mode() .with_item_factory() (+ _inplace)
  ComponentModel::StringStock(c, Mailbox::new()) ... core.rs:505-557    ComponentModel.StringStock(c, Mailbox.new()) ...
  connect_components!(&mut a, &mut b[, n])           core.rs:1405-1412  connect_components(a, b[, n])
  connect_logger!(&mut logger, &mut c)               core.rs:1415-1422  connect_logger(logger, c)
  register_component!(sim_init, c)                   core.rs:1424-1428  sim_init = register_component(sim_init, c)
  SimInit::new() / sim_init.init(t0)                 discrete_queue.rs:100-109
                                                                   BatchedSimInit.new() / .init(t0, n_replications, ...)
  create_scheduled_event!(&mut sched,&t,&cfg,&addr,&mut df) core.rs:1430-1434
                                                                   create_scheduled_event(sched, t, cfg, addr, df)
  simu.step_until(t)                                 discrete_queue.rs:111  sim.step_until(t) / sim.step_events(n)
  logger.write_csv(dir)                              core.rs:388-401    logger.write_csv(dir, sim, replication=r)

Everything here is bookkeeping: it appends rows to the C-ABI model tables.  All simulation work happens in
libquokka_b200.so on the GPU; nothing in this module can advance a replication on the CPU.
"""
import ctypes as C
import itertools
import os

import numpy as np

from . import _capi as capi

NANOS_PER_SEC = 1_000_000_000
_stamp = itertools.count()


# --------------------------------------------------------------------------------------- time helpers
class MonotonicTime:
    """tai-time MonotonicTime reduced to what the examples use: absolute integer ns since the TAI epoch."""

    EPOCH = 0

    @staticmethod
    def try_from_date_time(year, month, day, hour, minute, second, nanos):
        import datetime

        d = datetime.datetime(year, month, day, hour, minute, second, tzinfo=datetime.timezone.utc)
        e = datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)
        return int((d - e).total_seconds()) * NANOS_PER_SEC + int(nanos)


class Duration:
    @staticmethod
    def from_secs(s):
        return int(s) * NANOS_PER_SEC

    @staticmethod
    def from_nanos(n):
        return int(n)

    @staticmethod
    def from_millis(ms):
        return int(ms) * 1_000_000


# --------------------------------------------------------------------------------------- distributions
class DistributionParametersError(ValueError):
    """common.rs:51-62"""


class Distribution:
    """An instantiated distribution (common.rs:25-32). `stream` is the seed the factory assigned."""

    def __init__(self, kind, params, stream=0):
        self.kind = kind
        self.params = [float(p) for p in params]
        self.stream = int(stream)

    @staticmethod
    def Constant(value):
        return Distribution(capi.DIST_CONSTANT, [value])

    @staticmethod
    def default():
        return Distribution.Constant(1.0)  # common.rs:181-185

    def desc(self):
        d = capi.DistDesc()
        d.kind = self.kind
        d.stream = self.stream & 0xFFFFFFFF
        for i, p in enumerate(self.params):
            d.p[i] = p
        return d

    def __repr__(self):
        return "Distribution(kind=%d, params=%r, stream=%d)" % (self.kind, self.params, self.stream)


class DistributionConfig:
    """common.rs:37-44"""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.kw = kw

    @staticmethod
    def Uniform(min, max):
        return DistributionConfig("Uniform", min=min, max=max)

    @staticmethod
    def Triangular(min, max, mode):
        return DistributionConfig("Triangular", min=min, max=max, mode=mode)

    @staticmethod
    def Constant(value):
        return DistributionConfig("Constant", value=value)

    @staticmethod
    def Normal(mean, std):
        return DistributionConfig("Normal", mean=mean, std=std)

    @staticmethod
    def TruncNormal(mean, std, min=None, max=None):
        return DistributionConfig("TruncNormal", mean=mean, std=std, min=min, max=max)

    @staticmethod
    def Exponential(mean):
        return DistributionConfig("Exponential", mean=mean)


class DistributionFactory:
    """common.rs:46-148.  Seeds increment per created distribution, EXCEPT that the Normal / TruncNormal /
    Exponential arms return early and skip `next_seed += 1` (common.rs:98,121,134 vs :145) --
    reproduced, because the stream id decides which instances share a variate sequence."""

    def __init__(self, base_seed, next_seed=None):
        self.base_seed = int(base_seed)
        self.next_seed = int(base_seed if next_seed is None else next_seed)

    @staticmethod
    def new(base_seed):
        return DistributionFactory(base_seed)

    def create(self, config):
        k, kw = config.kind, config.kw
        if k == "Uniform":
            result = Distribution(capi.DIST_UNIFORM, [kw["min"], kw["max"]], self.next_seed)
        elif k == "Triangular":
            mn, mx, mode = kw["min"], kw["max"], kw["mode"]
            if not (mx >= mode >= mn) or not (mx > mn):  # rand_distr TriangularError
                raise DistributionParametersError("invalid Triangular parameters")
            result = Distribution(capi.DIST_TRIANGULAR, [mn, mx, mode], self.next_seed)
        elif k == "Constant":
            result = Distribution.Constant(kw["value"])
        elif k == "Normal":
            if not (kw["std"] >= 0 and np.isfinite(kw["std"])):
                raise DistributionParametersError("invalid Normal parameters")
            return Distribution(capi.DIST_NORMAL, [kw["mean"], kw["std"]], self.next_seed)
        elif k == "TruncNormal":
            mn = -np.finfo(np.float64).max if kw["min"] is None else kw["min"]  # f64::MIN
            mx = np.finfo(np.float64).max if kw["max"] is None else kw["max"]
            if mn >= mx:
                raise DistributionParametersError("Minimum value cannot be greater than or equal maximum value")
            return Distribution(capi.DIST_TRUNCNORMAL, [kw["mean"], kw["std"], mn, mx], self.next_seed)
        elif k == "Exponential":
            if not (kw["mean"] > 0):
                raise DistributionParametersError("invalid Exponential parameters")
            return Distribution(capi.DIST_EXPONENTIAL, [kw["mean"]], self.next_seed)
        else:
            raise DistributionParametersError("unknown distribution %r" % k)
        self.next_seed += 1
        return result


class DelayMode:
    """delays.rs:6-11"""

    def __init__(self, name, until_delay_distr, until_fix_distr):
        self.name = name
        self.until_delay_distr = until_delay_distr
        self.until_fix_distr = until_fix_distr


class DelayModeChange:
    """delays.rs:45-50 (only Add can be expressed before init; Remove/RemoveAll edit the builder list)"""

    def __init__(self, op, arg=None):
        self.op = op
        self.arg = arg

    @staticmethod
    def Add(mode):
        return DelayModeChange("Add", mode)

    @staticmethod
    def Remove(name):
        return DelayModeChange("Remove", name)

    @staticmethod
    def RemoveAll():
        return DelayModeChange("RemoveAll")


class StringItemFactory:
    """discrete.rs:664-686: items are "{prefix}_{index:0>num_digits}"."""

    def __init__(self, prefix="Item", next_index=0, num_digits=4):
        self.prefix = prefix
        self.next_index = next_index
        self.num_digits = num_digits

    def item_name(self, index):
        return "%s_%s" % (self.prefix, str(self.next_index + index).rjust(self.num_digits, "0"))


class ItemDeque(list):
    """discrete.rs:82-103"""

    @staticmethod
    def default():
        return ItemDeque()


class BasicEnvironmentState:
    Normal = 0
    Stopped = 1


# --------------------------------------------------------------------------------------- components
class _Component:
    KIND = None
    DEFAULT_NAME = ""

    def __init__(self):
        self.element_name = self.DEFAULT_NAME
        self.element_code = ""
        self.element_type = self.DEFAULT_NAME
        self._conns = []  # (stamp, a, b, n) recorded by connect_components
        self._id = None
        self._logged_by = None

    @classmethod
    def new(cls):
        return cls()

    def with_name(self, name):
        self.element_name = str(name)
        return self

    def with_code(self, code):
        self.element_code = str(code)
        return self

    def with_type(self, t):
        self.element_type = str(t)
        return self


class DiscreteStock(_Component):
    """discrete.rs:38-79"""

    KIND = capi.DISCRETE_STOCK
    DEFAULT_NAME = "DiscreteStock"

    def __init__(self):
        super().__init__()
        self.low_capacity = 0
        self.max_capacity = 1
        self.resource = ItemDeque()
        self.capacity_hint = 0

    def with_low_capacity(self, v):
        self.low_capacity = int(v)
        return self

    def with_max_capacity(self, v):
        self.max_capacity = int(v)
        return self

    def with_low_capacity_inplace(self, v):
        self.low_capacity = int(v)

    def with_max_capacity_inplace(self, v):
        self.max_capacity = int(v)

    def with_initial_resource(self, resource):
        self.resource = ItemDeque(resource)
        return self

    def with_capacity_hint(self, slots):
        """B200 executor only: item-ring slots reserved per replication (default: max_capacity + producers)."""
        self.capacity_hint = int(slots)
        return self


class _ProcessLike(_Component):
    def __init__(self):
        super().__init__()
        self.process_time_distr = Distribution.default()
        self.process_quantity_distr = Distribution.default()
        self.delay_modes = []
        self.capacity_hint = 0

    def with_process_time_distr(self, d):
        self.process_time_distr = d
        return self

    def with_process_time_distr_inplace(self, d):
        self.process_time_distr = d

    def with_process_quantity_distr(self, d):
        self.process_quantity_distr = d
        return self

    def with_process_quantity_distr_inplace(self, d):
        self.process_quantity_distr = d

    def with_delay_mode(self, change):
        self.with_delay_mode_inplace(change)
        return self

    def with_delay_mode_inplace(self, change):
        # DelayModes::modify delays.rs:157-174
        if change.op == "Add":
            for i, m in enumerate(self.delay_modes):
                if m.name == change.arg.name:  # IndexMap::insert on an existing key keeps its position
                    self.delay_modes[i] = change.arg
                    return
            self.delay_modes.append(change.arg)
        elif change.op == "Remove":
            self.delay_modes = [m for m in self.delay_modes if m.name != change.arg]
        else:
            self.delay_modes = []


class DiscreteProcess(_ProcessLike):
    KIND = capi.DISCRETE_PROCESS
    DEFAULT_NAME = "DiscreteProcess"


class DiscreteSource(_ProcessLike):
    KIND = capi.DISCRETE_SOURCE
    DEFAULT_NAME = "DiscreteSource"

    def __init__(self):
        super().__init__()
        self.item_factory = StringItemFactory()
        self._factory_set = False  # F::default() until with_item_factory: the variant decides which F

    def with_item_factory(self, f):
        self.item_factory = f
        self._factory_set = True
        return self

    def with_item_factory_inplace(self, f):
        self.item_factory = f
        self._factory_set = True


class DiscreteSink(_ProcessLike):
    KIND = capi.DISCRETE_SINK
    DEFAULT_NAME = "DiscreteSink"


class DiscreteParallelProcess(_ProcessLike):
    KIND = capi.DISCRETE_PARALLEL_PROCESS
    DEFAULT_NAME = "DiscreteParallelProcess"

    def with_capacity_hint(self, slots):
        """B200 executor only: in-flight item slots per replication (the reference Vec is unbounded)."""
        self.capacity_hint = int(slots)
        return self


class BasicEnvironment(_Component):
    """core.rs:218-290"""

    KIND = capi.ENVIRONMENT
    DEFAULT_NAME = ""

    def __init__(self):
        super().__init__()
        self.element_type = "BasicEnvironment"
        self.state = BasicEnvironmentState.Normal

    def with_state(self, state):
        self.state = state
        return self


class Mailbox:
    """Kept so that `ComponentModel::X(component, Mailbox::new())` reads like the reference; inert here."""

    @staticmethod
    def new():
        return Mailbox()


class ComponentModelAddress:
    def __init__(self, variant, component):
        self.variant = variant
        self.component = component


class ComponentModel:
    """The generated enum of core.rs:504-562, for the discrete family and BasicEnvironment."""

    _VARIANTS = {
        "StringStock": DiscreteStock,
        "StringProcess": DiscreteProcess,
        "StringParallelProcess": DiscreteParallelProcess,
        "StringSource": DiscreteSource,
        "StringSink": DiscreteSink,
        "BasicEnvironment": BasicEnvironment,
    }

    def __init__(self, variant, component, mailbox=None):
        if variant not in self._VARIANTS:
            raise ValueError("unknown ComponentModel variant %s" % variant)
        if not isinstance(component, self._VARIANTS[variant]):
            raise TypeError("ComponentModel.%s expects a %s" % (variant, self._VARIANTS[variant].__name__))
        self.variant = variant
        self.component = component
        self.mailbox = mailbox
        if variant in _VECTOR_VARIANTS:
            _bind_vector_variant(variant, component)
        if variant in _CONTAINER_VARIANTS:
            _bind_container_variant(variant, component)

    def __str__(self):  # #[derive(Display)] prints the variant name
        return self.variant

    def get_address(self):
        return ComponentModelAddress(self.variant, self.component)

    @staticmethod
    def connect_components(a, b, n=None):
        return connect_components(a, b, n)

    @staticmethod
    def register_component(sim_init, component):
        return register_component(sim_init, component)


def _variant_ctor(name):
    def ctor(component, mailbox=None):
        return ComponentModel(name, component, mailbox)

    return staticmethod(ctor)


for _v in ComponentModel._VARIANTS:
    setattr(ComponentModel, _v, _variant_ctor(_v))


class ComponentConnectionError(ValueError):
    """The Err(..) of connect_components for an undefined pair (core.rs:1174-1176, discrete_queue.rs:14-20)."""


_STOCK_TO_PROC = {"StringProcess", "StringParallelProcess", "StringSink"}
_PROC_TO_STOCK = {"StringProcess", "StringParallelProcess", "StringSource"}
_ENV_TARGETS = {"StringProcess", "StringParallelProcess", "StringSource", "StringSink"}


def connect_components(a, b, n=None):
    """connect_components!(&mut a, &mut b[, n]) (core.rs:565-1178, discrete arms :886-921, env arms :1129-1148)."""
    va, vb = a.variant, b.variant
    ok = (
        (va == "StringStock" and vb in _STOCK_TO_PROC)
        or (vb == "StringStock" and va in _PROC_TO_STOCK)
        or (va == "BasicEnvironment" and (vb in _ENV_TARGETS or vb in _CONTAINER_ENV_TARGETS
                                          or (vb in _VECTOR_VARIANTS and not vb.endswith("Stock"))))
        or _vector_connection_ok(va, vb, n)
        or _container_connection_ok(va, vb, n)
    )
    if not ok:
        raise ComponentConnectionError(
            "No component connection defined from %s to %s (n=%s)" % (a, b, "None" if n is None else "Some(%d)" % n)
        )
    a.component._conns.append((next(_stamp), a.component, b.component, -1 if n is None else int(n)))


# --------------------------------------------------------------------------------------- loggers
class _Logger:
    ACCEPTS = ()

    def __init__(self, name):
        self.name = str(name)
        self.components = []

    @classmethod
    def new(cls, name):
        return cls(name)

    def get_name(self):
        return self.name


class DiscreteStockLogger(_Logger):
    ACCEPTS = ("StringStock",)


class DiscreteProcessLogger(_Logger):
    ACCEPTS = ("StringProcess", "StringParallelProcess", "StringSource", "StringSink")


class ContainerStockLogger(DiscreteStockLogger):
    """DiscreteStockLogger<Vector3Container> (core.rs:1292, 1330)"""

    ACCEPTS = ("Vector3ContainerStock",)


class ContainerProcessLogger(DiscreteProcessLogger):
    """DiscreteProcessLogger<Vector3Container> (core.rs:1293, 1331-1334): no Source / Sink arm"""

    ACCEPTS = ("Vector3ContainerProcess", "Vector3ContainerParallelProcess", "Vector3ContainerLoadProcess",
               "Vector3ContainerUnloadProcess")


class BasicEnvironmentLogger(_Logger):
    ACCEPTS = ("BasicEnvironment",)


class VectorStockLogger(_Logger):
    """vector.rs:211-231"""

    ACCEPTS = ("F64Stock", "Vector3Stock")


class VectorProcessLogger(_Logger):
    """vector.rs:543-562: one logger type for process, source, sink, combiner and splitter records"""

    ACCEPTS = tuple("%s%s" % (p, k) for p in ("F64", "Vector3")
                    for k in ["Process", "Source", "Sink"] + ["Combiner%d" % i for i in range(1, 6)]
                    + ["Splitter%d" % i for i in range(1, 6)])


class ComponentLogger:
    """The generated logger enum (core.rs:1216-1365)."""

    _VARIANTS = {
        "StringStockLogger": DiscreteStockLogger,
        "StringProcessLogger": DiscreteProcessLogger,
        "BasicEnvironmentLogger": BasicEnvironmentLogger,
        "F64StockLogger": VectorStockLogger,
        "F64ProcessLogger": VectorProcessLogger,
        "Vector3StockLogger": VectorStockLogger,
        "Vector3ProcessLogger": VectorProcessLogger,
        "Vector3ContainerStockLogger": ContainerStockLogger,
        "Vector3ContainerProcessLogger": ContainerProcessLogger,
    }

    def __init__(self, variant, logger):
        if not isinstance(logger, self._VARIANTS[variant]):
            raise TypeError("ComponentLogger.%s expects a %s" % (variant, self._VARIANTS[variant].__name__))
        self.variant = variant
        self.logger = logger

    def __str__(self):
        return self.variant

    def write_csv(self, directory, sim, replication=0, record_map=None, header=None):
        from .csvlog import write_csv

        return write_csv(self, directory, sim, replication, record_map, header)


for _v in ComponentLogger._VARIANTS:
    setattr(ComponentLogger, _v, staticmethod(lambda logger, _v=_v: ComponentLogger(_v, logger)))


def connect_logger(logger, component, n=None):
    """connect_logger!(&mut logger, &mut component) (core.rs:1303-1338)."""
    if component.variant not in logger.logger.ACCEPTS:
        raise ComponentConnectionError(
            "No logger connection defined from %s to %s (n=%s)"
            % (logger, component, "None" if n is None else "Some(%d)" % n)
        )
    logger.logger.components.append(component.component)
    component.component._logged_by = logger.logger


# --------------------------------------------------------------------------------------- scheduled events
class ScheduledEventConfig:
    """core.rs:1367-1375"""

    def __init__(self, kind, value):
        self.kind = kind
        self.value = value

    def __str__(self):
        return self.kind

    @staticmethod
    def SetLowCapacity(v):
        return ScheduledEventConfig("SetLowCapacity", v)

    @staticmethod
    def SetMaxCapacity(v):
        return ScheduledEventConfig("SetMaxCapacity", v)

    @staticmethod
    def SetProcessQuantity(cfg):
        return ScheduledEventConfig("SetProcessQuantity", cfg)

    @staticmethod
    def SetProcessTime(cfg):
        return ScheduledEventConfig("SetProcessTime", cfg)

    @staticmethod
    def SetEnvironmentState(state):
        return ScheduledEventConfig("SetEnvironmentState", state)


ScheduledEvent = ScheduledEventConfig  # discrete_queue.rs / assembly_line.rs name the enum ScheduledEvent


class Scheduler:
    def __init__(self, sim):
        self._sim = sim


def create_scheduled_event(scheduler, time, event_config, component_addr, df=None):
    """create_scheduled_event!(&mut sched, &time, &cfg, &addr, &mut df) (core.rs:1377-1401, 1430-1434)."""
    if event_config.kind == "SetEnvironmentState" and component_addr.variant == "BasicEnvironment":
        scheduler._sim._schedule_env(int(time), component_addr.component, int(event_config.value))
        return
    if event_config.kind == "SetLowCapacity" and component_addr.variant == "F64Stock":  # core.rs:1382-1386
        scheduler._sim._schedule_low(int(time), component_addr.component, float(event_config.value))
        return
    if event_config.kind == "SetProcessQuantity" and component_addr.variant == "F64Process":  # core.rs:1387-1391
        # `df.create(distr.clone())?` runs NOW, at schedule time: the factory hands out its next seed (and skips the
        # increment for Normal / TruncNormal / Exponential); the instance replaces
        # process_quantity_distr at `time`, followed by update_state(SCH_000000)
        if df is None:
            raise TypeError("create_scheduled_event(SetProcessQuantity) needs the DistributionFactory (`&mut df`)")
        cfg = event_config.value
        dist = cfg if isinstance(cfg, Distribution) else df.create(cfg)
        scheduler._sim._schedule_qty(int(time), component_addr.component, dist)
        return dist
    raise NotImplementedError(
        "schedule_event not implemented for types (%s, %s)" % (event_config, component_addr.variant)
    )


# --------------------------------------------------------------------------------------- sim init / run
def register_component(sim_init, component):
    """register_component!(sim_init, c): registration order = component id = scheduler origin rank - 1."""
    sim_init._components.append(component)
    return sim_init


class BatchedSimInit:
    """Stands where nexosim::SimInit stands in the examples."""

    def __init__(self):
        self._components = []

    @staticmethod
    def new():
        return BatchedSimInit()

    def add_model(self, component):
        return register_component(self, component)

    def init(self, t0, n_replications=1, base_seed=0, device=0, rep_offset=0, log_capacity=0, threads_per_block=0,
             stream=None, lib=None, flags=0):
        """Returns (BatchedSimulation, Scheduler), like SimInit::init returns (Simulation, Scheduler).

        log_capacity: records kept per replication (power of two); 0 disables logging even if loggers are
        connected (statistics are always kept)."""
        sim = BatchedSimulation(self._components, int(t0), int(n_replications), int(base_seed), int(device),
                                int(rep_offset), int(log_capacity), int(threads_per_block), stream, lib, int(flags))
        return sim, Scheduler(sim)


class BatchedSimulation:
    """Many independent replications of one model, advanced in lockstep on one GPU.

    The device batch is created lazily at the first step/readout so that create_scheduled_event! calls made
    after init() (as in assembly_line.rs:240-246) still land in the model tables."""

    def __init__(self, components, t0, n_reps, base_seed, device, rep_offset, log_capacity, threads_per_block, stream,
                 lib, flags=0):
        self.L = capi.load_library(lib)
        self.components = [c.component for c in components]
        self.variants = [c.variant for c in components]
        self.t0 = t0
        self.n_reps = n_reps
        self.base_seed = base_seed
        self.device = device
        self.rep_offset = rep_offset
        self.log_capacity = log_capacity
        self.threads_per_block = threads_per_block
        self.stream = stream
        self.flags = flags
        self._model = C.c_void_p()
        self._batch = C.c_void_p()
        self._ext = []
        self._tapes = []
        self._dist_ids = {}
        for i, c in enumerate(self.components):
            if c._id is not None and c._id != i:
                raise ValueError("component %r registered twice" % c.element_name)
            c._id = i
        self.n_comp = len(self.components)

    # -- lowering: component objects -> C-ABI rows ---------------------------------------------------
    def _build_model(self):
        L = self.L
        capi.check(L, L.qk_model_create(C.byref(self._model)))
        registered = set(id(c) for c in self.components)
        for c in self.components:
            d = capi.ComponentDesc()
            d.kind = getattr(c, "_container_kind", None) or c.KIND
            is_container = d.kind >= capi.CONTAINER_STOCK
            if isinstance(c, DiscreteStock):
                d.low_capacity = c.low_capacity
                d.max_capacity = c.max_capacity
                d.n_initial_items = 0 if is_container else len(c.resource)
                d.capacity_hint = c.capacity_hint
            elif hasattr(c, "max_process_count"):
                d.max_capacity = c.max_process_count
            elif isinstance(c, BasicEnvironment):
                d.initial_env_state = c.state
            elif hasattr(c, "capacity_hint"):
                d.capacity_hint = c.capacity_hint
            out = C.c_uint32()
            capi.check(L, L.qk_model_add_component(self._model, C.byref(d), C.byref(out)))
            assert out.value == c._id
            if is_container and isinstance(c, DiscreteStock) and len(c.resource):
                n = len(c.resource)
                res = np.zeros((n, 3), dtype=np.float64)
                has = np.zeros(n, dtype=np.uint8)
                caps = np.zeros(n, dtype=np.float64)
                for k, it in enumerate(c.resource):
                    caps[k] = it.capacity
                    if it.resource is not None:
                        has[k] = 1
                        res[k] = it.resource.values
                capi.check(L, L.qk_model_set_initial_containers(self._model, c._id, n, res.ctypes.data, has.ctypes.data,
                                                                caps.ctypes.data))
            if is_container and isinstance(c, DiscreteSource):
                f = c.item_factory
                split = (C.c_double * 3)(*f.create_component_split)
                did = C.c_uint32()
                desc = f.create_quantity_distr.desc() if f.create_quantity_distr is not None else None
                capi.check(L, L.qk_model_set_container_factory(self._model, c._id, C.byref(desc) if desc else None, split,
                                                               f.capacity, C.byref(did)))
                if desc is not None:
                    self._dist_ids.setdefault(id(f.create_quantity_distr), []).append(did.value)
            if hasattr(c, "vector_rows"):
                if c.dim is None:
                    raise TypeError("vector component %r was never wrapped in a ComponentModel variant" % c.element_name)
                dim, low, maxc, init = c.vector_rows()
                arr = (C.c_double * 8)(*(list(init) + [0.0] * (8 - len(init))))
                if getattr(c, "total_mask", None) is None:
                    capi.check(L, L.qk_model_set_vector(self._model, c._id, dim, low, maxc, arr))
                else:  # a user-defined resource type (vector_api.define_resource)
                    capi.check(L, L.qk_model_set_vector_n(self._model, c._id, dim, c.total_mask, low, maxc, arr))
                if getattr(c, "n_ports", None):
                    ratios = (C.c_double * 5)(*(list(c.split_ratios) + [0.0] * (5 - len(c.split_ratios))))
                    capi.check(L, L.qk_model_set_ports(self._model, c._id, c.n_ports, ratios))
            if isinstance(c, _ProcessLike):
                for which, dist in ((0, c.process_time_distr), (1, c.process_quantity_distr)):
                    did = C.c_uint32()
                    desc = dist.desc()
                    capi.check(L, L.qk_model_set_dist(self._model, c._id, which, C.byref(desc), C.byref(did)))
                    self._dist_ids.setdefault(id(dist), []).append(did.value)
                for m in c.delay_modes:
                    a, b = C.c_uint32(), C.c_uint32()
                    da, db = m.until_delay_distr.desc(), m.until_fix_distr.desc()
                    capi.check(L, L.qk_model_add_delay_mode(self._model, c._id, C.byref(da), C.byref(db), C.byref(a),
                                                            C.byref(b)))
                    self._dist_ids.setdefault(id(m.until_delay_distr), []).append(a.value)
                    self._dist_ids.setdefault(id(m.until_fix_distr), []).append(b.value)
            if c._logged_by is not None:
                capi.check(L, L.qk_model_set_logged(self._model, c._id, 1))
        conns = sorted((x for c in self.components for x in c._conns), key=lambda x: x[0])
        for _, a, b, n in conns:
            if id(a) not in registered or id(b) not in registered:
                raise ValueError("connect_components involves a component that was never registered")
            capi.check(L, L.qk_model_connect(self._model, a._id, b._id, n))
        for kind, t, comp, value in self._ext:
            if kind == "env":
                capi.check(L, L.qk_model_schedule_env_state(self._model, t, comp._id, value))
            elif kind == "qty":
                did = C.c_uint32()
                desc = value.desc()
                capi.check(L, L.qk_model_schedule_process_quantity(self._model, t, comp._id, C.byref(desc), C.byref(did)))
                self._dist_ids.setdefault(id(value), []).append(did.value)
            else:
                capi.check(L, L.qk_model_schedule_low_capacity(self._model, t, comp._id, value))

    def _schedule_env(self, t_ns, env, state):
        if self._batch:
            raise RuntimeError("scheduled events must be created before the first step of a BatchedSimulation")
        self._ext.append(("env", t_ns, env, state))

    def _schedule_qty(self, t_ns, process, dist):
        if self._batch:
            raise RuntimeError("scheduled events must be created before the first step of a BatchedSimulation")
        self._ext.append(("qty", t_ns, process, dist))

    def _schedule_low(self, t_ns, stock, value):
        if self._batch:
            raise RuntimeError("scheduled events must be created before the first step of a BatchedSimulation")
        self._ext.append(("low", t_ns, stock, value))

    def set_tape(self, distribution, values):
        """Pre-drawn variate tape for one Distribution instance: values[replication, i] is what its i-th
        .sample() call returns in that replication (parity mode)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.ndim != 2 or v.shape[0] != self.n_reps:
            raise ValueError("tape must have shape (n_replications, draws)")
        if self._batch:
            raise RuntimeError("set_tape must be called before the first step")
        self._tapes.append((distribution, v))

    def _materialize(self):
        if self._batch:
            return
        L = self.L
        self._build_model()
        cfg = capi.BatchConfig()
        cfg.n_reps = self.n_reps
        cfg.rep_offset = self.rep_offset
        cfg.base_seed = self.base_seed
        cfg.device = self.device
        cfg.log_capacity = self.log_capacity
        cfg.threads_per_block = self.threads_per_block
        cfg.flags = self.flags & 0xFF
        cfg.lanes_per_replication = (self.flags >> 8) & 0xFF  # Python-level convention: flags = placement | lanes << 8
        cfg.stream = self.stream
        capi.check(L, L.qk_batch_create(self._model, C.byref(cfg), C.byref(self._batch)))
        for dist, v in self._tapes:
            for did in self._dist_ids.get(id(dist), []):
                capi.check(L, L.qk_batch_set_tape(self._batch, did, v.ctypes.data, v.shape[1]))
        capi.check(L, L.qk_batch_init(self._batch, self.t0))

    def close(self):
        if self._batch:
            if getattr(self, "_owns", True):  # a shard view of a MultiGpuSimulation does not own its batch
                self.L.qk_batch_destroy(self._batch)
            self._batch = C.c_void_p()
        if self._model:
            self.L.qk_model_destroy(self._model)
            self._model = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- stepping ---------------------------------------------------------------------------------------
    def step_until(self, t_ns):
        self._materialize()
        capi.check(self.L, self.L.qk_batch_step_until(self._batch, int(t_ns)))

    def step_events(self, n_events):
        self._materialize()
        capi.check(self.L, self.L.qk_batch_step_events(self._batch, int(n_events)))

    def step_events_chunked(self, n_events, chunk):
        """step_events cut into launches of `chunk` events (state round-trips HBM between launches); same results."""
        self._materialize()
        capi.check(self.L, self.L.qk_batch_step_events_chunked(self._batch, int(n_events), int(chunk)))

    def step_events_sliced(self, n_events, n_chunks):
        """step_events as wave-sized launches over (block, chunk) work items; same results."""
        self._materialize()
        capi.check(self.L, self.L.qk_batch_step_events_sliced(self._batch, int(n_events), int(n_chunks)))

    def synchronize(self):
        self._materialize()
        capi.check(self.L, self.L.qk_batch_synchronize(self._batch))

    def last_step_ms(self):
        ms, n = C.c_float(), C.c_uint32()
        capi.check(self.L, self.L.qk_batch_last_step_ms(self._batch, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def state_bytes(self):
        return self.info()["state_bytes_per_rep"]

    def info(self):
        """Launch geometry and memory footprint of the batch (qk_batch_info)."""
        self._materialize()
        bi = capi.BatchInfo()
        capi.check(self.L, self.L.qk_batch_info_get(self._batch, C.byref(bi)))
        return {f: getattr(bi, f) for f, _ in bi._fields_}

    # -- results ----------------------------------------------------------------------------------------
    def summary(self):
        self._materialize()
        s = capi.BatchSummary()
        capi.check(self.L, self.L.qk_batch_summary_get(self._batch, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def status(self, rep0=0, n=None):
        """(status, now_ns, n_events) arrays for replications [rep0, rep0+n)."""
        self._materialize()
        n = self.n_reps - rep0 if n is None else n
        st = np.zeros(n, dtype=np.uint32)
        now = np.zeros(n, dtype=np.uint64)
        ev = np.zeros(n, dtype=np.uint64)
        capi.check(self.L, self.L.qk_batch_read_status(self._batch, rep0, n, st.ctypes.data, now.ctypes.data,
                                                       ev.ctypes.data))
        return st, now, ev

    def _cid(self, component):
        if isinstance(component, ComponentModel):
            component = component.component
        return component if isinstance(component, int) else component._id

    def comp_stats(self, component, rep0=0, n=None):
        self._materialize()
        n = self.n_reps - rep0 if n is None else n
        out = np.zeros(n, dtype=capi.COMP_STATS_DTYPE)
        capi.check(self.L, self.L.qk_batch_comp_stats(self._batch, rep0, n, self._cid(component), out.ctypes.data))
        return out

    def reduce_stats(self):
        """Per-component summary over this batch's replications (exact integer sums, extrema, histograms)."""
        self._materialize()
        out = np.zeros(self.n_comp, dtype=capi.COMP_SUMMARY_DTYPE)
        capi.check(self.L, self.L.qk_batch_reduce_stats(self._batch, out.ctypes.data, self.n_comp))
        return out

    def read_log(self, replication):
        """All retained records of one replication in canonical order, as a numpy structured array."""
        self._materialize()
        n, tot = C.c_uint64(), C.c_uint64()
        capi.check(self.L, self.L.qk_batch_read_log(self._batch, replication, None, 0, C.byref(n), C.byref(tot)))
        kept = min(tot.value, self.log_capacity)
        buf = np.zeros(kept, dtype=capi.LOG_DTYPE)
        if kept:
            capi.check(self.L, self.L.qk_batch_read_log(self._batch, replication, buf.ctypes.data, kept, C.byref(n),
                                                        C.byref(tot)))
        return buf, tot.value

    def read_vector(self, component, replication):
        """Resource of a vector stock / in-flight vector of a vector process-like component (3 doubles)."""
        self._materialize()
        out = (C.c_double * 3)()
        capi.check(self.L, self.L.qk_batch_read_vector(self._batch, replication, self._cid(component), out))
        return list(out)

    def read_resource(self, component, replication):
        """... of any resource type: one double per field (a user-defined resource: all of its fields)."""
        self._materialize()
        out, dim = (C.c_double * 8)(), C.c_uint32()
        capi.check(self.L, self.L.qk_batch_read_vector_n(self._batch, replication, self._cid(component), out, 8,
                                                         C.byref(dim)))
        return list(out[: dim.value])

    def read_stock(self, component, replication):
        self._materialize()
        n = C.c_uint64()
        cid = self._cid(component)
        capi.check(self.L, self.L.qk_batch_read_stock(self._batch, replication, cid, None, 0, C.byref(n)))
        buf = np.zeros(max(n.value, 1), dtype=np.uint32)
        capi.check(self.L, self.L.qk_batch_read_stock(self._batch, replication, cid, buf.ctypes.data, buf.size,
                                                      C.byref(n)))
        return buf[: n.value]

    def read_containers(self, component, replication):
        """Containers held by a container stock (FIFO order) or a container process-like component: (handles,
        resources[n, 3]) with NaN rows for containers that carry nothing (resource None)."""
        self._materialize()
        n = C.c_uint64()
        cid = self._cid(component)
        capi.check(self.L, self.L.qk_batch_read_containers(self._batch, replication, cid, None, None, 0, C.byref(n)))
        h = np.zeros(max(n.value, 1), dtype=np.uint32)
        r = np.zeros((max(n.value, 1), 3), dtype=np.uint64)
        capi.check(self.L, self.L.qk_batch_read_containers(self._batch, replication, cid, h.ctypes.data, r.ctypes.data,
                                                           h.size, C.byref(n)))
        return h[: n.value], r[: n.value]

    def read_logs(self, rep0, n):
        """Retained records of replications [rep0, rep0 + n): ONE gather kernel and one device-to-host copy.
        Returns (list of record arrays, oldest first; array of records ever written)."""
        self._materialize()
        buf = np.zeros((n, max(self.log_capacity, 1)), dtype=capi.LOG_DTYPE)
        kept = np.zeros(n, dtype=np.uint32)
        total = np.zeros(n, dtype=np.uint64)
        capi.check(self.L, self.L.qk_batch_read_logs(self._batch, rep0, n, buf.ctypes.data, kept.ctypes.data,
                                                     total.ctypes.data))
        return [buf[i, : kept[i]] for i in range(n)], total

    def reduce_partials(self):
        """Packed partial sums of this batch, (n_comp, PARTIAL_WORDS) uint64 (quokka_b200.h QP_*)."""
        self._materialize()
        out = np.zeros((self.n_comp, capi.PARTIAL_WORDS), dtype=np.uint64)
        capi.check(self.L, self.L.qk_batch_reduce_partials(self._batch, out.ctypes.data, self.n_comp))
        return out

    # device-resident reduction for the multi-GPU path (see parallel.py)
    def reduce_partials_device(self, d_partials_ptr):
        self._materialize()
        capi.check(self.L, self.L.qk_batch_reduce_partials_device(self._batch, d_partials_ptr, self.n_comp))

    def combine_partials_device(self, d_parts_ptr, n_parts, d_out_ptr):
        self._materialize()
        capi.check(self.L, self.L.qk_batch_combine_partials_device(self._batch, d_parts_ptr, n_parts, d_out_ptr,
                                                                   self.n_comp))

    def histograms_device(self, d_partials_ptr, d_hist_ptr):
        self._materialize()
        capi.check(self.L, self.L.qk_batch_histograms_device(self._batch, d_partials_ptr, d_hist_ptr, self.n_comp))


# --------------------------------------------------------------------------------------- vector family
# (imported last: vector_api subclasses _Component / _ProcessLike defined above)
from . import vector_api as _va  # noqa: E402

_VECTOR_VARIANTS = _va.VECTOR_VARIANTS
_bind_vector_variant = _va.bind_variant
_vector_connection_ok = _va.connection_ok
for _v, (_cls, _dim, _ports) in _VECTOR_VARIANTS.items():
    ComponentModel._VARIANTS[_v] = _cls
    setattr(ComponentModel, _v, _variant_ctor(_v))
for _v in ("F64StockLogger", "F64ProcessLogger", "Vector3StockLogger", "Vector3ProcessLogger"):
    setattr(ComponentLogger, _v, staticmethod(lambda logger, _v=_v: ComponentLogger(_v, logger)))
from . import container_api as _ca  # noqa: E402

_CONTAINER_VARIANTS = _ca.CONTAINER_VARIANTS
_bind_container_variant = _ca.bind_variant
_container_connection_ok = _ca.connection_ok
_CONTAINER_ENV_TARGETS = _ca.ENV_TARGETS
for _v, (_cls, _kind) in _CONTAINER_VARIANTS.items():
    ComponentModel._VARIANTS[_v] = _cls
    setattr(ComponentModel, _v, _variant_ctor(_v))
for _v in ("Vector3ContainerStockLogger", "Vector3ContainerProcessLogger"):
    setattr(ComponentLogger, _v, staticmethod(lambda logger, _v=_v: ComponentLogger(_v, logger)))
Vector3Container, Vector3ContainerFactory = _ca.Vector3Container, _ca.Vector3ContainerFactory
ContainerLoadingProcess, ContainerUnloadingProcess = _ca.ContainerLoadingProcess, _ca.ContainerUnloadingProcess
Vector3, VectorStock, VectorProcess, VectorSource = _va.Vector3, _va.VectorStock, _va.VectorProcess, _va.VectorSource
VectorSink, VectorCombiner, VectorSplitter = _va.VectorSink, _va.VectorCombiner, _va.VectorSplitter
define_resource, define_vector_variants = _va.define_resource, _va.define_vector_variants
