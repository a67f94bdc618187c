"""Host-side mirror of the reference's continuous components (quokkasim/src/components/vector.rs) and the
resource types they carry (core.rs:7-128).  Pure bookkeeping, like api.py: the objects only hold the fields that
BatchedSimulation turns into C-ABI rows (qk_model_set_vector / qk_model_set_ports)."""
from . import _capi as capi
from .api import _Component, _ProcessLike


class Vector3:
    """core.rs:47-128"""

    def __init__(self, values=(0.0, 0.0, 0.0)):
        self.values = [float(x) for x in values]
        assert len(self.values) == 3

    @staticmethod
    def default():
        return Vector3()

    def total(self):
        return (self.values[0] + self.values[1]) + self.values[2]

    def __repr__(self):
        return "Vector3(%r)" % (self.values,)


MAX_DIM = 8  # QK_VEC_MAX_DIM


class _Resource:
    """Base of the user-defined resource types made by define_resource()."""

    NAME = None
    FIELDS = ()
    TOTAL_MASK = 0

    def __init__(self, *args, **kw):
        if args and kw:
            raise TypeError("%s takes its fields by position or by name, not both" % self.NAME)
        if kw:
            unknown = set(kw) - set(self.FIELDS)
            if unknown:
                raise TypeError("%s has no field %s" % (self.NAME, sorted(unknown)[0]))
            args = [kw.get(f, 0.0) for f in self.FIELDS]  # ..Default::default()
        if not args:
            args = [0.0] * len(self.FIELDS)
        if len(args) != len(self.FIELDS):
            raise TypeError("%s has %d fields" % (self.NAME, len(self.FIELDS)))
        self.values = [float(x) for x in args]

    @classmethod
    def default(cls):
        return cls()

    def total(self):
        t = None
        for k, v in enumerate(self.values):  # the counted fields, left to right
            if (self.TOTAL_MASK >> k) & 1:
                t = v if t is None else t + v
        return t

    def __getattr__(self, name):
        fields = type(self).FIELDS
        if name in fields:
            return self.values[fields.index(name)]
        raise AttributeError(name)

    def __repr__(self):
        return "%s(%s)" % (self.NAME, ", ".join("%s=%r" % fv for fv in zip(self.FIELDS, self.values)))


def define_resource(name, fields, total=None):
    """A user-defined resource type: what `struct IronOre { fe, other_elements, magnetite, hematite, limonite }` with its
    ResourceAdd / ResourceRemove / ResourceTotal / ResourceMultiply impls is in the reference
    (quokkasim_examples/src/bin/iron_ore.rs:21-96).  `fields`: up to 8 f64 field names; `total`: the fields
    ResourceTotal sums, in field order (default: all).  Add is field-wise, remove(q) takes q / total() of every field."""
    fields = tuple(fields)
    if not 1 <= len(fields) <= MAX_DIM or len(set(fields)) != len(fields):
        raise TypeError("a resource has 1..%d distinct fields" % MAX_DIM)
    counted = fields if total is None else tuple(total)
    if not counted or any(f not in fields for f in counted):
        raise TypeError("total= names fields of the resource")
    mask = 0
    for f in counted:
        mask |= 1 << fields.index(f)
    return type(name, (_Resource,), {"NAME": name, "FIELDS": fields, "TOTAL_MASK": mask})


def _as_list(resource):
    if isinstance(resource, (Vector3, _Resource)):
        return list(resource.values)
    return [float(resource)]


class VectorStock(_Component):
    """vector.rs:42-77: low_capacity / max_capacity default to 0.0, resource to Default::default()"""

    KIND = capi.VECTOR_STOCK
    DEFAULT_NAME = "VectorStock"

    def __init__(self):
        super().__init__()
        self.low_capacity = 0.0
        self.max_capacity = 0.0
        self.resource = None  # f64 0.0 or Vector3::default(), decided by the ComponentModel variant
        self.dim = None
        self.total_mask = None  # user-defined resources only (define_resource)
        self.resource_type = None

    def with_low_capacity(self, v):
        self.low_capacity = float(v)
        return self

    def with_max_capacity(self, v):
        self.max_capacity = float(v)
        return self

    def with_low_capacity_inplace(self, v):
        self.low_capacity = float(v)

    def with_max_capacity_inplace(self, v):
        self.max_capacity = float(v)

    def with_initial_resource(self, resource):
        self.resource = resource
        return self

    def vector_rows(self):
        init = _as_list(self.resource) if self.resource is not None else [0.0] * self.dim
        return self.dim, self.low_capacity, self.max_capacity, init


class _VectorProcessLike(_ProcessLike):
    def __init__(self):
        super().__init__()
        self.dim = None
        self.total_mask = None
        self.resource_type = None

    def vector_rows(self):
        return self.dim, 0.0, 0.0, [0.0] * self.dim


class VectorProcess(_VectorProcessLike):
    KIND = capi.VECTOR_PROCESS
    DEFAULT_NAME = "VectorProcess"


class VectorSource(_VectorProcessLike):
    """vector.rs:1324-1400: creates `source_vector` scaled to the sampled quantity"""

    KIND = capi.VECTOR_SOURCE
    DEFAULT_NAME = "VectorSource"

    def __init__(self):
        super().__init__()
        self.source_vector = None

    def with_source_vector(self, v):
        self.source_vector = v
        return self

    def with_source_vector_inplace(self, v):
        self.source_vector = v

    def vector_rows(self):
        init = _as_list(self.source_vector) if self.source_vector is not None else [0.0] * self.dim
        return self.dim, 0.0, 0.0, init


class VectorSink(_VectorProcessLike):
    KIND = capi.VECTOR_SINK
    DEFAULT_NAME = "VectorSink"


class VectorCombiner(_VectorProcessLike):
    """vector.rs:726-820; `M` (number of upstream ports) is fixed by the ComponentModel variant (Combiner1..5)"""

    KIND = capi.VECTOR_COMBINER
    DEFAULT_NAME = "VectorCombiner"

    def __init__(self):
        super().__init__()
        self.n_ports = None
        self.split_ratios = None

    def with_split_ratios(self, ratios):
        self.split_ratios = [float(x) for x in ratios]  # unused by the reference, kept for parity
        return self

    def with_split_ratios_inplace(self, ratios):
        self.split_ratios = [float(x) for x in ratios]


class VectorSplitter(_VectorProcessLike):
    """vector.rs:1026-1122; default split_ratios = [1/N; N]"""

    KIND = capi.VECTOR_SPLITTER
    DEFAULT_NAME = "VectorSplitter"

    def __init__(self):
        super().__init__()
        self.n_ports = None
        self.split_ratios = None

    def with_split_ratios(self, ratios):
        self.split_ratios = [float(x) for x in ratios]
        return self

    def with_split_ratios_inplace(self, ratios):
        self.split_ratios = [float(x) for x in ratios]


# ComponentModel variants of core.rs:505-533 -> (class, dim, ports)
VECTOR_VARIANTS = {}
for _prefix, _dim in (("F64", 1), ("Vector3", 3)):
    VECTOR_VARIANTS[_prefix + "Stock"] = (VectorStock, _dim, None)
    VECTOR_VARIANTS[_prefix + "Process"] = (VectorProcess, _dim, None)
    VECTOR_VARIANTS[_prefix + "Source"] = (VectorSource, _dim, None)
    VECTOR_VARIANTS[_prefix + "Sink"] = (VectorSink, _dim, None)
    for _n in range(1, 6):
        VECTOR_VARIANTS["%sCombiner%d" % (_prefix, _n)] = (VectorCombiner, _dim, _n)
        VECTOR_VARIANTS["%sSplitter%d" % (_prefix, _n)] = (VectorSplitter, _dim, _n)


RESOURCE_OF = {}  # variant -> resource class, for the variants made by define_vector_variants
PREFIXES = ["F64", "Vector3"]
PREFIX_OF = {v: ("F64" if v.startswith("F64") else "Vector3") for v in VECTOR_VARIANTS}


def define_vector_variants(prefix, resource_cls):
    """The ComponentModel / ComponentLogger variants a user adds to define_model_enums! for an own resource type, with the
    connection arms of `impl CustomComponentConnection` (iron_ore.rs:272-306): `<prefix>Stock`, `<prefix>Process`,
    `<prefix>Source`, `<prefix>Sink`, `<prefix>Combiner1..5`, `<prefix>Splitter1..5`, `<prefix>StockLogger`,
    `<prefix>ProcessLogger`.  They connect among themselves like the built-in vector variants (core.rs:572-883)."""
    from . import api

    if not (isinstance(resource_cls, type) and issubclass(resource_cls, _Resource)):
        raise TypeError("define_vector_variants takes a class made by define_resource")
    if prefix in PREFIXES or any((prefix + k) in api.ComponentModel._VARIANTS for k in ("Stock", "Process", "Source", "Sink")):
        raise ValueError("variant prefix %r collides with an existing one" % prefix)
    PREFIXES.append(prefix)
    dim = len(resource_cls.FIELDS)
    names = [(prefix + "Stock", VectorStock, None), (prefix + "Process", VectorProcess, None),
             (prefix + "Source", VectorSource, None), (prefix + "Sink", VectorSink, None)]
    for n in range(1, 6):
        names.append(("%sCombiner%d" % (prefix, n), VectorCombiner, n))
        names.append(("%sSplitter%d" % (prefix, n), VectorSplitter, n))
    for v, cls, ports in names:
        VECTOR_VARIANTS[v] = (cls, dim, ports)
        PREFIX_OF[v] = prefix
        RESOURCE_OF[v] = resource_cls
        api.ComponentModel._VARIANTS[v] = cls
        setattr(api.ComponentModel, v, api._variant_ctor(v))
    api.VectorStockLogger.ACCEPTS = api.VectorStockLogger.ACCEPTS + (prefix + "Stock",)
    api.VectorProcessLogger.ACCEPTS = api.VectorProcessLogger.ACCEPTS + tuple(v for v, _, _ in names[1:])
    for v, lcls in ((prefix + "StockLogger", api.VectorStockLogger), (prefix + "ProcessLogger", api.VectorProcessLogger)):
        api.ComponentLogger._VARIANTS[v] = lcls
        setattr(api.ComponentLogger, v, staticmethod(lambda logger, _v=v: api.ComponentLogger(_v, logger)))
    return resource_cls


def bind_variant(variant, component):
    """What the enum variant's type parameters fix in Rust: resource type (f64 / Vector3 / a user's struct) and port
    count."""
    cls, dim, ports = VECTOR_VARIANTS[variant]
    component.dim = dim
    rcls = RESOURCE_OF.get(variant)
    if rcls is not None:
        component.total_mask = rcls.TOTAL_MASK
        component.resource_type = rcls
        for what in ("resource", "source_vector"):
            r = getattr(component, what, None)
            if r is not None and not isinstance(r, rcls):
                raise TypeError("%s carries %s resources" % (variant, rcls.NAME))
    if ports is not None:
        component.n_ports = ports
        if component.split_ratios is None:
            component.split_ratios = [1.0 / ports] * ports
        if len(component.split_ratios) != ports:
            raise TypeError("%s needs %d split ratios" % (variant, ports))
    if isinstance(component, VectorStock) and component.resource is not None:
        if len(_as_list(component.resource)) != dim:
            raise TypeError("%s holds a %s resource" % (variant, "f64" if dim == 1 else "Vector3"))
    if isinstance(component, VectorSource) and component.source_vector is not None:
        if len(_as_list(component.source_vector)) != dim:
            raise TypeError("%s creates a %s resource" % (variant, "f64" if dim == 1 else "Vector3"))


def connection_ok(va, vb, n):
    """The vector arms of connect_components (core.rs:572-883): same resource type on both sides; Stock -> CombinerM
    and SplitterN -> Stock take Some(n)."""
    prefix = PREFIX_OF.get(va)
    if prefix is not None and PREFIX_OF.get(vb) == prefix:
        a, b = va[len(prefix):], vb[len(prefix):]
        if a == "Stock" and b in ("Process", "Sink"):
            return True
        if a == "Stock" and b.startswith("Splitter"):
            return True
        if a == "Stock" and b.startswith("Combiner"):
            return n is not None and 0 <= n < int(b[-1])
        if b == "Stock" and a in ("Process", "Source"):
            return True
        if b == "Stock" and a.startswith("Combiner"):
            return True
        if b == "Stock" and a.startswith("Splitter"):
            return n is not None and 0 <= n < int(a[-1])
    return False
