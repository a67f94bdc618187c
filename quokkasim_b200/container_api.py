"""Host-side mirror of the reference's container family (quokkasim/src/components/vector_container.rs) and the
ComponentModel variants built on it (core.rs:549-555).  Pure bookkeeping, like api.py / vector_api.py.

A Vector3Container {id, resource: Option<Vector3>, capacity} is a discrete item.  Vector3ContainerStock / Process /
ParallelProcess / Source / Sink are the DISCRETE component classes of api.py wrapped in a container variant (exactly as
in Rust, where they are DiscreteStock<Vector3Container> etc.); ContainerLoadingProcess / ContainerUnloadingProcess are
defined here.  The F64Container variants exist in the reference's enum but have no connect or register arms
(core.rs:923-995, 1186-1192), so they cannot be part of a model and are not mirrored.
"""
from . import _capi as capi
from .api import (DiscreteParallelProcess, DiscreteProcess, DiscreteSink, DiscreteSource, DiscreteStock, Distribution,
                  _ProcessLike)
from .vector_api import Vector3


class Vector3Container:
    """vector_container.rs:71-76"""

    def __init__(self, id="", resource=None, capacity=0.0):
        self.id = str(id)
        self.resource = resource if (resource is None or isinstance(resource, Vector3)) else Vector3(resource)
        self.capacity = float(capacity)

    @staticmethod
    def default():
        return Vector3Container()

    def __repr__(self):
        return "Vector3Container(%r, %r, %r)" % (self.id, self.resource, self.capacity)


class Vector3ContainerFactory:
    """vector_container.rs:126-156: items are "{prefix}{index:0num_digits}" (no underscore, unlike StringItemFactory)."""

    def __init__(self, prefix="", next_index=0, num_digits=0, create_quantity_distr=None,
                 create_component_split=(0.0, 0.0, 0.0), capacity=0.0):
        self.prefix = prefix
        self.next_index = int(next_index)
        self.num_digits = int(num_digits)
        self.create_quantity_distr = create_quantity_distr
        self.create_component_split = [float(x) for x in create_component_split]
        self.capacity = float(capacity)

    @staticmethod
    def default():
        return Vector3ContainerFactory()

    def item_name(self, index):
        return "%s%s" % (self.prefix, str(self.next_index + index).rjust(self.num_digits, "0"))


class _ContainerTransfer(_ProcessLike):
    def __init__(self):
        super().__init__()
        self.max_process_count = 1

    def with_max_process_count(self, n):
        self.max_process_count = int(n)
        return self

    def with_max_process_count_inplace(self, n):
        self.max_process_count = int(n)


class ContainerLoadingProcess(_ContainerTransfer):
    """vector_container.rs:158-233.  process_capacity_ratio_distr (default Constant(1.)): proportion of the container's
    free capacity that is loaded."""

    KIND = capi.CONTAINER_LOAD_PROCESS
    DEFAULT_NAME = "ContainerLoadingProcess"

    def __init__(self):
        super().__init__()
        self.element_code = "CLP"  # vector_container.rs:205

    @property
    def process_capacity_ratio_distr(self):
        return self.process_quantity_distr

    def with_process_capacity_ratio_distr(self, d):
        self.process_quantity_distr = d  # second distribution slot of the C ABI (qk_model_set_dist which = 1)
        return self

    def with_process_capacity_ratio_distr_inplace(self, d):
        self.process_quantity_distr = d


class ContainerUnloadingProcess(_ContainerTransfer):
    """vector_container.rs:445-520.  process_quantity_ratio_distr (default Constant(1.)): proportion of the held amount
    that is unloaded."""

    KIND = capi.CONTAINER_UNLOAD_PROCESS
    DEFAULT_NAME = "ContainerUnloadingProcess"

    @property
    def process_quantity_ratio_distr(self):
        return self.process_quantity_distr

    def with_process_quantity_ratio_distr(self, d):
        self.process_quantity_distr = d
        return self

    def with_process_quantity_ratio_distr_inplace(self, d):
        self.process_quantity_distr = d


# ComponentModel variant -> (class, C-ABI kind)
CONTAINER_VARIANTS = {
    "Vector3ContainerStock": (DiscreteStock, capi.CONTAINER_STOCK),
    "Vector3ContainerProcess": (DiscreteProcess, capi.CONTAINER_PROCESS),
    "Vector3ContainerParallelProcess": (DiscreteParallelProcess, capi.CONTAINER_PARALLEL_PROCESS),
    "Vector3ContainerSource": (DiscreteSource, capi.CONTAINER_SOURCE),
    "Vector3ContainerSink": (DiscreteSink, capi.CONTAINER_SINK),
    "Vector3ContainerLoadProcess": (ContainerLoadingProcess, capi.CONTAINER_LOAD_PROCESS),
    "Vector3ContainerUnloadProcess": (ContainerUnloadingProcess, capi.CONTAINER_UNLOAD_PROCESS),
}


def bind_variant(variant, component):
    """What the variant's type parameters fix in Rust: the item type is Vector3Container."""
    component._container_kind = CONTAINER_VARIANTS[variant][1]
    if isinstance(component, DiscreteStock):
        for it in component.resource:
            if not isinstance(it, Vector3Container):
                raise TypeError("Vector3ContainerStock holds Vector3Container items")
    if isinstance(component, DiscreteSource):
        if not component._factory_set:
            component.item_factory = Vector3ContainerFactory.default()  # F::default()
        elif not isinstance(component.item_factory, Vector3ContainerFactory):
            raise TypeError("Vector3ContainerSource takes a Vector3ContainerFactory")


_S = "Vector3ContainerStock"


def connection_ok(va, vb, n):
    """The container arms of connect_components (core.rs:923-995)."""
    if va == _S and vb in ("Vector3ContainerProcess", "Vector3ContainerParallelProcess", "Vector3ContainerSink",
                           "Vector3ContainerLoadProcess", "Vector3ContainerUnloadProcess"):
        return True
    if vb == _S and va in ("Vector3ContainerProcess", "Vector3ContainerParallelProcess", "Vector3ContainerSource",
                           "Vector3ContainerLoadProcess", "Vector3ContainerUnloadProcess"):
        return True
    if va == "Vector3Stock" and vb == "Vector3ContainerLoadProcess":
        return True
    if va == "Vector3ContainerUnloadProcess" and vb == "Vector3Stock":
        return True
    return False


# BasicEnvironment arms (core.rs:1149-1172): no ParallelProcess arm
ENV_TARGETS = {"Vector3ContainerProcess", "Vector3ContainerSource", "Vector3ContainerSink",
               "Vector3ContainerLoadProcess", "Vector3ContainerUnloadProcess"}
