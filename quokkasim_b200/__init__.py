"""quokkasim_b200 -- B200-native batched-replication executor for QuokkaSim stock-and-flow models.

Host-side mirror of the reference's model-builder API (api.py) over a C-ABI shared library
(libquokka_b200.so, include/quokka_b200.h) whose hot path is hand-written sm_100a CUDA.
There is no CPU execution path in this package.
"""
from .api import (  # noqa: F401
    BasicEnvironment, BasicEnvironmentLogger, BasicEnvironmentState, BatchedSimInit, BatchedSimulation,
    ComponentConnectionError, ComponentLogger, ComponentModel, ComponentModelAddress, DelayMode, DelayModeChange,
    DiscreteParallelProcess, DiscreteProcess, DiscreteProcessLogger, DiscreteSink, DiscreteSource, DiscreteStock,
    DiscreteStockLogger, Distribution, DistributionConfig, DistributionFactory, DistributionParametersError, Duration,
    ItemDeque, Mailbox, MonotonicTime, ScheduledEvent, ScheduledEventConfig, Scheduler, StringItemFactory,
    connect_components, connect_logger, create_scheduled_event, register_component,
)
from ._capi import QkError  # noqa: F401
from .api import VectorProcessLogger, VectorStockLogger  # noqa: F401
from .vector_api import (  # noqa: F401
    Vector3, VectorCombiner, VectorProcess, VectorSink, VectorSource, VectorSplitter, VectorStock,
    define_resource, define_vector_variants,
)

from .api import ContainerProcessLogger, ContainerStockLogger  # noqa: F401
from .container_api import (  # noqa: F401
    ContainerLoadingProcess, ContainerUnloadingProcess, Vector3Container, Vector3ContainerFactory,
)

__version__ = "0.1.0"
