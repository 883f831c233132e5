"""pedoni_b200 — B200-native backend for ONE path of qt2/pedoni: the per-timestep pedestrian update
(neighbor-grid rebuild, pair/wall repulsion, navigation-field steering, integration) behind the
reference's `PedestrianModel` plugin trait, through the C ABI in include/pedoni_cuda.h.

The product is libpedoni_cuda.so (hand-written sm_100a CUDA + a C++ host layer); this Python package
is the ctypes harness the tests and bench.py drive it with. There is no CPU fallback.
"""
from .model import Pedestrian, SlabGroup, SocialForceModelCuda, comm_unique_id, slab_rows  # noqa: F401
from .options import Backend, SimulatorOptions  # noqa: F401
from .scenario import Scenario  # noqa: F401
from .field import Field  # noqa: F401
from .simulator import Simulator, SpawnStream, StepMetrics  # noqa: F401
from . import observables  # noqa: F401
from ._capi import (PEDONI_MATH_FAST, PEDONI_MATH_STRICT, PedoniError)  # noqa: F401
