"""`SimulatorOptions` / `Backend` — mirrors pedoni-simulator/src/lib.rs:108-142."""
from __future__ import annotations

import enum
from dataclasses import dataclass


class Backend(enum.Enum):  # lib.rs:138-142 plus the new arm this repo adds
    CPU = "cpu"
    GPU = "gpu"
    CUDA = "cuda"


@dataclass
class SimulatorOptions:  # defaults: lib.rs:124-135
    backend: Backend = Backend.CUDA
    neighbor_grid_unit: float = 1.4
    field_grid_unit: float = 0.25
    use_neighbor_grid: bool = True
    use_distance_map: bool = True
    gpu_work_size: int = 64  # parsed but never applied in the reference (args.rs:40 vs :47-66); unused here too
