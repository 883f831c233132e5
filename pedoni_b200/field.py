"""`Field` — mirrors pedoni-simulator/src/field.rs:194-205: the one-time precompute whose OUTPUTS
(distance map + one potential map per waypoint) are inputs of the per-timestep hot path."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Field:
    unit: float                  # field.rs:196
    shape: tuple                 # (fy, fx), field.rs:198
    obstacle_exist: np.ndarray   # bool (fy, fx), field.rs:200
    distance_map: np.ndarray     # f32 (fy, fx), field.rs:202
    potential_maps: np.ndarray   # f32 (n_waypoints, fy, fx), field.rs:204

    @staticmethod
    def from_scenario(scenario, unit: float) -> "Field":
        """field.rs:220-232, computed by the host-side C++ builder in libpedoni_cuda.so."""
        from . import host
        return host.field_from_scenario(scenario, unit)
