"""`Field` — mirrors pedoni-simulator/src/field.rs:194-205: the one-time precompute whose OUTPUTS
(distance map + one potential map per waypoint) are inputs of the per-timestep hot path.

Building it (`Field::from_scenario`, field.rs:220-232: outline rasterisation + fast marching) is the
step BEFORE the path and stays with the caller, as it does for the Rust trait (lib.rs:30 builds the
field, then hands `&Field` to `PedestrianModel::new`). SURVEY.md section 8(f1) lists a product-side
builder as the first "next" row; until then the tests use the oracle's restatement
(oracle/field_oracle.cpp) and bench.py the closed form of pedoni_b200/synthetic.py."""
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
