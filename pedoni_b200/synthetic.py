"""Synthetic uniform crowd on a large open domain (BASELINE.json configs[4]).

Counter-based generator: every per-agent quantity is a pure function of (seed, agent id, stream)
through splitmix64, so any rank / the CPU oracle can regenerate identical bits for any id range.

  domain      square, side = cells_per_side * 1.4 m with cells_per_side = ceil(sqrt(N / density) / 1.4)
              (N = 10 M at 1 ped/m^2 -> 2260 cells -> 3164 m, 5 107 600 cells)
  positions   i.i.d. uniform in [2, side - 2]^2, velocity 0 (sfm.rs:53)
  v0          N(1.34, 0.26^2) (Box-Muller) clamped to [0.5, 2.2]
  destination {0, 1} with p = 1/2; waypoints are vertical lines at x = 1 and x = side - 1
  field       open domain, border ring only (field.rs:29-32). field(device=k) builds the maps with the
              library's device builder (what bench.py does: 1.1 s for 1.6e8 cells x 3 maps, where the
              reference's serial heap marching takes minutes). field() is a closed form for boxes without
              a GPU and small tests: for an axis-aligned full-height line source the first-order
              marching gives potential = h * |col - col_wp| and distance = h * (cells to the border
              ring); tests/test_field_builder.py checks it against the marching builder on a small
              domain (they differ near the diagonals of the distance map, where two walls are equally far).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .field import Field
from .scenario import FieldConfig, Scenario, WaypointConfig

SEED = 0x5EED0001
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """One splitmix64 output per input counter (uint64 array)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _uniform(seed: int, ids: np.ndarray, stream: int) -> np.ndarray:
    """float64 in [0, 1) from (seed, id, stream)."""
    with np.errstate(over="ignore"):
        ctr = (np.uint64(seed) ^ (ids.astype(np.uint64) * np.uint64(8) + np.uint64(stream)))
    return (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


@dataclass
class SyntheticCrowd:
    n: int
    density: float = 1.0
    seed: int = SEED
    neighbor_unit: float = 1.4
    field_unit: float = 0.25

    @property
    def cells_per_side(self) -> int:
        return max(8, math.ceil(math.sqrt(self.n / self.density) / self.neighbor_unit) + 1)

    @property
    def side(self) -> float:
        return float(np.float32(self.cells_per_side * self.neighbor_unit))

    def scenario(self) -> Scenario:
        s = self.side
        sc = Scenario(field=FieldConfig(size=(s, s)))
        sc.waypoints.append(WaypointConfig(line=((1.0, 0.0), (1.0, s)), width=1.0))
        sc.waypoints.append(WaypointConfig(line=((s - 1.0, 0.0), (s - 1.0, s)), width=1.0))
        return sc

    def agents(self, lo: int = 0, hi: int | None = None):
        """(pos[n,2], dest[n], vel[n,2], v0[n]) for agent ids [lo, hi)."""
        hi = self.n if hi is None else hi
        ids = np.arange(lo, hi, dtype=np.uint64)
        s = self.side
        x = 2.0 + _uniform(self.seed, ids, 0) * (s - 4.0)
        y = 2.0 + _uniform(self.seed, ids, 1) * (s - 4.0)
        pos = np.stack([x, y], 1).astype(np.float32)
        u1 = np.maximum(_uniform(self.seed, ids, 2), 1e-300)
        u2 = _uniform(self.seed, ids, 3)
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
        v0 = np.clip(1.34 + 0.26 * z, 0.5, 2.2).astype(np.float32)
        dest = (_uniform(self.seed, ids, 4) >= 0.5).astype(np.uint32)
        vel = np.zeros((hi - lo, 2), np.float32)
        return pos, dest, vel, v0

    def field(self, device=None) -> Field:
        """device=k: the field of this scenario built on GPU k by pedoni_field_build_device (what bench.py feeds
        the model). device=None: the closed-form open-domain field (see module docstring), for boxes without a
        GPU and for small tests; it equals the built one up to the corner effects of the border ring."""
        if device is not None:
            return Field.from_scenario(self.scenario(), self.field_unit, device=device)
        h = np.float32(self.field_unit)
        s = np.float32(self.side)
        n = int(math.ceil(float(s / h)))  # field.rs:25-26
        col = np.arange(n, dtype=np.float32)
        ring = np.minimum(col, np.float32(n - 1) - col)  # cells to the border ring along one axis
        dist = (np.minimum(ring[:, None], ring[None, :]) * h).astype(np.float32)
        pots = np.empty((2, n, n), np.float32)
        for k, xw in enumerate((1.0, float(s) - 1.0)):
            # the waypoint is rasterised as the OUTLINE of its width-1 rectangle (field.rs:66-88):
            # potential 0 on columns c_lo and c_hi, growing by h per cell away from them.
            c_lo = np.float32(math.floor((xw - 0.5) / float(h)))
            c_hi = np.float32(math.floor((xw + 0.5) / float(h)))
            row = np.where(col < c_lo, c_lo - col, np.where(col > c_hi, col - c_hi,
                                                            np.minimum(col - c_lo, c_hi - col))) * h
            # the border ring is obstacle: slowness 1e6 * h there (field.rs:102)
            row[0] = row[1] + np.float32(1e6) * h
            row[-1] = row[-2] + np.float32(1e6) * h
            pots[k] = row[None, :]
            pots[k, 0, :] = pots[k, 1, :] + np.float32(1e6) * h
            pots[k, -1, :] = pots[k, -2, :] + np.float32(1e6) * h
        exist = np.zeros((n, n), bool)
        exist[0, :] = exist[-1, :] = exist[:, 0] = exist[:, -1] = True
        return Field(unit=float(h), shape=(n, n), obstacle_exist=exist, distance_map=dist, potential_maps=pots)
