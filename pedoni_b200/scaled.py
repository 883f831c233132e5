"""BASELINE.json configs[3]: a shipped scenario "scaled to 1 M pedestrians"  — every
coordinate, every width and the field size multiplied by k, and N pedestrians seeded once, uniformly, in the
free space (lanes.toml x 46 -> a 2 760 m x 1 380 m field with a 368 m corridor; random.toml x 5 -> 1 000 m x
1 000 m with 1 000 obstacles of 25 m). Used by tests/test_gpu_scaled_scenarios.py and scripts/bench_scenarios.py."""
from __future__ import annotations

import numpy as np

from .scenario import FieldConfig, ObstacleConfig, Scenario, WaypointConfig


def scaled_scenario(sc: Scenario, k: float) -> Scenario:
    s = lambda p: (p[0] * k, p[1] * k)  # noqa: E731
    out = Scenario(field=FieldConfig(size=s(sc.field.size)))
    out.waypoints = [WaypointConfig(line=(s(w.line[0]), s(w.line[1])), width=w.width * k) for w in sc.waypoints]
    out.obstacles = [ObstacleConfig(line=(s(o.line[0]), s(o.line[1])), width=o.width * k) for o in sc.obstacles]
    return out


def seed_free_space(sc: Scenario, field, n: int, dests, seed: int, box=None, clearance: float = 0.6):
    """(pos, destination, desired_speed) of n pedestrians uniformly where the distance map says more than
    `clearance` metres to the nearest obstacle, inside `box` = (x0, y0, x1, y1) (default: the whole field)."""
    rng = np.random.default_rng(seed)
    x0, y0, x1, y1 = box or (1.0, 1.0, sc.field.size[0] - 1.0, sc.field.size[1] - 1.0)
    pos = np.empty((0, 2), np.float32)
    while len(pos) < n:
        p = np.stack([rng.uniform(x0, x1, n), rng.uniform(y0, y1, n)], 1).astype(np.float32)
        ij = np.floor(p / field.unit).astype(int)
        ok = field.distance_map[np.clip(ij[:, 1], 0, field.shape[0] - 1),
                                np.clip(ij[:, 0], 0, field.shape[1] - 1)] > clearance
        pos = np.concatenate([pos, p[ok]])[:n]
    dest = rng.choice(np.asarray(dests), n).astype(np.uint32)
    v0 = np.clip(rng.normal(1.34, 0.26, n), 0.5, 2.2).astype(np.float32)
    return pos, dest, v0


# the two configurations BASELINE.json names: (shipped scenario, k, destinations, seeding box as a function of the scenario)
MILLION = {
    "lanes": (46.0, (0, 1), lambda sc: (0.09 * sc.field.size[0], 0.5, 0.91 * sc.field.size[0], 8.0 * 46 - 0.5)),
    "random": (5.0, (0, 1, 2, 3), lambda sc: None),
}
