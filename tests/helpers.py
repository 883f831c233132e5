"""Shared builders for the parity tests (TEST INFRASTRUCTURE)."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

import oracle
from pedoni_b200 import Field, Scenario, SimulatorOptions
from pedoni_b200.scenario import FieldConfig, ObstacleConfig, WaypointConfig

# Stated fp32 tolerances (SURVEY.md §8d / BASELINE.json north_star "within a stated fp32 tolerance"):
# PEDONI_MATH_STRICT (IEEE ops, no FMA, reference summation order): measured drift ~1e-6 after 10 steps.
TOL_POS_ABS = 1e-4   # metres, per-step positions over a horizon of <= 10 steps
TOL_VEL_ABS = 1e-4   # m/s
# PEDONI_MATH_FAST (MUFU rcp/rsqrt/ex2 + FMA): ~1e-6 relative error per force evaluation, amplified by the
# chaotic dynamics of dense crowds (measured 2.5e-5 m / 8e-5..2e-4 m/s after 10-12 steps at 2.6 ped/m^2).
TOL_POS_ABS_FAST = 5e-4
TOL_VEL_ABS_FAST = 1e-3


def tolerances(math_mode):
    return (TOL_POS_ABS, TOL_VEL_ABS) if math_mode == 0 else (TOL_POS_ABS_FAST, TOL_VEL_ABS_FAST)


def scenario_of(size, obstacles=(), waypoints=()):
    sc = Scenario(field=FieldConfig(size=tuple(map(float, size))))
    for o in obstacles:
        sc.obstacles.append(ObstacleConfig(line=((o[0], o[1]), (o[2], o[3])), width=o[4]))
    for w in waypoints:
        sc.waypoints.append(WaypointConfig(line=((w[0], w[1]), (w[2], w[3])), width=w[4]))
    return sc


def arrays_of(sc):
    obs = np.array([[*o.line[0], *o.line[1], o.width] for o in sc.obstacles], np.float32).reshape(-1, 5)
    wps = np.array([[*w.line[0], *w.line[1], w.width] for w in sc.waypoints], np.float32).reshape(-1, 5)
    return obs, wps


def oracle_field(sc, unit=0.25) -> Field:
    obs, wps = arrays_of(sc)
    exist, dist, pots = oracle.field_build(sc.field.size, unit, obs, wps)
    return Field(unit=unit, shape=dist.shape, obstacle_exist=exist, distance_map=dist, potential_maps=pots)


def corridor_scenario(size=(60.0, 30.0)):
    """Two facing waypoint lines and two long walls: a lanes.toml-like counter-flow corridor."""
    w, h = size
    return scenario_of(size,
                       obstacles=[(0, 3.0, w, 3.0, 0.5), (0, h - 3.0, w, h - 3.0, 0.5), (w / 2, 10.0, w / 2, 14.0, 1.0)],
                       waypoints=[(4.0, 5.0, 4.0, h - 5.0, 1.0), (w - 4.0, 5.0, w - 4.0, h - 5.0, 1.0)])


def random_crowd(n, size, seed, n_dest=2, margin=2.0, speed=True):
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.uniform(margin, size[0] - margin, n), rng.uniform(margin, size[1] - margin, n)], 1)
    pos = pos.astype(np.float32)
    dest = rng.integers(0, n_dest, n).astype(np.uint32)
    v0 = np.clip(rng.normal(1.34, 0.26, n), 0.5, 2.2).astype(np.float32)
    vel = (rng.normal(0, 0.5, (n, 2)) if speed else np.zeros((n, 2))).astype(np.float32)
    return pos, dest, vel, v0


def make_pair(sc, field, options=None, math_mode=0, **kw):
    """(cuda model, oracle model) over the same scenario + field arrays."""
    from pedoni_b200 import SocialForceModelCuda
    options = options or SimulatorOptions()
    obs, _ = arrays_of(sc)
    cu = SocialForceModelCuda(options, sc, field, math_mode=math_mode, **kw)
    orc = oracle.OracleModel(sc.field.size, options.neighbor_grid_unit, field.unit, field.distance_map,
                             field.potential_maps, obstacles=obs, use_neighbor_grid=options.use_neighbor_grid,
                             use_distance_map=options.use_distance_map)
    return cu, orc


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---- shipped scenarios (tests/golden/scenarios.npz, made by tests/golden/make_scenarios.py) -----------
GOLDEN = __import__("pathlib").Path(__file__).resolve().parent / "golden"


def scenario_names():
    z = np.load(GOLDEN / "scenarios.npz")
    return sorted({k.split("/")[0] for k in z.files if "/" in k})


def load_scenario(name):
    """The reference's scenarios/<name>.toml as a Scenario object."""
    from pedoni_b200.scenario import PedestrianConfig, PedestrianSpawnConfig
    z = np.load(GOLDEN / "scenarios.npz")
    for a in z["__aliases__"]:
        alias, target = str(a).split("=")
        if alias == name:
            name = target
    sc = scenario_of(z[f"{name}/size"], obstacles=z[f"{name}/obstacles"].tolist(),
                     waypoints=z[f"{name}/waypoints"].tolist())
    for origin, dest, kind, value in z[f"{name}/pedestrians"]:
        spawn = PedestrianSpawnConfig(kind="once", count=int(value)) if kind == 1 else \
            PedestrianSpawnConfig(kind="periodic", frequency=float(value))
        sc.pedestrians.append(PedestrianConfig(origin=int(origin), destination=int(dest), spawn=spawn))
    return sc


class OracleAdapter:
    """The CPU oracle behind the model surface `Simulator` drives (spawn_arrays / rebuild / step / ...)."""

    def __init__(self, options, scenario, field):
        obs, _ = arrays_of(scenario)
        self.m = oracle.OracleModel(scenario.field.size, options.neighbor_grid_unit, field.unit, field.distance_map,
                                    field.potential_maps, obstacles=obs, use_neighbor_grid=options.use_neighbor_grid,
                                    use_distance_map=options.use_distance_map)
        self._pending = None

    def spawn_arrays(self, pos, dest, v0):
        assert self._pending is None
        self._pending = (pos, dest, v0)

    def rebuild(self):
        if self._pending is None:
            self.m.spawn()
        else:
            self.m.spawn(*self._pending)
        self._pending = None

    def step(self):
        self.m.update()

    def get_pedestrian_count(self):
        return self.m.count()

    def download(self, vel=True, v0=True):
        return self.m.get()

    def cell_table(self):
        return self.m.indices()


def simulator_pair(name_or_scenario, seed=1, math_mode=0, options=None, unit=0.25, **cuda_kw):
    """(cuda Simulator, oracle Simulator) over the same scenario, field arrays and seeded spawn stream."""
    from pedoni_b200 import SocialForceModelCuda
    from pedoni_b200.simulator import Simulator
    sc = load_scenario(name_or_scenario) if isinstance(name_or_scenario, str) else name_or_scenario
    options = options or SimulatorOptions(field_grid_unit=unit)
    field = oracle_field(sc, options.field_grid_unit)
    cu = Simulator(options, sc, field, SocialForceModelCuda(options, sc, field, math_mode=math_mode, **cuda_kw), seed=seed)
    orc = Simulator(options, sc, field, OracleAdapter(options, sc, field), seed=seed)
    return cu, orc


# ---- fast math at scale: the anisotropy switch is a discontinuity of the model ---------------------------------------
def pair_force_full(pos_i, pos_j, vel_j):
    """sfm.rs:131-149 in float64, WITHOUT the anisotropy factor of sfm.rs:150-152: the force pedestrian j exerts on i."""
    d = np.asarray(pos_i, np.float64) - np.asarray(pos_j, np.float64)
    vj = np.asarray(vel_j, np.float64)
    dist = np.linalg.norm(d, axis=-1, keepdims=True)
    t1 = d - vj * 0.1
    t1len = np.linalg.norm(t1, axis=-1, keepdims=True)
    t2 = dist + t1len
    vl = np.linalg.norm(vj, axis=-1, keepdims=True) * 0.1
    b = np.sqrt(t2 * t2 - vl * vl) * 0.5
    nabla = t2 * (d / dist + t1 / t1len) / (4.0 * b)
    return 2.1 / 0.3 * np.exp(-b / 0.3) * nabla


def explain_fast_outliers(pre_pos, pre_vel, v0, vel_cuda, vel_oracle, pos_cuda, pos_oracle, tol_p, tol_v, window=60000):
    """One step from IDENTICAL state, fast math vs oracle. The reference halves a pair force when
    `e . (-f) < |f| cos(phi)` (sfm.rs:150-152): a discontinuity — a pair within rounding distance of the switch is
    decided differently by any implementation that is not bit-identical, and the pedestrian's acceleration then
    differs by exactly HALF OF THAT PAIR'S FORCE (up to 3.5 m/s^2), not by a rounding error. Returns
    (outliers, explained, clamped): pedestrians beyond the tolerance; those whose acceleration difference equals
    +-f_j/2 for one neighbour j (or +-f_j/2 +- f_k/2 for two) to 1 % + 5e-3 m/s^2; and those not checked because
    the speed clamp (sfm.rs:251) hid the acceleration on either side."""
    dv = np.abs(vel_cuda - vel_oracle).max(1)
    dp = np.abs(pos_cuda - pos_oracle).max(1)
    bad = np.nonzero((dv > tol_v) | (dp > tol_p) | ~np.isfinite(dv))[0]
    explained = clamped = 0
    for i in bad:
        vmax = 1.3 * float(v0[i]) * 0.9999
        if np.linalg.norm(vel_cuda[i]) >= vmax or np.linalg.norm(vel_oracle[i]) >= vmax:
            clamped += 1
            continue
        da = (vel_cuda[i].astype(np.float64) - vel_oracle[i].astype(np.float64)) / 0.1
        # the arrays are cell-sorted, row-major: everybody within 2 m sits within a few grid rows of index i
        lo, hi = max(0, i - window), min(len(pre_pos), i + window)
        d2 = ((pre_pos[lo:hi] - pre_pos[i]).astype(np.float64) ** 2).sum(1)
        near = lo + np.nonzero((d2 <= 4.0) & (d2 > 0.0))[0]
        if len(near) == 0:
            continue
        half = 0.5 * pair_force_full(pre_pos[i][None, :], pre_pos[near], pre_vel[near])
        cands = np.concatenate([half, -half])                                   # one pair decided differently
        if len(near) > 1:                                                      # or two
            a, b = np.triu_indices(len(near), 1)
            cands = np.concatenate([cands, half[a] + half[b], half[a] - half[b], -half[a] + half[b], -half[a] - half[b]])
        err = np.linalg.norm(cands - da[None, :], axis=1)
        scale = np.linalg.norm(cands, axis=1)
        explained += int((err <= 0.01 * scale + 5e-3).any())
    return len(bad), explained, clamped
