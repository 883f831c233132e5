"""GPU: BASELINE.json configs[3] — lanes.toml and random.toml scaled to one million pedestrians on one B200
(pedoni_b200/scaled.py: coordinates, widths and field size x k, 1 M pedestrians seeded once in free space;
field maps by the product's host builder, the same arrays for device and oracle).

Short horizon (lockstep, see tests/test_gpu_headline.py): the device is handed the oracle's state every tick;
population, cell table and order bit-exact, positions / velocities of one step within the fast-math tolerance.
Aggregate observables after a free run of both sides from the same seeded state: mean speed, population by
destination, and for the counter-flow corridor the lane observables (band-wise mean x-velocity -> lane count,
and the order parameter <((n_right - n_left) / (n_right + n_left))^2> over the bands). One million pedestrians
self-average, so a single seed is compared with tight tolerances, stated at the asserts."""
import os

import numpy as np
import pytest

import helpers
import oracle
from helpers import bits
from pedoni_b200 import PEDONI_MATH_FAST, Field, SimulatorOptions, SocialForceModelCuda, observables
from pedoni_b200.scaled import MILLION, scaled_scenario, seed_free_space

pytestmark = pytest.mark.gpu

N = 1_000_000
FREE_RUN_TICKS = 120


@pytest.fixture(scope="module", params=sorted(MILLION))
def million(request):
    name = request.param
    k, dests, box_of = MILLION[name]
    sc = scaled_scenario(helpers.load_scenario(name), k)
    opts = SimulatorOptions()
    field = Field.from_scenario(sc, opts.field_grid_unit)
    pos, dest, v0 = seed_free_space(sc, field, N, dests, seed=7, box=box_of(sc))
    oracle.lib().oracle_set_threads(os.cpu_count() or 1)
    return name, sc, opts, field, (pos, dest, v0)


def _pair(sc, opts, field):
    obs, _ = helpers.arrays_of(sc)
    cu = SocialForceModelCuda(opts, sc, field, math_mode=PEDONI_MATH_FAST, capacity=int(1.05 * N))
    orc = oracle.OracleModel(sc.field.size, opts.neighbor_grid_unit, field.unit, field.distance_map,
                             field.potential_maps, obstacles=obs)
    return cu, orc


def test_million_pedestrian_scenario_lockstep_vs_oracle(million):
    name, sc, opts, field, (pos, dest, v0) = million
    cu, orc = _pair(sc, opts, field)
    assert cu.field_textures()
    orc.spawn(pos, dest, v0)
    for _ in range(5):
        orc.update()
        orc.spawn()
    tol_p, tol_v = helpers.tolerances(PEDONI_MATH_FAST)
    worst_v = 0.0
    total_bad = explained = clamped = 0
    for tick in range(6):
        op, od, ov, o0 = orc.get()
        pre = (op, od, ov, o0)
        cu.upload_state(op, od, ov, o0)
        cu.rebuild()
        assert cu.get_pedestrian_count() == orc.count() > 0.98 * N
        np.testing.assert_array_equal(cu.cell_table(), orc.indices(), err_msg=f"{name} tick {tick}: cell table")
        cp, cd, cv, c0 = cu.download()
        np.testing.assert_array_equal(bits(cp), bits(op), err_msg=f"{name} tick {tick}: order")
        np.testing.assert_array_equal(cd, od)
        cu.step()
        orc.update()
        cp, cd, cv, c0 = cu.download()
        op, od, ov, o0 = orc.get()
        # fast math vs the discontinuous anisotropy factor: see tests/test_gpu_headline.py
        # (a NaN on one side only counts as an outlier)
        n_bad, n_explained, n_clamped = helpers.explain_fast_outliers(pre[0], pre[2], pre[3], cv, ov, cp, op, tol_p, tol_v)
        total_bad, explained, clamped = total_bad + n_bad, explained + n_explained, clamped + n_clamped
        assert n_bad <= 5e-4 * len(od), f"{name} tick {tick}: {n_bad} pedestrians beyond the fast-math tolerance"
        dv = np.abs(cv - ov).max(1)
        worst_v = max(worst_v, float(np.quantile(dv[np.isfinite(dv)], 0.999)))
        orc.spawn()
    print(f"{name} x{MILLION[name][0]:g}, 1 M pedestrians, lockstep over 6 ticks: 99.9th percentile of |dvel| <= "
          f"{worst_v:.2e} m/s; beyond the tolerance {total_bad} pedestrians in total, {explained} explained as half a "
          f"pair force, {clamped} speed-clamped")
    assert worst_v <= tol_v
    assert explained >= 0.9 * (total_bad - clamped), (total_bad, explained, clamped)
    cu.close()


def _lane_observables(pos, vel, y_range, bins):
    """(lane count, order parameter) from band-wise walkers to the right / to the left."""
    edges = np.linspace(y_range[0], y_range[1], bins + 1)
    idx = np.digitize(pos[:, 1], edges) - 1
    ok = (idx >= 0) & (idx < bins) & np.isfinite(vel[:, 0])
    right = np.bincount(idx[ok & (vel[:, 0] > 0)], minlength=bins).astype(float)
    left = np.bincount(idx[ok & (vel[:, 0] < 0)], minlength=bins).astype(float)
    tot = right + left
    phi = float(np.mean(((right - left) / np.maximum(tot, 1.0))[tot > 0] ** 2))
    return observables.lane_count(pos, vel, y_range, bins=bins, min_agents=20), phi


def test_million_pedestrian_scenario_observables_vs_oracle(million):
    name, sc, opts, field, (pos, dest, v0) = million
    cu, orc = _pair(sc, opts, field)
    cu.spawn_arrays(pos, dest, v0)
    cu.rebuild()
    orc.spawn(pos, dest, v0)
    for _ in range(FREE_RUN_TICKS):
        cu.step()
        cu.rebuild()
    orc.run(FREE_RUN_TICKS)  # update + spawn per tick (lib.rs:85-90 order: spawn_pedestrians, update_states)
    orc.spawn()
    cp, cd, cv, _ = cu.download()
    op, od, ov, _ = orc.get()
    fin = lambda v: v[np.isfinite(v).all(1)]  # noqa: E731
    n_cu, n_or = len(cd), len(od)
    s_cu, s_or = observables.mean_speed(fin(cv)), observables.mean_speed(fin(ov))
    print(f"{name} x{MILLION[name][0]:g} after {FREE_RUN_TICKS} ticks: population cuda {n_cu} / oracle {n_or}; "
          f"mean speed {s_cu:.4f} / {s_or:.4f} m/s")
    # population: pedestrians leave by reaching a destination, leaving the grid or turning NaN; 0.1 % of N
    assert abs(n_cu - n_or) <= 1e-3 * N
    np.testing.assert_allclose(np.bincount(cd, minlength=4), np.bincount(od, minlength=4), atol=1e-3 * N)
    # mean speed: 0.5 % relative
    assert abs(s_cu - s_or) <= 5e-3 * s_or
    # the device-side reduction (pedoni_observe) sees the same crowd as the download
    dev = cu.observe((0.0, float(sc.field.size[1])), bins=64)
    assert dev["count"] == n_cu and abs(dev["mean_speed"] - observables.mean_speed(cv[np.isfinite(cv).all(1)])) < 2e-3
    if name == "lanes":
        y_range, bins = (0.0, 8.0 * 46.0), 64  # the corridor, in bands of 5.75 m
        l_cu, phi_cu = _lane_observables(cp, cv, y_range, bins)
        l_or, phi_or = _lane_observables(op, ov, y_range, bins)
        print(f"lanes x46: lane count cuda {l_cu} / oracle {l_or}; order parameter {phi_cu:.4f} / {phi_or:.4f}")
        # lane count over 64 bands: within 15 % + 2; order parameter: within 0.01
        assert abs(l_cu - l_or) <= 0.15 * l_or + 2
        assert abs(phi_cu - phi_or) <= 0.01
        lanes_dev = observables.lane_count_from_bins(dev["bin_mean_vx"], dev["bin_count"], 20)
        assert abs(lanes_dev - observables.lane_count(cp, cv, (0.0, float(sc.field.size[1])), bins=64, min_agents=20)) <= 1
    cu.close()
