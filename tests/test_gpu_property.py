"""GPU: randomised parity (hypothesis). Random domain sizes, grid units, obstacle / waypoint layouts, wall
variants and crowds (including pedestrians outside the grid, on obstacles and on their destination):
the rebuild is bit-exact against the oracle and a few ticks stay within the stated fp32 tolerance, on a
whole-domain handle and on a random number of slabs."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import helpers
from helpers import bits
from pedoni_b200 import PEDONI_MATH_FAST, PEDONI_MATH_STRICT, SimulatorOptions, SlabGroup, SocialForceModelCuda

pytestmark = pytest.mark.gpu


@st.composite
def worlds(draw):
    w = draw(st.floats(12.0, 60.0))
    h = draw(st.floats(12.0, 40.0))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    n_obs = draw(st.integers(0, 6))
    obstacles = [(*rng.uniform(1, [w - 1, h - 1]), *rng.uniform(1, [w - 1, h - 1]), float(rng.uniform(0.1, 2.0)))
                 for _ in range(n_obs)]
    n_wp = draw(st.integers(1, 3))
    waypoints = [(*rng.uniform(2, [w - 2, h - 2]), *rng.uniform(2, [w - 2, h - 2]), float(rng.uniform(0.5, 2.0)))
                 for _ in range(n_wp)]
    return dict(size=(w, h), obstacles=obstacles, waypoints=waypoints, seed=seed,
                neighbor_unit=draw(st.sampled_from([1.0, 1.4, 2.0, 2.5])),
                field_unit=draw(st.sampled_from([0.25, 0.3, 0.5])),
                use_distance_map=draw(st.booleans()),
                # up to 2.5 pedestrians / m^2: beyond that a crowd is a crush whose dynamics amplify the last bits
                # of fast math past any per-step tolerance within a tick or two (tests/test_gpu_edge_cases.py
                # covers the dense code paths on their own terms)
                n_agents=draw(st.integers(1, max(2, min(1500, int(2.5 * w * h))))),
                mode=draw(st.sampled_from([PEDONI_MATH_STRICT, PEDONI_MATH_FAST])),
                slabs=draw(st.integers(1, 3)))


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(worlds())
def test_random_worlds_match_the_oracle(wd):
    sc = helpers.scenario_of(wd["size"], obstacles=wd["obstacles"], waypoints=wd["waypoints"])
    opts = SimulatorOptions(neighbor_grid_unit=wd["neighbor_unit"], field_grid_unit=wd["field_unit"],
                            use_distance_map=wd["use_distance_map"])
    field = helpers.oracle_field(sc, wd["field_unit"])
    _, orc = helpers.make_pair(sc, field, options=opts, math_mode=wd["mode"])
    ny = orc.grid_shape()[0]
    n_slabs = wd["slabs"] if ny // max(wd["slabs"], 1) >= 2 else 1
    cu = (SocialForceModelCuda(opts, sc, field, math_mode=wd["mode"]) if n_slabs == 1
          else SlabGroup(opts, sc, field, n_slabs, math_mode=wd["mode"]))
    # Strict mode: a few pedestrians start outside the grid, or inside it but off the field maps (truncation
    # toward zero keeps x in (-unit, 0) in cell 0, neighbor_grid.rs:27) where every sample is built from 1e12
    # out-of-bounds taps — sums that cancel to 0 (-> NaN -> despawn) or to +-131072 depending on the last bit
    # of the position. Bit-identical arithmetic reproduces that; fast math, a few ulp away, legitimately lands
    # on the other side, so its crowd starts inside the maps.
    margin = -1.0 if wd["mode"] == PEDONI_MATH_STRICT else 0.6
    pos, dest, vel, v0 = helpers.random_crowd(wd["n_agents"], sc.field.size, seed=wd["seed"], n_dest=len(wd["waypoints"]),
                                              margin=margin)
    cu.upload_state(pos, dest, vel, v0)
    orc.set(pos, dest, vel, v0)
    tol_p, tol_v = helpers.tolerances(wd["mode"])
    for _ in range(3):
        cu.rebuild()
        orc.spawn()
        assert cu.get_pedestrian_count() == orc.count()
        np.testing.assert_array_equal(cu.cell_table(), orc.indices())
        cp, cd, cv, c0 = cu.download()
        op, od, ov, o0 = orc.get()
        np.testing.assert_array_equal(cd, od)
        np.testing.assert_array_equal(bits(c0), bits(o0))
        cu.step()
        orc.update()
    cp, cd, cv, _ = cu.download()
    op, od, ov, _ = orc.get()
    np.testing.assert_array_equal(np.isnan(cp).any(1), np.isnan(op).any(1))
    ok = ~np.isnan(op).any(1)
    if ok.any():
        # pedestrians starting inside walls see 1e6-scale gradients and speeds up to the clamp: scale the bound
        scale = max(1.0, float(np.abs(ov[ok]).max()))
        assert np.abs(cp[ok] - op[ok]).max() <= tol_p * scale * 4
        assert np.abs(cv[ok] - ov[ok]).max() <= tol_v * scale * 4
    cu.close()
