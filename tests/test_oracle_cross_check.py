"""CPU: the C++ oracle against a second restatement of the reference's step written independently in scalar
numpy-float32 Python (tests/py_restatement.py). Integer outputs — populations, `neighbor_grid_indices`,
order, destinations — must be identical; positions and velocities are bit-identical too, except where
glibc's expf (0.502 ulp) and a correctly rounded exp disagree in the last bit (about one call in 10^3)."""
import numpy as np
import pytest

import helpers
import oracle
from py_restatement import PyModel, bilinear, distance_from_line, sobel_filter


def test_samplers_bit_for_bit():
    rng = np.random.default_rng(1)
    grid = rng.normal(5, 3, (9, 11)).astype(np.float32)
    for _ in range(300):
        x, y = np.float32(rng.uniform(-2, 12)), np.float32(rng.uniform(-2, 10))  # includes out-of-bounds taps
        assert helpers.bits(bilinear(grid, x, y)) == helpers.bits(np.float32(oracle.bilinear(grid, x, y)))
        gx, gy = sobel_filter(grid, x, y)
        np.testing.assert_array_equal(helpers.bits(np.array([gx, gy])), helpers.bits(oracle.sobel_filter(grid, x, y)))
    for _ in range(200):
        p, a, b = (rng.uniform(-3, 3, 2).astype(np.float32) for _ in range(3))
        got = np.array(distance_from_line(p[0], p[1], a, b), np.float32)
        np.testing.assert_array_equal(helpers.bits(got), helpers.bits(oracle.distance_from_line(p, a, b)))


@pytest.mark.parametrize("use_distance_map", [True, False])
def test_step_agrees_with_the_cpp_oracle(use_distance_map):
    oracle.lib().oracle_set_threads(2)
    sc = helpers.corridor_scenario(size=(30.0, 16.0))
    field = helpers.oracle_field(sc)
    obs, _ = helpers.arrays_of(sc)
    pos, dest, vel, v0 = helpers.random_crowd(160, sc.field.size, seed=17, margin=-1.0, speed=False)
    pos[:6] = np.array(sc.waypoints[0].line[0], np.float32) + 0.1      # standing on their destination
    dest[:6] = 0
    py = PyModel(sc.field.size, 1.4, field.unit, field.distance_map, field.potential_maps, obstacles=obs,
                 use_distance_map=use_distance_map)
    cc = oracle.OracleModel(sc.field.size, 1.4, field.unit, field.distance_map, field.potential_maps, obstacles=obs,
                            use_distance_map=use_distance_map)
    py.spawn_pedestrians(pos, dest, v0)
    cc.spawn(pos, dest, v0)
    for tick in range(4):
        assert py.indices == cc.indices().tolist(), f"tick {tick}"
        pp, pd, pv, p0 = py.state()
        cp, cd, cv, c0 = cc.get()
        assert len(pd) == len(cd) and 100 < len(cd) < 160
        np.testing.assert_array_equal(pd, cd)
        np.testing.assert_array_equal(helpers.bits(p0), helpers.bits(c0))
        np.testing.assert_array_equal(np.isnan(pp), np.isnan(cp))
        np.testing.assert_allclose(pp, cp, rtol=0, atol=2e-6)
        np.testing.assert_allclose(pv, cv, rtol=0, atol=2e-5)
        # exp is the only operation not bit-identical by construction (correctly rounded here, glibc's 0.502-ulp
        # expf in the oracle): the overwhelming majority of the output floats must be identical in every bit
        differ = int((helpers.bits(pp) != helpers.bits(cp)).sum() + (helpers.bits(pv) != helpers.bits(cv)).sum())
        assert differ <= 0.03 * (tick + 1) * (pp.size + pv.size), (tick, differ, pp.size + pv.size)
        if tick == 0:  # nothing but the sort has run: bit for bit
            np.testing.assert_array_equal(helpers.bits(pp), helpers.bits(cp))
        py.update_states()
        cc.update()
        py.spawn_pedestrians()
        cc.spawn()
