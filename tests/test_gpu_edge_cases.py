"""GPU: the corners of the path the reference's behaviour defines and the kernel handles on separate code
paths — dense jams (neighbour lists overflowing a round, tiles overflowing shared memory), cell-boundary
and negative coordinates (truncation toward zero, neighbor_grid.rs:27), coincident pedestrians (NaN ->
despawn), populations that are not a multiple of the warp / CTA size."""
import numpy as np
import pytest

import helpers
from helpers import bits
from pedoni_b200 import PEDONI_MATH_FAST, PEDONI_MATH_STRICT, SimulatorOptions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hall():
    sc = helpers.scenario_of((40.0, 40.0), waypoints=[(3, 3, 3, 37, 1.0), (37, 3, 37, 37, 1.0)])
    return sc, helpers.oracle_field(sc)


def _run_both(sc, field, pos, dest, vel, v0, ticks, mode):
    cu, orc = helpers.make_pair(sc, field, math_mode=mode)
    cu.upload_state(pos, dest, vel, v0)
    orc.set(pos, dest, vel, v0)
    for _ in range(ticks):
        cu.rebuild()
        orc.spawn()
        assert cu.get_pedestrian_count() == orc.count()
        np.testing.assert_array_equal(cu.cell_table(), orc.indices())
        cu.step()
        orc.update()
    out = cu.download(), orc.get()
    cu.close()
    return out


@pytest.mark.parametrize("mode", [PEDONI_MATH_STRICT, PEDONI_MATH_FAST])
@pytest.mark.parametrize("n,spread", [(1500, 6.0),    # ~40 per cell: every list needs several rounds
                                      (3000, 2.5)])   # ~500 per cell: windows overflow the warp tile -> global path
def test_dense_jam(hall, mode, n, spread):
    sc, field = hall
    rng = np.random.default_rng(4)
    pos = (np.array([20.0, 20.0]) + rng.uniform(-spread / 2, spread / 2, (n, 2))).astype(np.float32)
    dest = rng.integers(0, 2, n).astype(np.uint32)
    vel = rng.normal(0, 0.3, (n, 2)).astype(np.float32)
    v0 = np.full(n, 1.3, np.float32)
    (cp, cd, cv, _), (op, od, ov, _) = _run_both(sc, field, pos, dest, vel, v0, 2, mode)
    np.testing.assert_array_equal(cd, od)
    # forces are huge in a crush (hundreds of m/s^2 before the speed clamp): compare relative to the scale
    tol = 2e-5 if mode == PEDONI_MATH_STRICT else 2e-3
    assert np.nanmax(np.abs(cp - op)) <= tol * max(1.0, np.nanmax(np.abs(op)))
    assert np.nanmax(np.abs(cv - ov)) <= tol * max(1.0, np.nanmax(np.abs(ov))) * 50


def test_cell_boundaries_negative_coordinates_and_coincident_agents(hall):
    sc, field = hall
    u = np.float32(1.4)
    pos = np.array([[u * 5, u * 7], [np.nextafter(u * 5, 0, dtype=np.float32), u * 7],        # on / just below a boundary
                    [-0.5, 10.0], [10.0, -1.39], [-1.5, 10.0],                               # (-unit, 0) truncates to cell 0
                    [20.0, 20.0], [20.0, 20.0],                                              # coincident: 0/0 -> NaN -> despawn
                    [39.99, 39.99], [40.5, 20.0]], np.float32)                               # last cell / outside
    n = len(pos)
    dest, vel, v0 = np.ones(n, np.uint32), np.zeros((n, 2), np.float32), np.full(n, 1.3, np.float32)
    (cp, cd, cv, _), (op, od, ov, _) = _run_both(sc, field, pos, dest, vel, v0, 3, PEDONI_MATH_STRICT)
    assert len(od) < n  # some were dropped (outside, NaN), identically on both sides
    np.testing.assert_array_equal(np.isnan(cp), np.isnan(op))
    assert np.nanmax(np.abs(cp - op)) <= helpers.TOL_POS_ABS


@pytest.mark.parametrize("n", [1, 31, 32, 33, 127, 128, 129, 1000])
def test_population_sizes_around_warp_and_cta_boundaries(hall, n):
    sc, field = hall
    pos, dest, vel, v0 = helpers.random_crowd(n, sc.field.size, seed=n, margin=6.0)
    (cp, cd, cv, c0), (op, od, ov, o0) = _run_both(sc, field, pos, dest, vel, v0, 3, PEDONI_MATH_STRICT)
    np.testing.assert_array_equal(cd, od)
    np.testing.assert_array_equal(bits(c0), bits(o0))
    assert np.abs(cp - op).max() <= helpers.TOL_POS_ABS and np.abs(cv - ov).max() <= helpers.TOL_VEL_ABS


def test_capacity_growth_under_inflow(hall):
    """Buffers start at 1024 and grow; host upper bounds lag the device counts by design."""
    sc, field = hall
    cu, orc = helpers.make_pair(sc, field, capacity=1024)
    rng = np.random.default_rng(8)
    for tick in range(12):
        k = 700
        p = np.stack([rng.uniform(6, 34, k), rng.uniform(6, 34, k)], 1).astype(np.float32)
        d = rng.integers(0, 2, k).astype(np.uint32)
        s = np.full(k, 1.3, np.float32)
        cu.spawn_arrays(p, d, s)
        cu.rebuild()
        orc.spawn(p, d, s)
        cu.step()
        orc.update()
    assert cu.get_pedestrian_count() == orc.count() > 7000
    cp, cd, _, _ = cu.download()
    op, od, _, _ = orc.get()
    np.testing.assert_array_equal(cd, od)
    assert np.nanmax(np.abs(cp - op)) <= 1e-3  # 12 ticks at ~7 ped/m^2
    cu.close()


def test_cell_table_beyond_1024_scan_tiles():
    """17.8 M cells = 1 085 scan tiles: the tile prefix sums more predecessors than a block has threads, and
    tiles outnumber the resident CTAs several times over."""
    from pedoni_b200.synthetic import SyntheticCrowd
    crowd = SyntheticCrowd(n=120_000, density=120_000 / 5900.0 ** 2, field_unit=2.0)
    sc, field = crowd.scenario(), crowd.field()
    assert crowd.cells_per_side ** 2 > 1024 * 16384
    pos, dest, vel, v0 = crowd.agents()
    cu, orc = helpers.make_pair(sc, field, math_mode=PEDONI_MATH_FAST)
    cu.upload_state(pos, dest, vel, v0)
    orc.set(pos, dest, vel, v0)
    for _ in range(3):
        cu.rebuild()
        orc.spawn()
        assert cu.get_pedestrian_count() == orc.count()
        np.testing.assert_array_equal(cu.cell_table(), orc.indices())
        cu.step()
        orc.update()
    (cp, cd, _, _), (op, od, _, _) = cu.download(), orc.get()
    np.testing.assert_array_equal(cd, od)
    assert np.nanmax(np.abs(cp - op)) <= 4 * np.spacing(np.float32(5900.0))  # coordinates up to 5.9 km: 1 ulp = 4.9e-4
    cu.close()
