"""Slab decomposition on ONE GPU: `slab_count` handles in one process exchanging ghost rows through the
in-process transport must reproduce the whole-domain handle BIT FOR BIT (same kernels, same neighbor
order, same summation order) — counts, cell table, order, positions and velocities — and therefore
the oracle within the stated tolerance. The NCCL transport differs only in how the strips travel
(tests/test_gpu_nccl_slabs.py, needs >= 2 GPUs)."""
import numpy as np
import pytest

import helpers
from helpers import bits
from pedoni_b200 import (PEDONI_MATH_FAST, PEDONI_MATH_STRICT, PedoniError, SimulatorOptions, SlabGroup,
                         SocialForceModelCuda)
from pedoni_b200.synthetic import SyntheticCrowd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corridor():
    sc = helpers.corridor_scenario()
    return sc, helpers.oracle_field(sc)


def _assert_same(whole, slabs, what=""):
    assert whole.get_pedestrian_count() == slabs.get_pedestrian_count(), what
    wp, wd, wv, w0 = whole.download()
    sp, sd, sv, s0 = slabs.download()
    np.testing.assert_array_equal(wd, sd, err_msg=what)
    np.testing.assert_array_equal(bits(wp), bits(sp), err_msg=what)
    np.testing.assert_array_equal(bits(wv), bits(sv), err_msg=what)
    np.testing.assert_array_equal(bits(w0), bits(s0), err_msg=what)


@pytest.mark.parametrize("n_slabs", [2, 3, 4])
@pytest.mark.parametrize("mode", [PEDONI_MATH_STRICT, PEDONI_MATH_FAST])
def test_slabs_equal_whole_domain_bitwise(corridor, n_slabs, mode):
    sc, field = corridor  # 22 x 43 neighbor grid: 4 slabs of 5-6 rows
    opts = SimulatorOptions()
    whole = SocialForceModelCuda(opts, sc, field, math_mode=mode)
    slabs = SlabGroup(opts, sc, field, n_slabs, math_mode=mode)
    pos, dest, vel, v0 = helpers.random_crowd(3000, sc.field.size, seed=21, margin=3.5)
    for m in (whole, slabs):
        m.upload_state(pos, dest, vel, v0)
    rng = np.random.default_rng(5)
    for tick in range(25):
        n = int(rng.poisson(15))  # spawn stream crossing slab boundaries: x fixed, y over the whole height
        sp = np.stack([np.full(n, 6.0), rng.uniform(5, 25, n)], 1).astype(np.float32)
        sd = np.ones(n, np.uint32)
        s0 = np.clip(rng.normal(1.34, 0.26, n), 0.5, 2.2).astype(np.float32)
        for m in (whole, slabs):
            m.spawn_arrays(sp, sd, s0)
            m.rebuild()
        np.testing.assert_array_equal(whole.cell_table(), slabs.cell_table(), err_msg=f"tick {tick}")
        _assert_same(whole, slabs, f"after rebuild, tick {tick}")
        for m in (whole, slabs):
            m.step()
        _assert_same(whole, slabs, f"after step, tick {tick}")
    assert whole.counters()[1] == slabs.counters()[1]  # ghost rows are integrated twice but counted once
    whole.close()
    slabs.close()


def test_slabs_match_oracle(corridor):
    sc, field = corridor
    slabs = SlabGroup(SimulatorOptions(), sc, field, 3)
    _, orc = helpers.make_pair(sc, field)
    pos, dest, vel, v0 = helpers.random_crowd(2500, sc.field.size, seed=8, margin=4.0, speed=False)
    slabs.spawn_arrays(pos, dest, v0)
    slabs.rebuild()
    orc.spawn(pos, dest, v0)
    for _ in range(10):
        np.testing.assert_array_equal(slabs.cell_table(), orc.indices())
        slabs.step()
        orc.update()
        slabs.rebuild()
        orc.spawn()
    cp, cd, cv, _ = slabs.download()
    op, od, ov, _ = orc.get()
    np.testing.assert_array_equal(cd, od)
    assert np.abs(cp - op).max() <= helpers.TOL_POS_ABS and np.abs(cv - ov).max() <= helpers.TOL_VEL_ABS


def test_slabs_on_uniform_crowd_with_walkers_crossing_every_boundary():
    """200k-agent synthetic crowd (BASELINE.json configs[4] at reduced N), 8 slabs, 30 ticks."""
    crowd = SyntheticCrowd(n=200_000)
    sc, field = crowd.scenario(), crowd.field()
    pos, dest, vel, v0 = crowd.agents()
    # make every agent walk along y so that each tick moves pedestrians across slab boundaries
    vel[:, 1] = np.where(np.arange(len(vel)) % 2 == 0, 1.2, -1.2).astype(np.float32)
    whole = SocialForceModelCuda(SimulatorOptions(), sc, field, math_mode=PEDONI_MATH_FAST, capacity=210_000)
    slabs = SlabGroup(SimulatorOptions(), sc, field, 8, math_mode=PEDONI_MATH_FAST, capacity=30_000)
    for m in (whole, slabs):
        m.upload_state(pos, dest, vel, v0)
        m.rebuild()
    owned0 = [s.get_pedestrian_count() for s in slabs.slabs]
    for tick in range(30):
        for m in (whole, slabs):
            m.step()
            m.rebuild()
    np.testing.assert_array_equal(whole.cell_table(), slabs.cell_table())
    _assert_same(whole, slabs)
    owned1 = [s.get_pedestrian_count() for s in slabs.slabs]
    assert owned0 != owned1  # pedestrians did migrate between slabs
    whole.close()
    slabs.close()


def test_halo_overflow_and_row_jump_are_reported(corridor):
    sc, field = corridor
    pos, dest, vel, v0 = helpers.random_crowd(3000, sc.field.size, seed=3, margin=4.0)
    small = SlabGroup(SimulatorOptions(), sc, field, 2, halo_capacity=16)
    small.upload_state(pos, dest, vel, v0)
    small.rebuild()
    with pytest.raises(PedoniError) as e:
        small.get_pedestrian_count()
    assert "halo_capacity" in str(e.value)
    small.close()

    fast = SlabGroup(SimulatorOptions(), sc, field, 2)
    v0_fast = np.full_like(v0, 40.0)  # speed limit 1.3 * 40 m/s: a 30 m/s walker covers 3 m = 2 rows per step
    vel_fast = vel.copy()
    vel_fast[:, 1] = 30.0
    fast.upload_state(pos, dest, vel_fast, v0_fast)
    for _ in range(6):
        fast.rebuild()
        fast.step()
    with pytest.raises(PedoniError) as e:
        fast.get_pedestrian_count()
    assert "two or more neighbor-grid rows" in str(e.value)
    fast.close()


def test_step_without_ghosts_is_refused(corridor):
    sc, field = corridor
    lone = SocialForceModelCuda(SimulatorOptions(), sc, field, slab_rank=0, slab_count=2)
    lone.rebuild()
    with pytest.raises(PedoniError):
        lone.step()
    lone.close()
    with pytest.raises(PedoniError):  # 22 rows cannot give 12 slabs two rows each
        SocialForceModelCuda(SimulatorOptions(), sc, field, slab_rank=0, slab_count=12)


def test_slabs_on_a_growing_non_uniform_crowd():
    """bottleneck.toml (200 pedestrians/s walking in, everything funnels through one gap): the buffers of
    every slab grow from the default capacity while ghost strips are in flight, populations are very
    unequal between slabs, and the result must still equal the whole-domain handle bit for bit."""
    from pedoni_b200.simulator import Simulator
    sc = helpers.load_scenario("bottleneck")
    opts = SimulatorOptions()
    field = helpers.oracle_field(sc, opts.field_grid_unit)
    whole = Simulator(opts, sc, field, SocialForceModelCuda(opts, sc, field, math_mode=PEDONI_MATH_FAST), seed=4,
                      count_every=50)
    slabs = Simulator(opts, sc, field, SlabGroup(opts, sc, field, 4, math_mode=PEDONI_MATH_FAST, halo_capacity=8192),
                      seed=4, count_every=50)
    for t in range(350):
        mw, ms = whole.tick(), slabs.tick()
        if (t + 1) % 50 == 0:
            assert mw.active_ped_count == ms.active_ped_count > 0
    _assert_same(whole.model, slabs.model)
    # device-side observables: the slabs' reductions add up to the whole domain's (arrivals are counted by
    # the owner only, although ghost-row pedestrians are integrated twice)
    ow, os_ = whole.model.observe((0.0, 200.0), 16), slabs.model.observe((0.0, 200.0), 16)
    assert ow["count"] == os_["count"] and abs(ow["mean_speed"] - os_["mean_speed"]) < 1e-4
    np.testing.assert_array_equal(ow["per_destination"], os_["per_destination"])
    np.testing.assert_array_equal(ow["arrived"], os_["arrived"])
    np.testing.assert_array_equal(ow["bin_count"], os_["bin_count"])
    per_slab = [s.get_pedestrian_count() for s in slabs.model.slabs]
    assert max(per_slab) > 0 and sum(per_slab) == whole.model.get_pedestrian_count() > 6000
    whole.model.close()
    slabs.model.close()


def test_top_slab_gains_immigrants_with_tight_bounds_and_no_spawns():
    """Every slab is given ONLY the pedestrians of its own rows, with a capacity far below its population, so each
    rebuild goes through the path that refreshes the host's population bound from the device before growing the
    buffers (the bound is then exact). Everybody walks up: the top slab, whose arrays start at the halo capacity,
    adopts immigrants every tick without any spawn. Its sort scratch (perm) is indexed with that offset and must
    be sized with it (the overrun the round-1 advisor found); the result must equal the whole domain bit for bit."""
    from pedoni_b200 import slab_rows
    crowd = SyntheticCrowd(n=6000)
    sc, field = crowd.scenario(), crowd.field()
    pos, dest, vel, v0 = crowd.agents()
    vel[:, 1] = 1.2
    opts = SimulatorOptions()
    whole = SocialForceModelCuda(opts, sc, field, math_mode=PEDONI_MATH_FAST)
    slabs = SlabGroup(opts, sc, field, 2, math_mode=PEDONI_MATH_FAST, capacity=1024)
    ny, _ = whole.grid_shape()
    row = np.trunc(pos[:, 1] / np.float32(1.4)).astype(int)
    whole.upload_state(pos, dest, vel, v0)
    for r, s in enumerate(slabs.slabs):
        r0, r1 = slab_rows(ny, 2, r)
        mine = (row >= r0) & (row < r1)
        s.upload_state(pos[mine], dest[mine], vel[mine], v0[mine])
    for tick in range(20):
        for m in (whole, slabs):
            m.rebuild()
        np.testing.assert_array_equal(whole.cell_table(), slabs.cell_table(), err_msg=f"tick {tick}")
        _assert_same(whole, slabs, f"after rebuild, tick {tick}")
        for m in (whole, slabs):
            m.step()
    r0, _ = slab_rows(ny, 2, 1)
    came_from_below = np.isin(bits(slabs.slabs[1].download()[3]), bits(v0[row < r0]))  # desired speed as identity
    assert came_from_below.sum() > 20  # immigration into the top slab, adopted without any spawn
    whole.close()
    slabs.close()
