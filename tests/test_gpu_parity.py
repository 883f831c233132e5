"""GPU parity tests proper: the CUDA path through the C ABI vs the CPU oracle on the same seeded
inputs. Bit-exact for cell keys / cell table / membership / order; stated fp32 tolerance for
positions and velocities (helpers.TOL_*)."""
import numpy as np
import pytest

import helpers
from helpers import TOL_POS_ABS, TOL_VEL_ABS, bits
from pedoni_b200 import PEDONI_MATH_FAST, PEDONI_MATH_STRICT, SimulatorOptions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corridor():
    sc = helpers.corridor_scenario()
    return sc, helpers.oracle_field(sc)


def _spawn_both(cu, orc, pos, dest, v0):
    cu.spawn_arrays(pos, dest, v0)
    cu.rebuild()
    orc.spawn(pos, dest, v0)


def _assert_rebuild_bit_exact(cu, orc):
    assert cu.get_pedestrian_count() == orc.count()
    np.testing.assert_array_equal(cu.cell_table(), orc.indices())
    cp, cd, cv, c0 = cu.download()
    op, od, ov, o0 = orc.get()
    np.testing.assert_array_equal(bits(cp), bits(op))  # same agents, same (stable) order
    np.testing.assert_array_equal(cd, od)
    np.testing.assert_array_equal(bits(cv), bits(ov))
    np.testing.assert_array_equal(bits(c0), bits(o0))


def test_rebuild_bit_exact_with_drops_and_despawns(corridor):
    sc, field = corridor
    cu, orc = helpers.make_pair(sc, field)
    pos, dest, vel, v0 = helpers.random_crowd(5000, sc.field.size, seed=1, margin=-3.0)  # some outside the grid
    pos[:50] = np.array(sc.waypoints[0].line[0], np.float32) + np.random.default_rng(2).normal(0, 0.3, (50, 2)).astype(np.float32)
    dest[:50] = 0  # standing on their destination: potential <= 0.25 -> despawn (sfm.rs:69)
    pos[50] = (np.nan, 1.0)  # NaN -> cell (0, 0), predicate false -> dropped
    _spawn_both(cu, orc, pos, dest, v0)
    assert orc.count() < 5000
    _assert_rebuild_bit_exact(cu, orc)


@pytest.mark.parametrize("mode", [PEDONI_MATH_STRICT, PEDONI_MATH_FAST])
def test_ten_steps_within_tolerance(corridor, mode):
    sc, field = corridor
    cu, orc = helpers.make_pair(sc, field, math_mode=mode)
    pos, dest, vel, v0 = helpers.random_crowd(3000, sc.field.size, seed=3, margin=4.0)
    cu.upload_state(pos, dest, vel, v0)
    cu.rebuild()
    orc.spawn(pos, dest, v0)
    # give the oracle the same non-zero velocities (spawn sets them to zero): re-sort is identity afterwards
    op, od, _, o0 = orc.get()
    # map velocities through the same permutation: positions are unique, match on bits
    order = {tuple(b): i for i, b in enumerate(bits(pos).tolist())}
    perm = np.array([order[tuple(b)] for b in bits(op).tolist()])
    orc.set(op, od, vel[perm], o0)
    worst_p = worst_v = 0.0
    for step in range(10):
        cu.rebuild() if step else None
        orc.spawn()
        _assert_counts_and_table(cu, orc)
        cu.step()
        orc.update()
        cp, cd, cv, _ = cu.download()
        op, od, ov, _ = orc.get()
        assert cp.shape == op.shape
        worst_p = max(worst_p, float(np.nanmax(np.abs(cp - op))))
        worst_v = max(worst_v, float(np.nanmax(np.abs(cv - ov))))
        np.testing.assert_array_equal(np.isnan(cp), np.isnan(op))
    print(f"mode={mode} worst |dpos|={worst_p:.3e} worst |dvel|={worst_v:.3e}")
    tol_p, tol_v = helpers.tolerances(mode)
    assert worst_p <= tol_p and worst_v <= tol_v


def _assert_counts_and_table(cu, orc):
    assert cu.get_pedestrian_count() == orc.count()
    np.testing.assert_array_equal(cu.cell_table(), orc.indices())


def test_strict_single_step_is_bit_exact_or_1ulp(corridor):
    sc, field = corridor
    cu, orc = helpers.make_pair(sc, field, math_mode=PEDONI_MATH_STRICT)
    pos, dest, vel, v0 = helpers.random_crowd(4000, sc.field.size, seed=5, margin=4.0, speed=False)
    _spawn_both(cu, orc, pos, dest, v0)
    cu.step()
    orc.update()
    cp, _, cv, _ = cu.download()
    op, _, ov, _ = orc.get()
    mism = int((bits(cp) != bits(op)).sum() + (bits(cv) != bits(ov)).sum())
    print(f"strict: {mism} of {cp.size + cv.size} floats differ in bits; max |dpos| = {np.abs(cp - op).max():.3e}")
    # exp is the only op not bit-identical by construction: the device rounds a correctly rounded fp64 exp
    # once, glibc's expf (what Rust's f32::exp calls; FMA or non-FMA variant picked per CPU) is a 0.502-ulp
    # routine. About 1 exp in 10^3 differs in the last bit, ~12 exps feed each velocity: allow 0.5 %
    # of the outputs to differ, and only in the last bits.
    assert mism <= 0.005 * (cp.size + cv.size)
    assert np.abs(cp - op).max() <= 1e-6 and np.abs(cv - ov).max() <= 1e-5


def test_segment_walls_variant(corridor):
    sc, field = corridor
    opts = SimulatorOptions(use_distance_map=False)
    cu, orc = helpers.make_pair(sc, field, options=opts)
    pos, dest, vel, v0 = helpers.random_crowd(2000, sc.field.size, seed=7, margin=4.0, speed=False)
    _spawn_both(cu, orc, pos, dest, v0)
    for _ in range(5):
        cu.step()
        orc.update()
        cu.rebuild()
        orc.spawn()
        assert cu.get_pedestrian_count() == orc.count()
    cp, _, cv, _ = cu.download()
    op, _, ov, _ = orc.get()
    assert np.abs(cp - op).max() <= TOL_POS_ABS and np.abs(cv - ov).max() <= TOL_VEL_ABS


def test_spawn_stream_over_ticks(corridor):
    """lib.rs:64-100 tick order: spawn_pedestrians(new) then update_states, new agents every tick."""
    sc, field = corridor
    cu, orc = helpers.make_pair(sc, field)
    rng = np.random.default_rng(11)
    for tick in range(30):
        n = int(rng.poisson(20))
        pos = np.stack([np.full(n, 6.0), rng.uniform(6, 24, n)], 1).astype(np.float32)
        dest = np.ones(n, np.uint32)
        v0 = np.clip(rng.normal(1.34, 0.26, n), 0.5, 2.2).astype(np.float32)
        cu.spawn_arrays(pos, dest, v0)
        cu.rebuild()
        orc.spawn(pos, dest, v0)
        _assert_counts_and_table(cu, orc)
        cu.step()
        orc.update()
    cp, cd, cv, _ = cu.download()
    op, od, ov, _ = orc.get()
    np.testing.assert_array_equal(cd, od)
    assert np.abs(cp - op).max() <= TOL_POS_ABS


def test_empty_model_and_errors(corridor):
    sc, field = corridor
    cu, orc = helpers.make_pair(sc, field)
    cu.rebuild()
    cu.step()
    assert cu.get_pedestrian_count() == 0
    assert (cu.cell_table() == 0).all()
    from pedoni_b200 import PedoniError
    pos, dest, vel, v0 = helpers.random_crowd(10, sc.field.size, seed=1)
    cu.spawn_arrays(pos, dest, v0)
    with pytest.raises(PedoniError):  # step with un-rebuilt spawns
        cu.step()
    with pytest.raises(PedoniError):  # O(N^2) path is not built
        helpers.make_pair(sc, field, options=SimulatorOptions(use_neighbor_grid=False))


@pytest.mark.parametrize("pack", ["1", "0", "u8"])
def test_pipelined_download_equals_blocking_download(corridor, monkeypatch, pack):
    """pedoni_download_begin/_end: the snapshot is of the moment of the call, whatever is enqueued after it.
    With at most 256 potential maps the destinations cross PCIe as bytes and are widened on the host — or are
    delivered as bytes (pedoni_download_begin_u8)."""
    import torch
    monkeypatch.setenv("PEDONI_DOWNLOAD_PACK", "0" if pack == "u8" else pack)
    sc, field = corridor
    cu, _ = helpers.make_pair(sc, field)
    assert cu.download_wire_bytes() == (9 if pack == "1" else 12)
    pos, dest, vel, v0 = helpers.random_crowd(3001, sc.field.size, seed=9, margin=4.0)
    cu.upload_state(pos, dest, vel, v0)
    cu.rebuild()
    h_pos = torch.empty((4000, 2), dtype=torch.float32).pin_memory().numpy()
    h_dest = torch.empty(4000, dtype=torch.uint8).pin_memory().numpy() if pack == "u8" else \
        torch.empty(4000, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    for _ in range(5):
        cu.step()
        want_pos, want_dest = cu.download(vel=False, v0=False)[:2]
        cu.download_begin(h_pos, h_dest)
        cu.rebuild()          # the model moves on while the copy is in flight
        cu.step()
        cu.rebuild()
        got_pos, got_dest = cu.download_end()
        np.testing.assert_array_equal(bits(got_pos), bits(want_pos))
        np.testing.assert_array_equal(got_dest, want_dest)
    from pedoni_b200 import PedoniError
    with pytest.raises(PedoniError):
        cu.download_end()     # nothing in flight
    cu.close()


def test_published_count_trails_but_never_blocks(corridor):
    sc, field = corridor
    cu, orc = helpers.make_pair(sc, field)
    pos, dest, vel, v0 = helpers.random_crowd(2000, sc.field.size, seed=2, margin=4.0)
    cu.upload_state(pos, dest, vel, v0)
    seen = []
    for _ in range(5):
        cu.rebuild()
        cu.step()
        seen.append(cu.count_published())  # no synchronisation: may still show an older rebuild
    cu.synchronize()
    n, ordinal = cu.count_published()
    cu.rebuild()                           # count() reports the state after this rebuild
    assert n + 0 >= cu.get_pedestrian_count() > 0
    assert all(a[1] <= b[1] for a, b in zip(seen, seen[1:])) and seen[-1][1] <= ordinal
    cu.close()


def test_texture_gather_path_is_bit_identical_to_loads(corridor, monkeypatch):
    """Fast math fetches the 4x4 field footprints with texture gathers when the handle could build the 2D
    arrays (pedoni_field_textures); the texel values are the same, so ten ticks must agree bit for bit with
    a handle forced onto plain loads."""
    sc, field = corridor
    from pedoni_b200 import SocialForceModelCuda
    pos, dest, vel, v0 = helpers.random_crowd(4000, sc.field.size, seed=11, margin=2.0)
    outs = []
    for knob in ("1", "0"):
        monkeypatch.setenv("PEDONI_FIELD_TEXTURES", knob)
        cu = SocialForceModelCuda(SimulatorOptions(), sc, field, math_mode=PEDONI_MATH_FAST)
        assert cu.field_textures() == (knob == "1")
        cu.upload_state(pos, dest, vel, v0)
        for _ in range(10):
            cu.rebuild()
            cu.step()
        outs.append(cu.download())
        cu.close()
    for a, b in zip(*outs):
        np.testing.assert_array_equal(bits(a), bits(b))
    monkeypatch.setenv("PEDONI_FIELD_TEXTURES", "1")
    strict = SocialForceModelCuda(SimulatorOptions(), sc, field, math_mode=PEDONI_MATH_STRICT)
    assert not strict.field_textures()
    strict.close()


def test_wall_far_mask_changes_nothing(monkeypatch):
    """Fast math skips the wall term (sfm.rs:188-192) in cells that pedoni_wall_far_cells marks: more than 8 m from
    every obstacle (the term is below 1e-17 m/s^2 there) and off the ridges of the distance map (where the reference's
    normalize() may yield NaN). Twenty ticks of an open 144 m square agree with a handle that evaluates every wall
    term (PEDONI_WALL_CUTOFF=0) bit for bit (up to two last-place flips allowed: 1e-17 can decide a rounding)."""
    from pedoni_b200 import SocialForceModelCuda
    from pedoni_b200.synthetic import SyntheticCrowd
    crowd = SyntheticCrowd(20000)
    sc, field = crowd.scenario(), crowd.field()
    pos, dest, vel, v0 = crowd.agents()
    outs, marked = [], []
    for knob in ("1", "0"):
        monkeypatch.setenv("PEDONI_WALL_CUTOFF", knob)
        cu = SocialForceModelCuda(SimulatorOptions(), sc, field, math_mode=PEDONI_MATH_FAST)
        marked.append(cu.wall_far_cells())
        cu.upload_state(pos, dest, vel, v0)
        for _ in range(20):
            cu.rebuild()
            cu.step()
        outs.append(cu.download())
        cu.close()
    (far, total), (far_off, _) = marked
    # 144 m square, 8 m cut-off + footprint margin, minus the four diagonal ridges of min(x, y, W - x, H - y)
    assert far_off == 0 and 0.45 * total < far < 0.8 * total, marked
    for a, b in zip(*outs):
        assert a.shape == b.shape
        differ = bits(a) != bits(b)
        assert differ.sum() <= 2 and (differ.sum() == 0 or np.abs(a - b).max() <= 1e-6), int(differ.sum())
    monkeypatch.setenv("PEDONI_WALL_CUTOFF", "1")
    # a 16 m wide corridor (walls 10 m apart) has no far blocks at all, and strict handles never carry a mask
    sc2 = helpers.corridor_scenario(size=(60.0, 16.0))
    f2 = helpers.oracle_field(sc2)
    for mode in (PEDONI_MATH_FAST, PEDONI_MATH_STRICT):
        cu = SocialForceModelCuda(SimulatorOptions(), sc2, f2, math_mode=mode)
        assert cu.wall_far_cells()[0] == 0
        cu.close()
    strict = SocialForceModelCuda(SimulatorOptions(), sc, field, math_mode=PEDONI_MATH_STRICT)
    assert strict.wall_far_cells()[0] == 0
    strict.close()


def test_two_pipelined_downloads_in_flight(corridor):
    """begin k, begin k+1, end k, end k+1: _end completes the oldest; a third begin is refused."""
    import torch
    from pedoni_b200 import PedoniError
    sc, field = corridor
    cu, _ = helpers.make_pair(sc, field)
    pos, dest, vel, v0 = helpers.random_crowd(2500, sc.field.size, seed=12, margin=4.0)
    cu.upload_state(pos, dest, vel, v0)
    cu.rebuild()
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    bufs = [(pin((3000, 2), torch.float32), pin(3000, torch.int32).view(np.uint32)) for _ in range(3)]
    want = []
    for k in range(2):
        cu.step()
        cu.rebuild()
        want.append(cu.download(vel=False, v0=False)[:2])
        cu.download_begin(*bufs[k])
    with pytest.raises(PedoniError):
        cu.download_begin(*bufs[2])
    for k in range(2):
        got_pos, got_dest = cu.download_end()
        np.testing.assert_array_equal(bits(got_pos), bits(want[k][0]))
        np.testing.assert_array_equal(got_dest, want[k][1])
    cu.close()


def test_profile_timeline_lists_every_launch_of_a_tick(corridor):
    """pedoni_profile_timeline: with profiling on, every launch between timer_begin and the read appears once, with
    its stream and offsets (the two-stream timeline bench.py --timeline writes for a slab handle)."""
    sc, field = corridor
    cu, _ = helpers.make_pair(sc, field, math_mode=PEDONI_MATH_FAST)
    pos, dest, vel, v0 = helpers.random_crowd(3000, sc.field.size, seed=4, margin=4.0)
    cu.upload_state(pos, dest, vel, v0)
    cu.rebuild()
    cu.profile_enable(True)
    cu.profile_reset()
    cu.timer_begin()
    for _ in range(3):
        cu.step()
        cu.rebuild()
    ms = cu.timer_end()
    tl = cu.profile_timeline()
    cu.profile_enable(False)
    assert [k for k, _, _, _ in tl] == ["force", "sort"] * 3 and all(s == "main" for _, s, _, _ in tl)
    starts = [a for _, _, a, _ in tl]
    assert starts == sorted(starts) and all(0 <= a <= b <= ms + 1e-3 for _, _, a, b in tl)
    times = cu.profile_read()
    assert times["force_launches"] == 3 and times["gather_launches"] == 3 and times["force_edge_launches"] == 0
    cu.close()


def test_library_pinned_buffers_carry_a_download(corridor):
    """pedoni_host_alloc / pedoni_host_free: page-locked staging for a host that does not link the CUDA runtime (the
    Rust shim's list_pedestrians)."""
    import ctypes as C
    from pedoni_b200 import _capi
    sc, field = corridor
    cu, _ = helpers.make_pair(sc, field)
    pos, dest, vel, v0 = helpers.random_crowd(1500, sc.field.size, seed=6, margin=4.0)
    cu.upload_state(pos, dest, vel, v0)
    cu.rebuild()
    lib = _capi.load()
    n = cu.get_pedestrian_count()
    p_pos, p_dest = lib.pedoni_host_alloc(8 * n), lib.pedoni_host_alloc(4 * n)
    assert p_pos and p_dest
    h_pos = np.ctypeslib.as_array(C.cast(p_pos, C.POINTER(C.c_float)), shape=(n, 2))
    h_dest = np.ctypeslib.as_array(C.cast(p_dest, C.POINTER(C.c_uint32)), shape=(n,))
    want_pos, want_dest = cu.download(vel=False, v0=False)[:2]
    got_pos, got_dest = cu.download(vel=False, v0=False, out=(h_pos, h_dest, None, None))[:2]
    np.testing.assert_array_equal(bits(got_pos), bits(want_pos))
    np.testing.assert_array_equal(got_dest, want_dest)
    lib.pedoni_host_free(p_pos)
    lib.pedoni_host_free(p_dest)
    cu.close()
