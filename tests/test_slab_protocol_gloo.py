"""CPU, world_size 2 and 3 over gloo: the slab PROTOCOL of the multi-GPU path, executed with the CPU
oracle standing in for the device kernels.

What libpedoni_cuda does per rank and tick (include/pedoni_cuda.h "multi-GPU slabs"):
  step      integrate the owned rows [r0, r1) plus ONE ghost row each side, with neighbours taken from
            TWO ghost rows each side;
  rebuild   of the agents just integrated (and the replicated spawn list) keep those whose new cell
            row is owned, stably ordered by (cell, previous index);
  exchange  ship the first / last two owned rows to rank-1 / rank+1 (send/recv), no other messages.
The claim tested here, bit for bit and independent of any GPU: the rank-order concatenation of the
slabs equals the undecomposed model at every tick — i.e. two ghost rows are enough, redundant
integration of a ghost row reproduces its owner's result, migrants need no message of their own, and
the in-cell order (hence the force summation order) is preserved.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

UNIT = np.float32(1.4)


def _rows(pos):
    """neighbor_grid.rs:27 `(pos / unit).as_ivec2().y` in f32."""
    return np.trunc(pos[:, 1].astype(np.float32) / UNIT).astype(np.int64)


def _send(arrs, dst):
    n = torch.tensor([len(arrs[1])], dtype=torch.int64)
    dist.send(n, dst)
    if n.item():
        for a in arrs:
            dist.send(torch.from_numpy(np.ascontiguousarray(a)), dst)


def _recv(src):
    n = torch.zeros(1, dtype=torch.int64)
    dist.recv(n, src)
    n = int(n.item())
    out = [np.zeros((n, 2), np.float32), np.zeros(n, np.uint32), np.zeros((n, 2), np.float32), np.zeros(n, np.float32)]
    if n:
        for a in out:
            dist.recv(torch.from_numpy(a), src)
    return out


def _cat(parts):
    return [np.concatenate([p[k] for p in parts]) for k in range(4)]


def _exchange(rank, world, owned, r0, r1):
    """Ghost rows for this rank: (below = rows r0-2, r0-1 of rank-1, above = rows r1, r1+1 of rank+1)."""
    rows = _rows(owned[0])
    lo = [a[(rows >= r0) & (rows < r0 + 2)] for a in owned]
    hi = [a[(rows >= r1 - 2) & (rows < r1)] for a in owned]
    empty = [np.zeros((0, 2), np.float32), np.zeros(0, np.uint32), np.zeros((0, 2), np.float32), np.zeros(0, np.float32)]
    below, above = empty, empty
    # even ranks send first, odd ranks receive first: no deadlock with blocking gloo send/recv
    for phase in (0, 1):
        if rank % 2 == phase:
            if rank > 0:
                _send(lo, rank - 1)
            if rank < world - 1:
                _send(hi, rank + 1)
        else:
            if rank < world - 1:
                above = _recv(rank + 1)
            if rank > 0:
                below = _recv(rank - 1)
    return below, above


def _worker(rank, world, port, n_agents, ticks, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers
    import oracle
    from pedoni_b200 import slab_rows

    oracle.lib().oracle_set_threads(1)
    sc = helpers.corridor_scenario()
    field = helpers.oracle_field(sc)
    obs, _ = helpers.arrays_of(sc)
    new = lambda: oracle.OracleModel(sc.field.size, float(UNIT), field.unit, field.distance_map,  # noqa: E731
                                     field.potential_maps, obstacles=obs)
    ny, _nx = new().grid_shape()
    r0, r1 = slab_rows(ny, world, rank)
    pos, dest, vel, v0 = helpers.random_crowd(n_agents, sc.field.size, seed=33, margin=3.5)

    def rebuild(candidates):
        """Keep the candidates whose row is owned; oracle.spawn() does despawn + stable cell sort."""
        rows = _rows(candidates[0])
        keep = (rows >= r0) & (rows < r1)
        m = new()
        m.set(*[a[keep] for a in candidates])
        m.spawn()
        return list(m.get())

    owned = rebuild([pos, dest, vel, v0])  # upload_state: every rank sees the whole list, keeps its rows
    history = []
    rng = np.random.default_rng(5)
    for _tick in range(ticks):
        below, above = _exchange(rank, world, owned, r0, r1)
        history.append([a.copy() for a in owned])
        # step: integrate everything held; only rows [r0-1, r1+1) have complete neighbourhoods
        local = _cat([below, owned, above])
        m = new()
        m.set(*local)
        m.spawn()  # builds the cell table of the local set; it is already sorted and filtered -> no reorder
        assert m.count() == len(local[1])
        old_rows = _rows(m.get()[0])
        m.update()
        moved = list(m.get())
        valid = (old_rows >= r0 - 1) & (old_rows < r1 + 1)
        n = int(rng.poisson(15))  # replicated spawn list (same RNG stream on every rank)
        sp = np.stack([np.full(n, 6.0), rng.uniform(5, 25, n)], 1).astype(np.float32)
        spawn = [sp, np.ones(n, np.uint32), np.zeros((n, 2), np.float32),
                 np.clip(rng.normal(1.34, 0.26, n), 0.5, 2.2).astype(np.float32)]
        owned = rebuild(_cat([[a[valid] for a in moved], spawn]))
    history.append([a.copy() for a in owned])
    np.savez(Path(result_dir) / f"rank{rank}.npz", **{f"t{t}_{k}": h[k] for t, h in enumerate(history) for k in range(4)})
    dist.barrier()
    dist.destroy_process_group()


def _whole_domain(n_agents, ticks):
    import helpers
    import oracle
    sc = helpers.corridor_scenario()
    field = helpers.oracle_field(sc)
    obs, _ = helpers.arrays_of(sc)
    m = oracle.OracleModel(sc.field.size, float(UNIT), field.unit, field.distance_map, field.potential_maps,
                           obstacles=obs)
    pos, dest, vel, v0 = helpers.random_crowd(n_agents, sc.field.size, seed=33, margin=3.5)
    m.set(pos, dest, vel, v0)
    m.spawn()
    history = [list(m.get())]
    rng = np.random.default_rng(5)
    for _ in range(ticks):
        m.update()
        n = int(rng.poisson(15))
        sp = np.stack([np.full(n, 6.0), rng.uniform(5, 25, n)], 1).astype(np.float32)
        m.spawn(sp, np.ones(n, np.uint32), np.clip(rng.normal(1.34, 0.26, n), 0.5, 2.2).astype(np.float32))
        history.append(list(m.get()))
    return history


@pytest.mark.parametrize("world", [2, 3])
def test_slab_protocol_equals_whole_domain(world, tmp_path):
    n_agents, ticks = 1500, 12
    port = 29700 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, n_agents, ticks, str(tmp_path)), nprocs=world, join=True)
    whole = _whole_domain(n_agents, ticks)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    migrated = False
    for t, ref in enumerate(whole):
        for k in range(4):
            got = np.concatenate([p[f"t{t}_{k}"] for p in parts])
            assert got.shape == ref[k].shape, (t, k)
            np.testing.assert_array_equal(np.ascontiguousarray(got).view(np.uint32),
                                          np.ascontiguousarray(ref[k]).view(np.uint32), err_msg=f"tick {t} column {k}")
        migrated |= [len(p[f"t{t}_1"]) for p in parts] != [len(p["t0_1"]) for p in parts]
    assert migrated, "the scenario must move pedestrians across slab boundaries"


def test_slab_rows_partition_the_grid():
    from pedoni_b200 import slab_rows
    for ny in (2, 7, 22, 143, 2260):
        for count in (1, 2, 3, 4, 8):
            if ny < count:
                continue
            spans = [slab_rows(ny, count, r) for r in range(count)]
            assert spans[0][0] == 0 and spans[-1][1] == ny
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
