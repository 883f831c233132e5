"""Worker of tests/test_gpu_nccl_slabs.py: one process per GPU under torch.distributed.run.
Each rank drives ONE slab handle joined over NCCL (pedoni_comm_init); rank 0 also runs the
whole-domain handle and checks that the rank-order concatenation of the slabs equals it bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import pedoni_b200 as pb  # noqa: E402
from pedoni_b200.synthetic import SyntheticCrowd  # noqa: E402


def scenario_mode(rank, world, local, name, ticks):
    """A shipped scenario through the Simulator facade on every rank (same seed: the spawn lists are
    replicated, each slab keeps its rows): growing, very unevenly distributed crowd."""
    sys.path.insert(0, str(ROOT / "tests"))
    import helpers
    from pedoni_b200.simulator import Simulator
    sc = helpers.load_scenario(name)
    opts = pb.SimulatorOptions()
    field = pb.Field.from_scenario(sc, opts.field_grid_unit)
    slab = pb.SocialForceModelCuda(opts, sc, field, device=local, math_mode=pb.PEDONI_MATH_FAST, slab_rank=rank,
                                   slab_count=world, halo_capacity=8192)
    uid = [pb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    slab.comm_init(uid[0])
    sim = Simulator(opts, sc, field, slab, seed=4, count_every=10 ** 9, device_spawn=True)
    for _ in range(ticks):
        sim.tick()
    part = slab.download()
    parts = [None] * world
    dist.gather_object(part, parts if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        whole = Simulator(opts, sc, field, pb.SocialForceModelCuda(opts, sc, field, device=local,
                                                                   math_mode=pb.PEDONI_MATH_FAST),
                          seed=4, count_every=10 ** 9, device_spawn=True)
        for _ in range(ticks):
            whole.tick()
        want = whole.model.download()
        u32 = lambda a: np.ascontiguousarray(a).view(np.uint32)  # noqa: E731
        got = [np.concatenate([p[k] for p in parts]) for k in range(4)]
        ok = all(g.shape == w.shape and (u32(g) == u32(w)).all() for g, w in zip(got, want))
        print(f"NCCL-SLABS {'OK' if ok else 'MISMATCH'} scenario={name} world={world} n={len(want[1])} "
              f"transport={slab.slab_transport()!r} owned={[len(p[1]) for p in parts]}", flush=True)
    slab.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # plumbing only: the data path is the library's own NCCL communicator
    if sys.argv[1].startswith("scenario:"):
        return scenario_mode(rank, world, local, sys.argv[1].split(":", 1)[1], int(sys.argv[2]))
    n_agents, ticks = int(sys.argv[1]), int(sys.argv[2])
    crowd = SyntheticCrowd(n=n_agents)
    sc, field = crowd.scenario(), crowd.field()
    pos, dest, vel, v0 = crowd.agents()
    vel[:, 1] = np.where(np.arange(len(vel)) % 2 == 0, 1.2, -1.2).astype(np.float32)
    opts = pb.SimulatorOptions()
    slab = pb.SocialForceModelCuda(opts, sc, field, device=local, math_mode=pb.PEDONI_MATH_FAST,
                                   capacity=int(1.3 * n_agents / world), slab_rank=rank, slab_count=world)
    uid = [pb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    slab.comm_init(uid[0])
    transport = slab.slab_transport()
    slab.upload_state(pos, dest, vel, v0)
    dist.barrier()
    slab.rebuild()
    n0 = slab.get_pedestrian_count()
    for _ in range(ticks):
        slab.step()
        slab.rebuild()
    part = slab.download()
    table = slab.cell_table()
    parts = [None] * world
    dist.gather_object((part, table, n0), parts if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        whole = pb.SocialForceModelCuda(opts, sc, field, device=local, math_mode=pb.PEDONI_MATH_FAST,
                                        capacity=int(1.1 * n_agents))
        whole.upload_state(pos, dest, vel, v0)
        whole.rebuild()
        for _ in range(ticks):
            whole.step()
            whole.rebuild()
        wp, wd, wv, w0 = whole.download()
        cat = [np.concatenate([p[0][k] for p in parts]) for k in range(4)]
        u32 = lambda a: np.ascontiguousarray(a).view(np.uint32)  # noqa: E731
        ok = (cat[0].shape == wp.shape and (u32(cat[0]) == u32(wp)).all() and (cat[1] == wd).all()
              and (u32(cat[2]) == u32(wv)).all() and (u32(cat[3]) == u32(w0)).all())
        stitched, base = [np.zeros(1, np.uint64)], 0
        for p in parts:
            t = p[1].astype(np.uint64)
            stitched.append(t[1:] + base)
            base += int(t[-1])
        ok = ok and (np.concatenate(stitched) == whole.cell_table()).all()
        migrated = [p[2] for p in parts] != [len(p[0][1]) for p in parts]
        print(f"NCCL-SLABS {'OK' if ok and migrated else 'MISMATCH'} world={world} n={len(wd)} transport={transport!r} "
              f"owned_before={[p[2] for p in parts]} owned_after={[len(p[0][1]) for p in parts]}", flush=True)
        ok = ok and migrated
    slab.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
