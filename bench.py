#!/usr/bin/env python
"""bench.py — pedestrian-updates/sec of the per-timestep pedestrian update on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--agents 10000000] [--density 1.0] [--math fast|strict]

One "step" = one tick of the hot path over the whole crowd: neighbor-grid rebuild (pedoni_rebuild)
+ forces and integration (pedoni_step). Workload = BASELINE.json configs[4]: the synthetic uniform
crowd of 10 M pedestrians on a large open domain (pedoni_b200/synthetic.py); it fits one GPU, so it is
also the N = 1 workload. With --gpus N the SAME 10 M crowd is slab-decomposed over N ranks
(strong scaling, as the north star states the target on a fixed 10 M crowd).

Prints ONE JSON line (rank 0). `value` = updates/s with state resident in HBM; `e2e` = the same tick
driven through the C ABI with HOST buffers: per step a spawn batch is copied host->device and the
trait's list_pedestrians payload (pos + destination) is copied device->host, inside the timed region.
`roofline` is for the dominant kernel (force+integrate) from per-launch CUDA events; `cpu_baseline`
is the C++ restatement of the Rust reference (oracle/) on this box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ALGO_BYTES_PER_UPDATE = 48  # read 24 B state + write 24 B state
# DRAM bytes the force kernel actually moves per update: dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` launch at this workload (profiles/r01i_force_10M_full.md: 2.094 GB + 0.343 GB for
# 9 999 438 pedestrians). The excess over 48 B is the field maps: two 4x4 texel footprints per pedestrian.
NCU_TRAFFIC_BYTES_PER_UPDATE = (2.093599e9 + 343.457536e6) / 9999438
RELAX_STEPS = 50            # untimed: lets the zero-velocity seed crowd reach walking state
E2E_SPAWN_PER_STEP = 1024   # host->device spawn batch per e2e step


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=10_000_000)
    ap.add_argument("--density", type=float, default=1.0)
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--cpu-agents", type=int, default=4_000_000, help="bounded sample for the CPU legs (~0.35 s/tick on 16 cores)")
    ap.add_argument("--cpu-steps", type=int, default=20, help="ticks of the cpu_baseline leg (~10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--relax", type=int, default=RELAX_STEPS)
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.samples, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def window(self, t0, t1):
        rows = [s.split(", ") for t, s in self.samples if t0 - 0.05 <= t <= t1 + 0.15 and s]
        if not rows:
            rows = [s.split(", ") for _, s in self.samples[-3:] if s]
        sm = sorted(int(r[0]) for r in rows if r[0].isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4),
                          ("sw_power_cap", 5)):
            if any(len(r) > col and r[col].strip() == "Active" for r in rows):
                reasons.append(name)
        smax = max((int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": reasons,
                "samples": len(rows)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path, i.e. (no Rust toolchain in
    this image) its C++ restatement in oracle/, on all host threads, bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle
    from pedoni_b200.synthetic import SyntheticCrowd

    crowd = SyntheticCrowd(n=args.cpu_agents, density=args.density)
    field = crowd.field()
    sc = crowd.scenario()
    pos, dest, vel, v0 = crowd.agents()
    m = oracle.OracleModel(sc.field.size, 1.4, field.unit, field.distance_map, field.potential_maps)
    m.spawn(pos, dest, v0)
    oracle.lib().oracle_set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: use every host core
    cores = oracle.lib().oracle_max_threads()
    m.run(max(args.warmup, 1))
    t0 = time.time()
    updates, ts, tc = m.run(args.steps)
    wall = time.time() - t0
    value = updates / (ts + tc)
    sample = f"{args.cpu_agents} agents of the same synthetic crowd (density {args.density}/m^2), {args.steps} ticks"
    print(json.dumps({
        "impl": "reference", "metric": "pedestrian-updates/sec", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (ts + tc) / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as the CUDA arm; each timed step is a tick of a bounded sample of it
        "config": {"workload": f"synthetic uniform crowd, {args.agents} pedestrians, open domain "
                               f"{SyntheticCrowd(n=args.agents, density=args.density).side:.0f} m x "
                               f"{SyntheticCrowd(n=args.agents, density=args.density).side:.0f} m (BASELINE.json configs[4])",
                   "density_per_m2": args.density, "neighbor_unit_m": 1.4, "field_unit_m": 0.25,
                   "sample_agents": args.cpu_agents,
                   "implementation": "C++ restatement of the Rust reference (oracle/): the Rust toolchain is absent"},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "time_spawn_s": ts, "time_calc_state_s": tc, "wall_s": wall},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(args):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle
    from pedoni_b200.synthetic import SyntheticCrowd

    crowd = SyntheticCrowd(n=args.cpu_agents, density=args.density)
    field = crowd.field()
    sc = crowd.scenario()
    pos, dest, vel, v0 = crowd.agents()
    m = oracle.OracleModel(sc.field.size, 1.4, field.unit, field.distance_map, field.potential_maps)
    m.spawn(pos, dest, v0)
    oracle.lib().oracle_set_threads(os.cpu_count() or 1)
    m.run(1)
    updates, ts, tc = m.run(args.cpu_steps)
    return {"value": updates / (ts + tc), "unit": "updates/s", "cores": oracle.lib().oracle_max_threads(),
            "kind": "port",
            "sample": f"{args.cpu_agents} agents of the same synthetic crowd, {args.cpu_steps} ticks; "
                      "C++ restatement of the Rust reference (rayon force loop -> OpenMP; rebuild and "
                      "integration serial as in sfm.rs)",
            "time_spawn_s": ts, "time_calc_state_s": tc}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import pedoni_b200 as pb
    from pedoni_b200.synthetic import SyntheticCrowd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    crowd = SyntheticCrowd(n=args.agents, density=args.density)
    sc = crowd.scenario()
    field = crowd.field()
    opts = pb.SimulatorOptions()
    math_mode = pb.PEDONI_MATH_FAST if args.math == "fast" else pb.PEDONI_MATH_STRICT
    model = pb.SocialForceModelCuda(opts, sc, field, device=local_rank, math_mode=math_mode,
                                    capacity=int(args.agents * 1.05 / world) + 65536,
                                    slab_rank=rank, slab_count=world)
    if world > 1:
        uid = [pb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        model.comm_init(uid[0])
    del field

    # Every rank generates the crowd in chunks and keeps the agents whose rows it owns (the library
    # drops foreign rows of replicated spawn lists).
    chunk = 2_000_000
    for lo in range(0, args.agents, chunk):
        pos, dest, vel, v0 = crowd.agents(lo, min(lo + chunk, args.agents))
        model.spawn_arrays(pos, dest, v0)
    barrier()  # ranks generate the crowd at different speeds; the first ghost exchange should find everybody there
    model.rebuild()
    n0 = model.get_pedestrian_count()
    for _ in range(args.relax):
        model.step()
        model.rebuild()
    model.synchronize()

    def tick():
        model.step()
        model.rebuild()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: state resident in HBM ------------------------------------------------------------
    for _ in range(args.warmup):
        tick()
    model.profile_enable(True)
    model.profile_reset()
    l0, u0 = model.counters()
    barrier()
    t_wall0 = time.time()
    model.timer_begin()
    for _ in range(args.steps):
        tick()
    ms = model.timer_end()
    barrier()
    t_wall1 = time.time()
    l1, u1 = model.counters()
    prof = model.profile_read()
    model.profile_enable(False)
    updates = u1 - u0

    t = torch.tensor([ms, float(updates), float(l1 - l0)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_max, updates_all, launches_all = tmax[0].item(), tsum[1].item(), tsum[2].item()
    else:
        ms_max, updates_all, launches_all = ms, float(updates), float(l1 - l0)
    value = updates_all / (ms_max * 1e-3)

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ---------------------
    e2e = None
    if not args.no_e2e:
        cap = int(n0 * 1.02) + 8 * E2E_SPAWN_PER_STEP * (args.steps + args.warmup)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
        h_pos, h_dest = pin((cap, 2), torch.float32), pin((cap,), torch.int32).view(np.uint32)
        nb = E2E_SPAWN_PER_STEP
        s_pos, s_dest, s_v0 = pin((nb, 2), torch.float32), pin((nb,), torch.int32).view(np.uint32), \
            pin((nb,), torch.float32)
        extra = SyntheticCrowd(n=args.agents, density=args.density, seed=crowd.seed ^ 0xE2E)
        h_pos2, h_dest2 = pin((cap, 2), torch.float32), pin((cap,), torch.int32).view(np.uint32)
        bufs = [(h_pos, h_dest), (h_pos2, h_dest2)]
        state = {"inflight": 0, "n": 0, "bytes": 0}

        def collect(keep=0):
            """Finish the pipelined list_pedestrians of earlier ticks (all but the newest `keep`): their payload
            is on the host, in the API's types, when this returns."""
            while state["inflight"] > keep:
                pos, dest = model.download_end()
                state["n"] += pos.shape[0]
                state["bytes"] = pos.shape[0] * model.download_wire_bytes()  # what crossed PCIe for this tick
                state["inflight"] -= 1

        # the synthetic inflow (uniform over the domain) is drawn before the clock starts: generating random
        # numbers in numpy is not part of the path, copying them into the pinned staging buffers is
        n_ticks = max(args.warmup, 1) + args.steps + 2
        inflow = [extra.agents(k * nb, (k + 1) * nb) for k in range(n_ticks)]

        def compute(k):
            p, d, _, v = inflow[k]
            s_pos[:], s_dest[:], s_v0[:] = p, d, v
            model.spawn_arrays(s_pos, s_dest, s_v0)   # H2D (spawn_pedestrians, first half)
            model.rebuild()                           # spawn_pedestrians, second half
            model.step()                              # update_states

        # One GPU: the next tick is enqueued BEFORE the previous tick's download is finished on the host, so
        # the GPU computes while the host waits for the copy and widens the destinations (PCIe stays
        # saturated: 1.88 -> 1.61 ms per tick). One process per slab: the plain order is faster (measured at
        # 2 GPUs: 1.31 vs 2.31 ms per tick — the ranks' rebuilds exchange ghost rows and want to stay in
        # lockstep; a rank sitting in a long host-side wait right after enqueueing delays its neighbour).
        eager = world == 1

        def e2e_tick(k):
            """Steady state of the pipeline: tick k has been enqueued. Start its list_pedestrians (device
            snapshot + async D2H), enqueue tick k+1 and finish tick k-1 on the host (wait for its copy, widen
            its destinations) — in the order `eager` says."""
            model.download_begin(*bufs[k % 2])
            state["inflight"] += 1
            if eager:
                compute(k + 1)
                collect(keep=1)
            else:
                collect(keep=1)
                compute(k + 1)

        compute(0)
        for k in range(max(args.warmup, 1)):
            e2e_tick(k)
        collect()
        state["n"] = 0
        model.synchronize()
        barrier()
        t0 = time.perf_counter()
        k0 = max(args.warmup, 1)
        for k in range(k0, k0 + args.steps):          # K x (one tick computed, one tick's pedestrians delivered)
            e2e_tick(k)
        collect()                                     # every timed tick's result has been read on the host
        model.synchronize()
        wall_e2e = (time.perf_counter() - t0) * 1e3
        barrier()
        n_e2e, d2h = state["n"], state["bytes"]
        te = torch.tensor([wall_e2e, float(n_e2e), float(d2h)], dtype=torch.float64, device="cuda")
        if world > 1:
            tm = te.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            tsu = te.clone()
            dist.all_reduce(tsu, op=dist.ReduceOp.SUM)
            e_ms, e_updates, e_d2h = tm[0].item(), tsu[1].item(), tsu[2].item()
        else:
            e_ms, e_updates, e_d2h = te[0].item(), te[1].item(), te[2].item()
        e2e = {"value": e_updates / (e_ms * 1e-3), "unit": "updates/s",
               "h2d_bytes_per_step": nb * 16 * world, "d2h_bytes_per_step": int(e_d2h),
               "ms_per_step": e_ms / args.steps,
               "timer": "host wall clock around the K ticks (device events cannot see the D2H stream)",
               "api": "pedoni_spawn + pedoni_rebuild + pedoni_step + pedoni_download_begin/_end(pos, destination): "
                      "software pipeline: inside the clock, K ticks are computed (k+1 .. k+K) and K ticks' "
                      "pedestrians are delivered to host buffers in the API's types (k .. k+K-1); the payload of "
                      "tick k travels while tick k+1 is computed and tick k-1 is finished on the host (two "
                      "downloads in flight; a whole-domain handle sends destinations as bytes and widens them on "
                      "the host)"}

    clocks = sampler.window(t_wall0, t_wall1) if rank == 0 else None
    sampler.stop()

    if rank == 0:
        peak, peak_src = peaks()
        # One force launch per step on a whole-domain handle; a slab handle adds two small edge launches
        # (ghost-adjacent rows) on its second stream. Quote per step: all force launches of one step
        # and the agents they integrated (device-side live count, not the host's upper bound).
        force_ms = prof["force_ms"] / args.steps
        agents_per_launch = updates / args.steps
        achieved = ALGO_BYTES_PER_UPDATE * agents_per_launch / (force_ms * 1e-3) / 1e9
        step_kernel_ms = {k[:-3]: prof[k] / args.steps for k in prof if k.endswith("_ms")}
        out = {
            "metric": "pedestrian-updates/sec", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"synthetic uniform crowd, {args.agents} pedestrians, open domain "
                                   f"{crowd.side:.0f} m x {crowd.side:.0f} m (BASELINE.json configs[4])",
                       "density_per_m2": args.density, "neighbor_unit_m": 1.4, "field_unit_m": 0.25,
                       "math_mode": args.math, "decomposition": f"{world} row slab(s)",
                       "slab_transport": model.slab_transport(),
                       "field_fetch": "texture gather (atlas)" if model.field_textures() else "global loads",
                       "relax_steps_untimed": args.relax, "active_pedestrians": int(updates_all / args.steps),
                       "l2": "inputs larger than L2 (2 x 24 B x N state + 3 field maps >> 126 MB); no flush"},
            "e2e": e2e,
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": "force_integrate_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (NCU_TRAFFIC_BYTES_PER_UPDATE * agents_per_launch
                                     if args.math == "fast" and args.density == 1.0 else None),
                         "traffic_unit": "bytes per launch (ncu dram read + write, profiles/r01i_force_10M_full.md, "
                                         "scaled by pedestrians per launch)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                         "kernel_ms_per_launch": force_ms, "agents_per_launch": agents_per_launch,
                         "note": "at 1 ped/m^2 the kernel is instruction-issue bound (75 % of issue slots, ncu), "
                                 "not HBM bound: ~18 candidate pairs per update; the field maps add ~205 B of "
                                 "compulsory reads per update to the 48 algorithmic bytes; see DESIGN.md"},
            "kernel_ms_per_step": step_kernel_ms,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:  # the CPU leg is measured once, at N = 1
            out["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(out))
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # Rank 0 prints exactly ONE line on stdout: libraries (NCCL prints its version banner) get stderr.
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real_stdout, "w")
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
