#!/usr/bin/env python
"""bench.py — pedestrian-updates/sec of the per-timestep pedestrian update on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--agents 10000000] [--density 1.0] [--math fast|strict]

One "step" = one tick of the hot path over the whole crowd: neighbor-grid rebuild (pedoni_rebuild)
+ forces and integration (pedoni_step). Workload = BASELINE.json configs[4]: the synthetic uniform
crowd of 10 M pedestrians on a large open domain (pedoni_b200/synthetic.py); it fits one GPU, so it is
also the N = 1 workload. With --gpus N the SAME 10 M crowd is slab-decomposed over N ranks
(strong scaling, as the north star states the target on a fixed 10 M crowd).

Prints ONE JSON line (rank 0).
  value        updates/s with state resident in HBM, K ticks between two CUDA events (max over ranks).
  e2e          the same tick driven through the C ABI with HOST buffers: per step a spawn batch is copied
               host->device and the trait's list_pedestrians payload (pos + destination) is copied device->host,
               inside the timed region, software-pipelined (pedoni_download_begin / _end).
  e2e_blocking the call sequence the reference's loop makes (pedoni/src/main.rs:86-97): tick, then the whole list on
               the host (blocking pedoni_download into pinned buffers), then the next tick.
  roofline     the dominant kernel (force + integrate) from per-launch CUDA events inside the timed region;
               `traffic` from profiles/r02_traffic.json (one `ncu --set full` launch of the benched binary).
  slab_parity  N > 1: after the timed region rank 0 runs the whole-domain handle for the same ticks and compares
               a hash of every slab's pedestrians with the matching range of its own: "bitwise" or "mismatch".
  cpu_baseline the C++ restatement of the Rust reference (oracle/) on this box's host cores, bounded sample.
`--impl reference` times that restatement on the FULL workload (there is no Rust toolchain to build the reference).
"""
from __future__ import annotations

import argparse
import hashlib
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
TRAFFIC_FILE = ROOT / "profiles" / "r02_traffic.json"  # written by scripts/ncu_summary.py traffic (ncu --set full)
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
    ap.add_argument("--cpu-agents", type=int, default=None,
                    help="pedestrians of the CPU legs: default = --agents for --impl reference (the full workload, "
                         "~0.8 s/tick on 16 cores), 4 000 000 for the in-line cpu_baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=20, help="ticks of the cpu_baseline leg (~10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-slab-parity", action="store_true")
    ap.add_argument("--relax", type=int, default=RELAX_STEPS)
    ap.add_argument("--timeline", default=None, help="write rank 0's two-stream launch timeline of the timed ticks here")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(args, side):
    return (f"synthetic uniform crowd, {args.agents} pedestrians, open domain {side:.0f} m x {side:.0f} m "
            "(BASELINE.json configs[4])")


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


def bind_to_gpu_numa_node(pci_bus_id: str):
    """Run this rank (and allocate its pinned buffers) on the NUMA node its GPU hangs off, so that the per-tick
    device-to-host payload does not cross the socket interconnect. Returns a description or None."""
    try:
        node = int(Path(f"/sys/bus/pci/devices/{pci_bus_id.lower()}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"NUMA node {node} ({len(cpus)} cpus)"
    except (OSError, ValueError):
        return None


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path, i.e. (no Rust toolchain in
    this image) its C++ restatement in oracle/, on all host threads, on the full workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle
    from pedoni_b200.synthetic import SyntheticCrowd

    n_cpu = args.cpu_agents or args.agents
    crowd = SyntheticCrowd(n=n_cpu, density=args.density)
    field = build_field_for_cpu_leg(crowd)
    sc = crowd.scenario()
    m = oracle.OracleModel(sc.field.size, 1.4, field.unit, field.distance_map, field.potential_maps)
    chunk = 2_000_000
    parts = [crowd.agents(lo, min(lo + chunk, n_cpu)) for lo in range(0, n_cpu, chunk)]
    import numpy as np
    m.spawn(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
            np.concatenate([p[3] for p in parts]))
    del parts
    oracle.lib().oracle_set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: use every host core
    cores = oracle.lib().oracle_max_threads()
    m.run(max(args.warmup, 1))
    t0 = time.time()
    updates, ts, tc = m.run(args.steps)
    wall = time.time() - t0
    value = updates / (ts + tc)
    full = n_cpu == args.agents
    sample = (f"the full workload: {n_cpu} pedestrians" if full else f"{n_cpu} pedestrians of the same synthetic crowd") + \
        f" (density {args.density}/m^2), {args.steps} ticks"
    full_side = SyntheticCrowd(n=args.agents, density=args.density).side
    config = {"workload": workload_name(args, full_side), "density_per_m2": args.density, "neighbor_unit_m": 1.4,
              "field_unit_m": 0.25,
              "implementation": "C++ restatement of the Rust reference (oracle/): the Rust toolchain is absent"}
    if not full:
        config["sample_agents"] = n_cpu
    print(json.dumps({
        "impl": "reference", "metric": "pedestrian-updates/sec", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (ts + tc) / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "time_spawn_s": ts, "time_calc_state_s": tc, "wall_s": wall},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def build_field_for_cpu_leg(crowd):
    """The CPU legs walk on the same kind of field as the CUDA arm (device-built) when the box has a GPU; the
    closed form stands in on a box without one (the field is an input; building it is not timed)."""
    try:
        import torch
        if torch.cuda.is_available():
            return crowd.field(device=0)
    except Exception:
        pass
    return crowd.field()


def cpu_baseline(args):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle
    from pedoni_b200.synthetic import SyntheticCrowd

    n_cpu = args.cpu_agents or min(args.agents, 4_000_000)
    crowd = SyntheticCrowd(n=n_cpu, density=args.density)
    field = build_field_for_cpu_leg(crowd)
    sc = crowd.scenario()
    pos, dest, vel, v0 = crowd.agents()
    m = oracle.OracleModel(sc.field.size, 1.4, field.unit, field.distance_map, field.potential_maps)
    m.spawn(pos, dest, v0)
    oracle.lib().oracle_set_threads(os.cpu_count() or 1)
    m.run(1)
    updates, ts, tc = m.run(args.cpu_steps)
    return {"value": updates / (ts + tc), "unit": "updates/s", "cores": oracle.lib().oracle_max_threads(),
            "kind": "port",
            "sample": f"{n_cpu} agents of the same synthetic crowd, {args.cpu_steps} ticks; "
                      "C++ restatement of the Rust reference (rayon force loop -> OpenMP; rebuild and "
                      "integration serial as in sfm.rs)",
            "time_spawn_s": ts, "time_calc_state_s": tc}


def traffic_record():
    """DRAM bytes per pedestrian-update of the force kernel from the tracked ncu capture, or None."""
    try:
        d = json.loads(TRAFFIC_FILE.read_text())
        k = d["kernels"]["force_integrate_kernel"]
        return (k["dram_read_bytes"] + k["dram_write_bytes"]) / k["agents"], d
    except (OSError, KeyError, ValueError, ZeroDivisionError):
        return None, None


def state_digest(parts):
    """One hash over the bytes of (pos, destination, velocity, desired speed) in the model's order."""
    import numpy as np
    h = hashlib.blake2b(digest_size=16)
    for a in parts:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


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
    all_cpus = os.sched_getaffinity(0)
    torch.cuda.set_device(local_rank)
    props = torch.cuda.get_device_properties(local_rank)
    affinity = None
    if os.environ.get("PEDONI_BENCH_NUMA", "1") != "0" and hasattr(props, "pci_bus_id"):
        affinity = bind_to_gpu_numa_node(f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(values):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if world == 1:
            return t.tolist(), t.tolist()
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        return tmax.tolist(), tsum.tolist()

    crowd = SyntheticCrowd(n=args.agents, density=args.density)
    sc = crowd.scenario()
    # the field maps are an input of the path: built here by the library's device builder (f1), like the
    # reference's Simulator::new builds them with Field::from_scenario before the model exists (lib.rs:30)
    t_field = time.time()
    field = crowd.field(device=local_rank)
    t_field = time.time() - t_field
    one_map = field.distance_map.size * 4
    map_bytes = (1 + field.potential_maps.shape[0]) * one_map
    opts = pb.SimulatorOptions()
    math_mode = pb.PEDONI_MATH_FAST if args.math == "fast" else pb.PEDONI_MATH_STRICT
    model = pb.SocialForceModelCuda(opts, sc, field, device=local_rank, math_mode=math_mode,
                                    capacity=int(args.agents * 1.05 / world) + 65536,
                                    slab_rank=rank, slab_count=world)
    if world > 1:
        uid = [pb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        model.comm_init(uid[0])
    keep_field = world > 1 and rank == 0 and not args.no_slab_parity
    if not keep_field:
        del field

    # Every rank generates the crowd in chunks and keeps the agents whose rows it owns (the library
    # drops foreign rows of replicated spawn lists).
    chunk = 2_000_000

    def seed_crowd(m):
        for lo in range(0, args.agents, chunk):
            pos, dest, vel, v0 = crowd.agents(lo, min(lo + chunk, args.agents))
            m.spawn_arrays(pos, dest, v0)

    seed_crowd(model)
    barrier()  # ranks generate the crowd at different speeds; the first ghost exchange should find everybody there
    model.rebuild()
    n0 = model.get_pedestrian_count()
    ticks_done = 0

    def tick():
        model.step()
        model.rebuild()

    for _ in range(args.relax):
        tick()
    ticks_done += args.relax
    model.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: state resident in HBM ------------------------------------------------------------
    for _ in range(args.warmup):
        tick()
    ticks_done += args.warmup
    l0, u0 = model.counters()
    barrier()
    t_wall0 = time.time()
    model.timer_begin()
    for _ in range(args.steps):
        tick()
    ms = model.timer_end()
    barrier()
    t_wall1 = time.time()
    ticks_done += args.steps
    l1, u1 = model.counters()
    updates = u1 - u0
    (ms_max, _, _), (_, updates_all, launches_all) = reduce_max_sum([ms, float(updates), float(l1 - l0)])
    value = updates_all / (ms_max * 1e-3)

    # the same K ticks once more with a CUDA event pair around every launch (per-kernel times for the roofline and
    # the two-stream timeline); the events themselves cost a few per cent of a short tick, so `value` is the pass above
    model.profile_enable(True)
    model.profile_reset()
    barrier()
    model.timer_begin()
    for _ in range(args.steps):
        tick()
    ms_prof = model.timer_end()
    barrier()
    ticks_done += args.steps
    prof = model.profile_read()
    if args.timeline and rank == 0:
        Path(args.timeline).write_text(json.dumps({"n_gpus": world, "steps": args.steps, "ms_total": ms_prof,
                                                   "launches": model.profile_timeline()}))
    model.profile_enable(False)
    (ms_prof_max, _), _ = reduce_max_sum([ms_prof, 0.0])

    # ---- slab parity: the N slabs against the whole-domain handle, same ticks ------------------------
    slab_parity = None
    if world > 1 and not args.no_slab_parity:
        part = model.download()  # pos, dest, vel, v0 of the owned rows
        mine = (len(part[1]), state_digest(part))
        del part
        got = [None] * world
        dist.all_gather_object(got, mine)
        if rank == 0:
            whole = pb.SocialForceModelCuda(opts, sc, field, device=local_rank, math_mode=math_mode,
                                            capacity=int(args.agents * 1.02) + 65536)
            del field
            seed_crowd(whole)
            whole.rebuild()
            for _ in range(ticks_done):
                whole.step()
                whole.rebuild()
            w = whole.download()
            whole.close()
            ok, lo = len(w[1]) == sum(g[0] for g in got), 0
            for n_r, digest in got:
                ok = ok and state_digest([a[lo:lo + n_r] for a in w]) == digest
                lo += n_r
            slab_parity = {"result": "bitwise" if ok else "mismatch", "ticks": ticks_done, "pedestrians": len(w[1]),
                           "per_slab": [g[0] for g in got],
                           "what": "blake2b of every slab's (pos, destination, vel, desired_speed) bytes == the same hash "
                                   "of the matching range of a whole-domain handle run for the same ticks on rank 0"}
            del w
        barrier()

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ---------------------
    e2e = e2e_blocking = None
    if not args.no_e2e:
        cap = int(n0 * 1.02) + 8 * E2E_SPAWN_PER_STEP * (2 * (args.steps + args.warmup) + 8)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
        # list_pedestrians' payload: position + destination, the destination as one byte (pedoni_download_begin_u8)
        h_pos, h_dest = pin((cap, 2), torch.float32), pin((cap,), torch.uint8)
        nb = E2E_SPAWN_PER_STEP
        s_pos, s_dest, s_v0 = pin((nb, 2), torch.float32), pin((nb,), torch.int32).view(np.uint32), \
            pin((nb,), torch.float32)
        extra = SyntheticCrowd(n=args.agents, density=args.density, seed=crowd.seed ^ 0xE2E)
        h_pos2, h_dest2 = pin((cap, 2), torch.float32), pin((cap,), torch.uint8)
        h_dest32 = pin((cap,), torch.int32).view(np.uint32)  # the blocking variant delivers the trait's 4-byte type
        bufs = [(h_pos, h_dest), (h_pos2, h_dest2)]
        state = {"inflight": 0, "n": 0, "bytes": 0}

        # concurrent pinned device->host ceiling of this box at N ranks: what any per-tick read-back is bound by
        d_probe = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        h_probe = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        h_probe.copy_(d_probe)
        barrier()
        t0 = time.perf_counter()
        for _ in range(8):
            h_probe.copy_(d_probe, non_blocking=True)
        torch.cuda.synchronize()
        t_probe = time.perf_counter() - t0
        (t_probe_max, ), _ = reduce_max_sum([t_probe])
        d2h_ceiling = world * 8 * (64 << 20) / t_probe_max / 1e9
        del d_probe, h_probe

        def collect(keep=0):
            """Finish the pipelined list_pedestrians of earlier ticks (all but the newest `keep`): their payload
            is on the host, in the API's types, when this returns."""
            while state["inflight"] > keep:
                pos, dest = model.download_end()
                state["n"] += pos.shape[0]
                state["bytes"] = pos.shape[0] * (8 + dest.itemsize)  # what crossed PCIe for this tick
                state["inflight"] -= 1

        # the synthetic inflow (uniform over the domain) is drawn before the clock starts: generating random
        # numbers in numpy is not part of the path, copying them into the pinned staging buffers is
        n_ticks = 2 * (max(args.warmup, 1) + args.steps) + 4
        inflow = [extra.agents(k * nb, (k + 1) * nb) for k in range(n_ticks)]

        def compute(k):
            p, d, _, v = inflow[k]
            s_pos[:], s_dest[:], s_v0[:] = p, d, v
            model.spawn_arrays(s_pos, s_dest, s_v0)   # H2D (spawn_pedestrians, first half)
            model.rebuild()                           # spawn_pedestrians, second half
            model.step()                              # update_states

        # `eager`: tick k+1 is enqueued BEFORE tick k-1's download is finished on the host, so the GPU computes while
        # the host waits for the copy (and, on a whole-domain handle, widens the byte-sized destinations).
        eager = os.environ.get("PEDONI_BENCH_E2E_ORDER", "eager") != "plain"

        def e2e_tick(k):
            """Steady state of the pipeline: tick k has been enqueued. Start its list_pedestrians (device
            snapshot + async D2H), enqueue tick k+1 and finish tick k-1 on the host — in the order `eager` says."""
            model.download_begin(*bufs[k % 2])
            state["inflight"] += 1
            if eager:
                compute(k + 1)
                collect(keep=1)
            else:
                collect(keep=1)
                compute(k + 1)

        compute(0)
        k0 = max(args.warmup, 1)
        for k in range(k0):
            e2e_tick(k)
        collect()
        state["n"] = 0
        model.synchronize()
        barrier()
        t0 = time.perf_counter()
        for k in range(k0, k0 + args.steps):          # K x (one tick computed, one tick's pedestrians delivered)
            e2e_tick(k)
        collect()                                     # every timed tick's result has been read on the host
        model.synchronize()
        wall_e2e = (time.perf_counter() - t0) * 1e3
        barrier()
        (e_ms, _, _), (_, e_updates, e_d2h) = reduce_max_sum([wall_e2e, float(state["n"]), float(state["bytes"])])
        e2e = {"value": e_updates / (e_ms * 1e-3), "unit": "updates/s",
               "h2d_bytes_per_step": nb * 16 * world, "d2h_bytes_per_step": int(e_d2h),
               "ms_per_step": e_ms / args.steps, "d2h_ceiling_gbs": d2h_ceiling,
               "d2h_achieved_gbs": e_d2h / (e_ms / args.steps * 1e-3) / 1e9,
               "order": "eager" if eager else "plain",
               "timer": "host wall clock around the K ticks (device events cannot see the D2H stream)",
               "api": "pedoni_spawn + pedoni_rebuild + pedoni_step + pedoni_download_begin_u8/_end(pos f32x2, destination u8): "
                      "software pipeline: inside the clock, K ticks are computed (k+1 .. k+K) and K ticks' "
                      "pedestrians are delivered to pinned host buffers (k .. k+K-1); the payload of tick k travels "
                      "while tick k+1 is computed and tick k-1 is finished on the host (two downloads in flight). "
                      "d2h_ceiling_gbs = aggregate pinned D2H rate of this box with all N ranks copying "
                      "at once (8 x 64 MiB each)"}

        # ---- the reference's own call sequence: tick, whole list on the host, next tick (main.rs:86-97) --------
        kb = k0 + args.steps + 1
        out = (h_pos, h_dest32, None, None)

        def blocking_tick(k):
            p, d, _, v = inflow[k]
            s_pos[:], s_dest[:], s_v0[:] = p, d, v
            model.spawn_arrays(s_pos, s_dest, s_v0)
            model.rebuild()
            model.step()
            return model.download(vel=False, v0=False, out=out)[1].shape[0]   # blocks: list_pedestrians

        for k in range(kb, kb + max(args.warmup, 1)):
            blocking_tick(k)
        barrier()
        t0 = time.perf_counter()
        n_blk = 0
        for k in range(kb + max(args.warmup, 1), kb + max(args.warmup, 1) + args.steps):
            n_blk += blocking_tick(k)
        wall_blk = (time.perf_counter() - t0) * 1e3
        barrier()
        (b_ms, _), (_, b_updates) = reduce_max_sum([wall_blk, float(n_blk)])
        e2e_blocking = {"value": b_updates / (b_ms * 1e-3), "unit": "updates/s", "ms_per_step": b_ms / args.steps,
                        "d2h_bytes_per_step": int(b_updates / args.steps * 12),
                        "api": "pedoni_spawn + pedoni_rebuild + pedoni_step + blocking pedoni_download(pos, destination) "
                               "into pinned buffers, every tick: Simulator::tick() then list_pedestrians() "
                               "(pedoni/src/main.rs:86-97), nothing overlapped"}

    clocks = sampler.window(t_wall0, t_wall1) if rank == 0 else None
    sampler.stop()

    if rank == 0:
        peak, peak_src = peaks()
        # One interior force launch per step (the whole crowd on a whole-domain handle); a slab handle adds
        # launches on the rows next to a slab boundary on its second stream, reported separately.
        force_ms = prof["force_ms"] / args.steps
        agents_interior = prof["force_agents"] / max(prof["force_launches"], 1)  # host upper bound of the grid
        agents_per_launch = updates / args.steps  # device-side live count of this rank (owned pedestrians)
        if world > 1:  # the interior launch integrates the owned rows minus the two boundary rows each side
            agents_per_launch = min(agents_per_launch, agents_interior)
        achieved = ALGO_BYTES_PER_UPDATE * agents_per_launch / (force_ms * 1e-3) / 1e9
        step_kernel_ms = {name: prof[key] / args.steps for name, key in (
            ("key", "key_ms"), ("sort", "gather_ms"), ("force", "force_ms"), ("force_edge", "force_edge_ms"),
            ("pack", "pack_ms"), ("exchange_unpack", "comm_ms"))}
        per_update, traffic_doc = traffic_record()
        use_traffic = per_update is not None and args.math == "fast" and args.density == 1.0 and \
            abs(traffic_doc.get("agents_total", 0) - args.agents) <= 0.01 * args.agents  # captured at this workload
        # the distance map is not fetched where the wall term is below 1e-17 m/s^2 (pedoni_wall_far_cells)
        far_blocks, mask_blocks = model.wall_far_cells()
        near_fraction = 1.0 - far_blocks / mask_blocks if mask_blocks else 1.0
        compulsory = ALGO_BYTES_PER_UPDATE + (map_bytes - one_map + near_fraction * one_map) / args.agents
        out = {
            "metric": "pedestrian-updates/sec", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, crowd.side),
                       "density_per_m2": args.density, "neighbor_unit_m": 1.4, "field_unit_m": 0.25,
                       "math_mode": args.math, "decomposition": f"{world} row slab(s)",
                       "slab_transport": model.slab_transport(),
                       "field_fetch": "texture gather (atlas)" if model.field_textures() else "global loads",
                       "wall_term": (f"distance map not fetched in {far_blocks} of {mask_blocks} 2 m blocks: more than 8 m "
                                     "from every obstacle, term < 1e-17 m/s^2 (pedoni_wall_far_cells; PEDONI_WALL_CUTOFF=0 "
                                     "evaluates it everywhere)") if far_blocks else "evaluated everywhere",
                       "field_builder": f"pedoni_field_build_device (block-iterative eikonal on the GPU), {t_field:.1f} s, untimed",
                       "relax_steps_untimed": args.relax, "active_pedestrians": int(updates_all / args.steps),
                       "cpu_affinity": affinity,
                       "l2": "inputs larger than L2 (2 x 24 B x N state + 3 field maps >> 126 MB); no flush"},
            "ms_per_step_with_profiling_events": ms_prof_max / args.steps,
            "kernel_timing": "kernel_ms_per_step and roofline.kernel_ms_per_launch come from a second pass of the same K "
                             "ticks with a CUDA event pair around every launch (ms_per_step_with_profiling_events); "
                             "`value` is the pass without them",
            "e2e": e2e,
            "e2e_blocking": e2e_blocking,
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": "force_integrate_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": per_update * agents_per_launch if use_traffic else None,
                         "traffic_unit": "bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of one ncu "
                                         "--set full launch (profiles/r02_traffic.json), per pedestrian, x the "
                                         "pedestrians of this launch",
                         "traffic_source": ({"file": str(TRAFFIC_FILE.relative_to(ROOT)), "head": traffic_doc.get("head"),
                                             "bytes_per_update": per_update} if use_traffic else None),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                         "kernel_ms_per_launch": force_ms, "agents_per_launch": agents_per_launch,
                         "note": "at 1 ped/m^2 the kernel is instruction-issue bound, not HBM bound: ~18 candidate "
                                 "pairs per update; see roofline_with_field_maps and DESIGN.md"},
            # the bytes the path cannot avoid moving at this density: the 48 algorithmic bytes plus every field map
            # once per tick (each pedestrian gathers 4x4-texel footprints of two 0.25 m maps; at 1 ped/m^2 nearly
            # every sector of the maps is touched once per step)
            "roofline_with_field_maps": {"bound": "hbm", "kernel": "force_integrate_kernel",
                                         "bytes_per_update": compulsory,
                                         "achieved": compulsory * agents_per_launch / (force_ms * 1e-3) / 1e9,
                                         "peak": peak, "unit": "GB/s",
                                         "frac": compulsory * agents_per_launch / (force_ms * 1e-3) / 1e9 / peak,
                                         "formula": "48 + (n_potential_maps + fraction of the distance map within 8 m of an "
                                                    "obstacle or on a ridge) * field_ny * field_nx * 4 / N"},
            "kernel_ms_per_step": step_kernel_ms,
            "clocks": clocks,
        }
        if slab_parity is not None:
            out["slab_parity"] = slab_parity["result"]
            out["slab_parity_detail"] = slab_parity
        if not args.no_cpu_baseline and world == 1:  # the CPU leg is measured once, at N = 1
            os.sched_setaffinity(0, all_cpus)        # the oracle gets every host core, not one NUMA node
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
