"""`python -m pedoni_b200 <scenario.toml> [options]` — the headless loop of the reference binary
(pedoni/src/main.rs:43-136, pedoni/src/args.rs:11-66) over the CUDA backend: parse the scenario TOML,
build the field, tick until --max-steps (or Ctrl-C), write the diagnostic log
`logs/<timestamp>_log.json` in the reference's shape (diagnostic.rs:6-50).

Only the headless mode exists here (the renderer is out of scope), the backend is always CUDA, and the
reference's rate limiter (main.rs:99-103) is not reproduced: the loop runs as fast as the device allows.
"""
from __future__ import annotations

import argparse
import json
import signal
import sys
import time
from pathlib import Path

from . import PEDONI_MATH_FAST, PEDONI_MATH_STRICT, Field, Scenario, SimulatorOptions, SocialForceModelCuda
from .simulator import Simulator, StepMetricsCollection


def parse_args(argv=None):
    ap = argparse.ArgumentParser(prog="python -m pedoni_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("scenario", nargs="?", default="scenarios/default.toml", help="path to scenario file")  # args.rs:14-15
    ap.add_argument("-H", "--headless", action="store_true", help="accepted for compatibility; always headless")
    ap.add_argument("-b", "--backend", default="cuda", choices=["cuda"], help="only the CUDA backend lives here")
    ap.add_argument("--no-distance-map", action="store_true", help="segment walls (sfm.rs:193-237)")  # args.rs:31-32
    ap.add_argument("--field-unit", type=float, default=None)      # args.rs:34-35
    ap.add_argument("--neighbor-unit", type=float, default=None)   # args.rs:37-38
    ap.add_argument("--max-steps", type=int, default=1000)         # args.rs:43-44 (required here: no Ctrl-C loop in CI)
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0x5EED0001, help="spawn stream seed (the reference is unseeded)")
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--log-dir", default="logs")
    ap.add_argument("--device-spawn", nargs="?", const="positions", default=None, choices=["positions", "poisson"],
                    help="draw arrivals on the device: `positions` = positions / desired speeds (pedoni_spawn_groups), "
                         "`poisson` = the per-group Poisson counts as well (pedoni_spawn_poisson); same run, no upload")
    ap.add_argument("--device-field", action="store_true",
                    help="build the field maps on the GPU (pedoni_field_build_device) instead of on the host")
    ap.add_argument("--count-every", type=int, default=1, help="read the population back every N ticks (1 = reference behaviour)")
    return ap.parse_args(argv)


def main(argv=None) -> int:
    args = parse_args(argv)
    opts = SimulatorOptions(use_distance_map=not args.no_distance_map)
    if args.field_unit:
        opts.field_grid_unit = args.field_unit
    if args.neighbor_unit:
        opts.neighbor_grid_unit = args.neighbor_unit
    scenario = Scenario.from_toml(args.scenario)                     # main.rs:54-55
    t0 = time.perf_counter()
    field = Field.from_scenario(scenario, opts.field_grid_unit,      # lib.rs:30
                                device=args.device if args.device_field else None)
    time_calc_field = time.perf_counter() - t0
    model = SocialForceModelCuda(opts, scenario, field, device=args.device,
                                 math_mode=PEDONI_MATH_FAST if args.math == "fast" else PEDONI_MATH_STRICT)
    sim = Simulator(opts, scenario, field, model, seed=args.seed, count_every=args.count_every,
                    device_spawn=args.device_spawn or False)
    stop = {"now": False}
    signal.signal(signal.SIGINT, lambda *_: stop.__setitem__("now", True))   # main.rs:108
    log = StepMetricsCollection()
    t_run = time.perf_counter()
    while not stop["now"] and sim.step < args.max_steps:
        m = sim.tick()                                               # main.rs:86
        log.push(m)
        if sim.step % 100 == 0:                                      # main.rs:87-92
            print(f"Step: {sim.step}, Active pedestrians: {m.active_ped_count}", file=sys.stderr)
    model.synchronize()
    wall = time.perf_counter() - t_run
    out = {  # diagnostic.rs:6-19; `model` / `scenario` are declared but never filled by the reference
        "model": "SocialForceModelCuda", "scenario": str(args.scenario), "total_steps": sim.step,
        "preprocess_metrics": {"time_calc_field": time_calc_field},
        "step_metrics": {"active_ped_count": log.active_ped_count, "time_spawn": log.time_spawn,
                         "time_calc_state": log.time_calc_state, "time_calc_state_kernel": log.time_calc_state_kernel},
    }
    log_dir = Path(args.log_dir)
    log_dir.mkdir(parents=True, exist_ok=True)
    path = log_dir / time.strftime("%Y-%m-%d_%H%M%S_log.json")       # main.rs:118-124
    path.write_text(json.dumps(out))
    updates = sum(log.active_ped_count)
    print(f"Exported log file: {path}  ({sim.step} steps, {updates} pedestrian-updates, {wall:.3f} s)", file=sys.stderr)
    model.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
