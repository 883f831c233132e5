#!/usr/bin/env python
"""Throughput of the CUDA path and of the CPU oracle on every BASELINE.json config other than the 10 M
synthetic crowd (which is bench.py's job): the shipped scenarios as they are, and lanes / random scaled
to one million pedestrians (all coordinates, widths and the field size multiplied
by k, 1 M pedestrians seeded once in free space).

    python scripts/bench_scenarios.py [--quick] > profiles/rNN_scenarios.md      (on the B200 box)

updates/s = sum of active pedestrians over the timed ticks / time of (rebuild + step) for those ticks;
GPU time from CUDA events on the handle's stream, CPU time from the oracle's own split timers
(time_spawn + time_calc_state, lib.rs:86,91).
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
import oracle  # noqa: E402
from pedoni_b200 import PEDONI_MATH_FAST, Field, SimulatorOptions, SocialForceModelCuda  # noqa: E402
from pedoni_b200.scenario import ObstacleConfig, Scenario, WaypointConfig, FieldConfig  # noqa: E402
from pedoni_b200.simulator import Simulator  # noqa: E402

QUICK = "--quick" in sys.argv
MILLION_ONLY = "--million-only" in sys.argv  # skip the shipped scenarios
NO_CPU = "--no-cpu" in sys.argv              # skip the CPU oracle legs (kernel A/B runs)


from pedoni_b200.scaled import scaled_scenario as scaled, seed_free_space  # noqa: E402


def time_cuda(model, ticks):
    model.synchronize()
    _, u0 = model.counters()
    model.timer_begin()
    for _ in range(ticks):
        model.step()
        model.rebuild()
    ms = model.timer_end()
    _, u1 = model.counters()
    return (u1 - u0) / (ms * 1e-3), ms / ticks


def row_shipped(name, warm, ticks):
    """A shipped scenario as it is: warm it up through the Simulator (seeded spawn stream), then time."""
    cu, orc = helpers.simulator_pair(name, seed=1, math_mode=PEDONI_MATH_FAST)
    cu.count_every = 10 ** 9
    for _ in range(warm):
        cu.tick()
        orc.tick()
    n = cu.model.get_pedestrian_count()
    cu.model.rebuild()
    gpu, ms = time_cuda(cu.model, ticks)
    updates, ts, tc = orc.model.m.run(ticks)
    cu.model.close()
    return name, n, gpu, ms, updates / (ts + tc)


def row_million(name, k, dests, box_of):
    sc = scaled(helpers.load_scenario(name), k)
    opts = SimulatorOptions()
    t0 = time.time()
    field = Field.from_scenario(sc, opts.field_grid_unit)
    t_field = time.time() - t0
    n = 200_000 if QUICK else 1_000_000
    pos, dest, v0 = seed_free_space(sc, field, n, dests, seed=7, box=box_of(sc))
    model = SocialForceModelCuda(opts, sc, field, math_mode=PEDONI_MATH_FAST, capacity=int(1.05 * n))
    model.spawn_arrays(pos, dest, v0)
    model.rebuild()
    for _ in range(50):  # relax from the zero-velocity seed
        model.step()
        model.rebuild()
    n_live = model.get_pedestrian_count()
    gpu, ms = time_cuda(model, 20)
    p, d, v, s = model.download()
    model.close()
    if NO_CPU:
        return f"{name} x{k:g} ({sc.field.size[0]:.0f} m x {sc.field.size[1]:.0f} m, field build {t_field:.0f} s)", n_live, gpu, ms, float("nan")
    obs = np.array([[*o.line[0], *o.line[1], o.width] for o in sc.obstacles], np.float32).reshape(-1, 5)
    om = oracle.OracleModel(sc.field.size, opts.neighbor_grid_unit, field.unit, field.distance_map, field.potential_maps,
                            obstacles=obs)
    om.set(p, d, v, s)
    om.run(1)
    updates, ts, tc = om.run(3)
    return f"{name} x{k:g} ({sc.field.size[0]:.0f} m x {sc.field.size[1]:.0f} m, field build {t_field:.0f} s)", n_live, gpu, ms, updates / (ts + tc)


def main():
    oracle.lib().oracle_set_threads(__import__("os").cpu_count() or 1)
    rows = []
    for name, warm, ticks in [] if MILLION_ONLY else [("default", 600, 200), ("narrow-gap", 50, 100), ("bottleneck", 600, 100),
                                                     ("evacuation", 30, 100), ("lanes", 600, 200), ("random", 600, 100)]:
        rows.append(row_shipped(name, warm // (4 if QUICK else 1), ticks))
        print("done", rows[-1][0], file=sys.stderr)
    rows.append(row_million("lanes", 46.0, [0, 1], lambda sc: (0.09 * sc.field.size[0], 0.5, 0.91 * sc.field.size[0], 8.0 * 46 - 0.5)))
    print("done lanes 1M", file=sys.stderr)
    rows.append(row_million("random", 5.0, [0, 1, 2, 3], lambda sc: None))
    print("| config | pedestrians | CUDA updates/s | CUDA ms/tick | CPU oracle updates/s (%d threads) | ratio |" % oracle.lib().oracle_max_threads())
    print("|---|---|---|---|---|---|")
    for name, n, gpu, ms, cpu in rows:
        print(f"| {name} | {n} | {gpu:.3e} | {ms:.4f} | {cpu:.3e} | {gpu / cpu:.1f} |")


if __name__ == "__main__":
    main()
