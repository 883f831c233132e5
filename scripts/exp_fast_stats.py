#!/usr/bin/env python
"""Experiment: which part of PEDONI_MATH_FAST shifts the evacuation.toml statistic? Builds variants with
-DPEDONI_FAST_FIELD / -DPEDONI_FAST_PAIR and compares the 40 %-evacuation time over 20 seeds with the
strict device path and the oracle.  python scripts/exp_fast_stats.py build | run"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
OUT = ROOT / "build" / "variants"
VARIANTS = {"field1_pair1": (1, 1)}

if sys.argv[1] == "build":
    from pedoni_b200 import build as b
    OUT.mkdir(parents=True, exist_ok=True)
    for name, (f, p) in VARIANTS.items():
        b.build(out=OUT / f"libpedoni_{name}.so", extra=[f"-DPEDONI_FAST_FIELD={f}", f"-DPEDONI_FAST_PAIR={p}"])
elif sys.argv[1] == "run":
    for name in list(VARIANTS) + ["strict"]:
        env = dict(os.environ)
        if name != "strict":
            env["PEDONI_CUDA_LIB"] = str(OUT / f"libpedoni_{name}.so")
        subprocess.run([sys.executable, __file__, "one", name], env=env)
else:
    import numpy as np
    import helpers
    from pedoni_b200 import observables
    name = sys.argv[2]
    mode = 0 if name == "strict" else 1
    scen = sys.argv[3] if len(sys.argv) > 3 else "evacuation"
    t_cu, t_or = [], []
    for seed in range(20):
        cu, orc = helpers.simulator_pair(scen, seed=100 + seed, math_mode=mode)
        for sim, ts in ((cu, t_cu), (orc, t_or)):
            n0 = sim.model.get_pedestrian_count()
            ts.append(observables.evacuation_time(sim.run(800, until_empty=True).active_ped_count, fraction=0.4, initial=n0))
        cu.model.close()
    f = lambda a: f"{np.mean(a):.2f} +- {np.std(a, ddof=1) / np.sqrt(len(a)):.2f}"  # noqa: E731
    print(f"{name:14s} cuda {f(t_cu)}   oracle {f(t_or)}", flush=True)
