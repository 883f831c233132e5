#!/usr/bin/env python
"""Diagnostic (GPU box): where does PEDONI_MATH_FAST differ from the oracle on the 1 M synthetic crowd after ONE step
from identical state? Prints the error distribution and, for the worst pedestrians, how close their nearest
neighbour is and how large the oracle's acceleration is (ill-conditioned pairs: the pair term divides by
sqrt(t2^2 - |0.1 v|^2), which cancels when a neighbour sits almost exactly where the other will be in 0.1 s)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from pedoni_b200 import PEDONI_MATH_FAST, PEDONI_MATH_STRICT, SimulatorOptions, SocialForceModelCuda  # noqa: E402
from pedoni_b200.synthetic import SyntheticCrowd  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
crowd = SyntheticCrowd(n=N)
sc, field = crowd.scenario(), crowd.field()
oracle.lib().oracle_set_threads(os.cpu_count() or 1)
orc = oracle.OracleModel(sc.field.size, 1.4, field.unit, field.distance_map, field.potential_maps)
pos, dest, vel, v0 = crowd.agents()
orc.spawn(pos, dest, v0)
for _ in range(5):
    orc.update()
    orc.spawn()
op, od, ov, o0 = orc.get()
orc.update()
acc = orc.accelerations()
np_, _, nv, _ = orc.get()
for mode in (PEDONI_MATH_STRICT, PEDONI_MATH_FAST):
    cu = SocialForceModelCuda(SimulatorOptions(), sc, field, math_mode=mode, capacity=N + 4096)
    cu.upload_state(op, od, ov, o0)
    cu.rebuild()
    cu.step()
    cp, cd, cv, _ = cu.download()
    dp, dv = np.abs(cp - np_).max(1), np.abs(cv - nv).max(1)
    print(f"mode {mode}: n = {len(cd)}")
    for name, d in (("dpos", dp), ("dvel", dv)):
        qs = np.nanquantile(d, [0.5, 0.9, 0.99, 0.999, 0.9999, 0.99999])
        print(f"  {name}: median {qs[0]:.2e} p90 {qs[1]:.2e} p99 {qs[2]:.2e} p99.9 {qs[3]:.2e} p99.99 {qs[4]:.2e} "
              f"p99.999 {qs[5]:.2e} max {np.nanmax(d):.2e}; > 1e-4: {(d > 1e-4).sum()}  > 1e-3: {(d > 1e-3).sum()}")
    amag = np.linalg.norm(acc, axis=1)
    rel = dv / np.maximum(amag * 0.1, 1e-6)
    print(f"  dvel relative to |a| dt: p99.99 {np.nanquantile(rel, 0.9999):.2e} max {np.nanmax(rel):.2e}")
    worst = np.argsort(-np.nan_to_num(dv))[:12]
    for i in worst:
        d = np.linalg.norm(op - op[i], axis=1)
        d[i] = np.inf
        j = int(np.argmin(d))
        t1 = (op[i] - op[j]) - 0.1 * ov[j]
        t2 = d[j] + np.linalg.norm(t1)
        q = t2 * t2 - 0.01 * float(ov[j] @ ov[j])
        clamp = np.linalg.norm(nv[i]) >= 1.3 * o0[i] * 0.999
        print(f"    i={i} dvel={dv[i]:.2e} dpos={dp[i]:.2e} |a|={amag[i]:.3e} nearest={d[j]:.4f} m  q={q:.3e} t2={t2:.3e} "
              f"speed-clamped={bool(clamp)} v_or={nv[i]} v_cu={cv[i]}")
    cu.close()
