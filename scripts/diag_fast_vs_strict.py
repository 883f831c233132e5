"""One tick from identical states: PEDONI_MATH_FAST vs PEDONI_MATH_STRICT, worst per-agent deviations."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import helpers
from pedoni_b200 import SimulatorOptions, SocialForceModelCuda

name = sys.argv[1] if len(sys.argv) > 1 else "evacuation"
cu, orc = helpers.simulator_pair(name, seed=100, math_mode=0)
field, sc = orc.field, orc.scenario
fast = SocialForceModelCuda(SimulatorOptions(), sc, field, math_mode=1)
for warm in (20, 60, 150, 300):
    while orc.step < warm:
        orc.tick()
    pos, dest, vel, v0 = orc.model.download()
    out = {}
    for key, m in (("strict", cu.model), ("fast", fast)):
        m.upload_state(pos, dest, vel, v0); m.rebuild(); m.step()
        out[key] = m.download()
    ps, ds, vs, _ = out["strict"]; pf, df, vf, _ = out["fast"]
    assert (ds == df).all()
    dv = np.linalg.norm(vs - vf, axis=1)
    order = np.argsort(-np.nan_to_num(dv, nan=1e9))[:6]
    print(f"tick {warm}: n={len(ds)} max|dv|={np.nanmax(dv):.3e} nan strict/fast={np.isnan(vs).any(1).sum()}/{np.isnan(vf).any(1).sum()}")
    for i in order:
        x, y = ps[i] if np.isfinite(ps[i]).all() else pf[i]
        fx, fy = int(x / 0.25 - 0.5), int(y / 0.25 - 0.5)
        print(f"   dv={dv[i]:.3e} pos={ps[i]} dest={ds[i]} v_strict={vs[i]} v_fast={vf[i]} "
              f"dist_tex={field.distance_map[fy-1:fy+3, fx-1:fx+3].round(3).tolist()}")
