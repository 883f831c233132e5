"""Run PEDONI_MATH_STRICT and PEDONI_MATH_FAST side by side from the same seed; report when and where the
trajectories first separate by more than a threshold, and how the gap of that pedestrian grew."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import helpers
from pedoni_b200 import SimulatorOptions, SocialForceModelCuda
from pedoni_b200.simulator import Simulator

name, seed = (sys.argv[1] if len(sys.argv) > 1 else "evacuation"), int(sys.argv[2]) if len(sys.argv) > 2 else 100
sc = helpers.load_scenario(name)
opts = SimulatorOptions()
field = helpers.oracle_field(sc, 0.25)
sims = [Simulator(opts, sc, field, SocialForceModelCuda(opts, sc, field, math_mode=m), seed=seed) for m in (0, 1)]
hist = []
for t in range(600):
    for s in sims:
        s.tick()
    a, b = sims[0].model.download(), sims[1].model.download()
    if len(a[1]) != len(b[1]):
        print(f"tick {t}: populations differ {len(a[1])} vs {len(b[1])}")
        break
    d = np.linalg.norm(a[0] - b[0], axis=1)
    hist.append((a, b, d))
    if np.nanmax(d) > 1e-3:
        i = int(np.nanargmax(d))
        print(f"tick {t}: max |dpos| = {d[i]:.3e} at sorted index {i}: strict pos {a[0][i]} vel {a[2][i]} | fast pos {b[0][i]} vel {b[2][i]} dest {a[1][i]}")
        x, y = a[0][i]
        fx, fy = int(x / 0.25 - 0.5), int(y / 0.25 - 0.5)
        print("dist footprint:\n", field.distance_map[fy - 2:fy + 4, fx - 2:fx + 4].round(3))
        print("pot footprint:\n", field.potential_maps[a[1][i]][fy - 2:fy + 4, fx - 2:fx + 4].round(3))
        # follow the same pedestrian back in time by desired speed (unique)
        v0 = a[3][i]
        for k in range(max(0, len(hist) - 14), len(hist)):
            aa, bb, dd = hist[k]
            j = np.nonzero(aa[3] == v0)[0]
            jb = np.nonzero(bb[3] == v0)[0]
            if len(j) and len(jb):
                print(f"   t={k} strict pos {aa[0][j[0]]} vel {aa[2][j[0]]}  fast pos {bb[0][jb[0]]} vel {bb[2][jb[0]]}  gap {np.linalg.norm(aa[0][j[0]] - bb[0][jb[0]]):.3e}")
        break
else:
    print("no separation > 1e-3 within 600 ticks; final max", np.nanmax(hist[-1][2]))
