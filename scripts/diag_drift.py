"""Per-step |CUDA - oracle| drift for both math modes on an unrelaxed and a relaxed dense crowd."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import helpers

sc = helpers.corridor_scenario()
field = helpers.oracle_field(sc)
for relax in (0, 40):
    for mode in (0, 1):
        cu, orc = helpers.make_pair(sc, field, math_mode=mode)
        pos, dest, vel, v0 = helpers.random_crowd(3000, sc.field.size, seed=3, margin=4.0, speed=False)
        orc.spawn(pos, dest, v0)
        for _ in range(relax):
            orc.update(); orc.spawn()
        p, d, v, s = orc.get()
        cu.upload_state(p, d, v, s); cu.rebuild()
        row = []
        for step in range(12):
            cu.step(); orc.update()
            cp, _, cv, _ = cu.download(); op, _, ov, _ = orc.get()
            row.append((float(np.abs(cp - op).max()), float(np.abs(cv - ov).max())))
            cu.rebuild(); orc.spawn()
            assert cu.get_pedestrian_count() == orc.count()
        print(f"relax={relax} mode={mode} n={orc.count()}: " + " ".join(f"{a:.1e}/{b:.1e}" for a, b in row))
        cu.close()
