"""Small end-to-end flows for compute-sanitizer (memcheck / racecheck / initcheck, one tool per gpurun call):
whole-domain handle in both math modes and wall variants, spawn stream, a 3-slab group through the
in-process transport, pipelined download, capacity growth.
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import helpers
from pedoni_b200 import SimulatorOptions, SlabGroup, SocialForceModelCuda

sc = helpers.corridor_scenario()
field = helpers.oracle_field(sc)
rng = np.random.default_rng(0)
for mode in (0, 1):
    for use_map in (True, False):
        m = SocialForceModelCuda(SimulatorOptions(use_distance_map=use_map), sc, field, math_mode=mode, capacity=256)
        pos, dest, vel, v0 = helpers.random_crowd(700, sc.field.size, seed=1, margin=-2.0)  # some out of grid; grows capacity
        m.upload_state(pos, dest, vel, v0)
        for t in range(6):
            n = int(rng.poisson(10))
            sp = np.stack([np.full(n, 6.0), rng.uniform(5, 25, n)], 1).astype(np.float32)
            m.spawn_arrays(sp, np.ones(n, np.uint32), np.full(n, 1.3, np.float32))
            m.rebuild()
            m.step()
        h_pos, h_dest = np.empty((2000, 2), np.float32), np.empty(2000, np.uint32)
        m.download_begin(h_pos, h_dest)
        m.rebuild(); m.step()
        p, d = m.download_end()
        print("whole", mode, use_map, m.get_pedestrian_count(), len(d), m.cell_table()[-1])
        m.close()
g = SlabGroup(SimulatorOptions(), sc, field, 3, math_mode=1)
pos, dest, vel, v0 = helpers.random_crowd(900, sc.field.size, seed=2, margin=3.5)
vel[:, 1] = np.where(np.arange(len(vel)) % 2 == 0, 1.2, -1.2)
g.upload_state(pos, dest, vel, v0)
for t in range(8):
    g.rebuild(); g.step()
print("slabs", g.get_pedestrian_count(), [s.get_pedestrian_count() for s in g.slabs])
g.close()
print("SANITIZE-SMOKE DONE")
