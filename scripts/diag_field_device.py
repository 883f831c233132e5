#!/usr/bin/env python
"""GPU box: pedoni_field_build_device (block-iterative eikonal) against pedoni_field_build (the reference's marching
restated on the host) on the shipped scenarios, and its run time on the 10 M synthetic domain."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
from pedoni_b200 import Field  # noqa: E402
from pedoni_b200.synthetic import SyntheticCrowd  # noqa: E402

for name in helpers.scenario_names():
    sc = helpers.load_scenario(name)
    t0 = time.time()
    host = Field.from_scenario(sc, 0.25)
    t1 = time.time()
    dev = Field.from_scenario(sc, 0.25, device=0)
    t2 = time.time()
    free = ~host.obstacle_exist
    assert (host.obstacle_exist == dev.obstacle_exist).all()
    d = np.abs(dev.distance_map - host.distance_map)[free] / 0.25
    line = f"{name:14s} {host.shape} host {t1 - t0:6.2f} s dev {t2 - t1:6.2f} s | distance: max {d.max():.4f} mean {d.mean():.5f} cells"
    worst, rel = 0.0, 0.0
    for k in range(host.potential_maps.shape[0]):
        h, g = host.potential_maps[k][free], dev.potential_maps[k][free]
        reach = h < 1e5 * 0.25  # not through a wall
        dd = np.abs(g - h)[reach] / 0.25
        worst = max(worst, float(dd.max()))
        rel = max(rel, float((np.abs(g - h)[reach] / np.maximum(h[reach], 0.25)).max()))
        above = float(((g - h)[reach] > 1e-4).mean())
    print(line + f" | potentials: max {worst:.4f} cells, max rel {rel:.4f}, device above host on {above:.3%} of cells", flush=True)

crowd = SyntheticCrowd(n=10_000_000)
sc = crowd.scenario()
t0 = time.time()
dev = Field.from_scenario(sc, 0.25, device=0)
t1 = time.time()
closed = crowd.field()
free = ~closed.obstacle_exist
print(f"synthetic 10 M domain {dev.shape}: device builder {t1 - t0:.1f} s; vs closed form: distance max "
      f"{np.abs(dev.distance_map - closed.distance_map)[free].max():.3f} m, potentials max "
      f"{np.abs(dev.potential_maps - closed.potential_maps)[:, free].max():.3f} m")
