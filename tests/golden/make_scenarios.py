#!/usr/bin/env python
"""Regenerates tests/golden/scenarios.npz from the reference's shipped scenario files.

Run in the build container only (`/root/reference` does not exist on the GPU box):
    python tests/golden/make_scenarios.py
Each scenario TOML (scenarios/*.toml, parsed with pedoni_b200.Scenario — the same serde-shaped loader
users call) becomes four numeric arrays: `<name>/size` (2,), `<name>/waypoints` (n, 5) and
`<name>/obstacles` (n, 5) as x0, y0, x1, y1, width, and `<name>/pedestrians` (n, 4) as origin,
destination, kind (0 periodic / 1 once), value (frequency or count). tests/helpers.load_scenario()
turns them back into Scenario objects. Byte-identical duplicates (s-shape == default, bottleneck2 ==
bottleneck) are stored once and aliased.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from pedoni_b200 import Scenario  # noqa: E402

SRC = Path("/root/reference/scenarios")


def main():
    out, seen, alias = {}, {}, []
    for path in sorted(SRC.glob("*.toml")):
        name = path.stem
        digest = hashlib.sha256(path.read_bytes()).hexdigest()
        if digest in seen:
            alias.append(f"{name}={seen[digest]}")
            continue
        seen[digest] = name
        sc = Scenario.from_toml(path)
        out[f"{name}/size"] = np.asarray(sc.field.size, np.float64)
        out[f"{name}/waypoints"] = np.asarray([[*w.line[0], *w.line[1], w.width] for w in sc.waypoints],
                                              np.float64).reshape(-1, 5)
        out[f"{name}/obstacles"] = np.asarray([[*o.line[0], *o.line[1], o.width] for o in sc.obstacles],
                                              np.float64).reshape(-1, 5)
        out[f"{name}/pedestrians"] = np.asarray(
            [[p.origin, p.destination, 0 if p.spawn.kind == "periodic" else 1,
              p.spawn.frequency if p.spawn.kind == "periodic" else p.spawn.count] for p in sc.pedestrians],
            np.float64).reshape(-1, 4)
    out["__aliases__"] = np.asarray(alias)
    np.savez_compressed(Path(__file__).with_name("scenarios.npz"), **out)
    print("scenarios:", sorted(seen.values()), "aliases:", alias)


if __name__ == "__main__":
    main()
