#!/usr/bin/env python
"""Regenerates tests/golden/step_vectors.npz: state of the CPU oracle (oracle/, the C++ restatement of
sfm.rs) after 0, 1, 5 and 10 ticks on two small seeded cases — the distance-map wall variant and the
segment wall variant (--no-distance-map). The Rust reference cannot be built in this environment, so
these vectors pin the ORACLE (guarding it against accidental edits) and give the GPU tests a fixture
that needs no oracle call; they are not outputs of the reference binary ("parity unpinned").
    python tests/golden/make_step_vectors.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
import oracle  # noqa: E402
from pedoni_b200 import SimulatorOptions  # noqa: E402


def main():
    oracle.lib().oracle_set_threads(1)
    out = {}
    sc = helpers.corridor_scenario()
    field = helpers.oracle_field(sc)
    for case, use_map in (("distance_map", True), ("segments", False)):
        opts = SimulatorOptions(use_distance_map=use_map)
        m = helpers.OracleAdapter(opts, sc, field)
        pos, dest, vel, v0 = helpers.random_crowd(600, sc.field.size, seed=101, margin=4.0, speed=False)
        out[f"{case}/in_pos"], out[f"{case}/in_dest"], out[f"{case}/in_v0"] = pos, dest, v0
        m.spawn_arrays(pos, dest, v0)
        m.rebuild()
        for tick in range(11):
            if tick in (0, 1, 5, 10):
                p, d, v, s = m.download()
                out[f"{case}/t{tick}_pos"], out[f"{case}/t{tick}_dest"] = p, d
                out[f"{case}/t{tick}_vel"], out[f"{case}/t{tick}_v0"] = v, s
                out[f"{case}/t{tick}_table"] = m.cell_table()
            m.step()
            m.rebuild()
    np.savez_compressed(Path(__file__).with_name("step_vectors.npz"), **out)
    print({k: v.shape for k, v in out.items() if "t10" in k})


if __name__ == "__main__":
    main()
