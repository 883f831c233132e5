#!/usr/bin/env python
"""Build tuning variants of the force kernel (launch bounds / tile / list depth) into build/variants/ and,
on a GPU box, time each with bench.py (kernel_ms_per_step.force). Usage:
    python scripts/sweep_force.py build            # here (nvcc cross-compiles)
    python scripts/sweep_force.py run [--agents N] # on the B200, via gpurun
`--set sort` sweeps the elements per thread of the rebuild's scan / scatter / gather kernels instead.
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VARIANTS = {  # name: extra -D flags
    "base": [],
    "unroll1": ["-DPEDONI_FORCE_UNROLL=1"],
    "unroll3": ["-DPEDONI_FORCE_UNROLL=3"],
    "unroll4": ["-DPEDONI_FORCE_UNROLL=4"],
    "list24": ["-DPEDONI_LIST_DEPTH=24"],
    "list40_tile160": ["-DPEDONI_LIST_DEPTH=40", "-DPEDONI_TILE_ENTRIES=160"],
    "tile160": ["-DPEDONI_TILE_ENTRIES=160"],
    "tile224_list24": ["-DPEDONI_TILE_ENTRIES=224", "-DPEDONI_LIST_DEPTH=24"],
    "t192_b6": ["-DPEDONI_FORCE_THREADS=192", "-DPEDONI_FORCE_MIN_BLOCKS=6"],
    "t256_b4": ["-DPEDONI_FORCE_THREADS=256", "-DPEDONI_FORCE_MIN_BLOCKS=4"],
}
SORT_VARIANTS = {"base": [],  # 2 CTAs of 512 threads per SM, 2 pedestrians in flight per thread in the move pass
                 "b3": ["-DPEDONI_SORT_MIN_BLOCKS=3"],
                 "b3_u1": ["-DPEDONI_SORT_MIN_BLOCKS=3", "-DPEDONI_SORT_UNROLL=1"],
                 "t256_b6": ["-DPEDONI_SORT_THREADS=256", "-DPEDONI_SORT_MIN_BLOCKS=6"],
                 "t256_b8_u1": ["-DPEDONI_SORT_THREADS=256", "-DPEDONI_SORT_MIN_BLOCKS=8", "-DPEDONI_SORT_UNROLL=1"]}
if "--set" in sys.argv and sys.argv[sys.argv.index("--set") + 1] == "sort":
    VARIANTS = SORT_VARIANTS
WALL_VARIANTS = {"late_add": ["-DPEDONI_FAR_LOOKUP_EARLY=0"],
                 "late_noadd": ["-DPEDONI_FAR_LOOKUP_EARLY=0", "-DPEDONI_WALL_EARLY_ADD=0"],
                 "early_add": [],  # the defaults
                 "early_noadd": ["-DPEDONI_WALL_EARLY_ADD=0"]}
PREFETCH_VARIANTS = {"base": [], "ahead_1wave": ["-DPEDONI_PREFETCH_AHEAD=170496"], "ahead_2waves": ["-DPEDONI_PREFETCH_AHEAD=340992"],
                     "ahead_half": ["-DPEDONI_PREFETCH_AHEAD=85248"]}
if "--set" in sys.argv and sys.argv[sys.argv.index("--set") + 1] == "prefetch":  # L2 prefetch of a later warp's state
    VARIANTS = PREFETCH_VARIANTS
if "--set" in sys.argv and sys.argv[sys.argv.index("--set") + 1] == "wall":  # far-from-walls mask: where to ask, when to add
    VARIANTS = WALL_VARIANTS
if "--set" in sys.argv and sys.argv[sys.argv.index("--set") + 1] == "debug":  # bounds-checked build for the test suite
    VARIANTS = {"debug": ["-DPEDONI_DEBUG_CHECKS=1"]}
OUT = ROOT / "build" / "variants"


def main():
    mode = sys.argv[1]
    if mode == "build":
        from pedoni_b200 import build as b
        OUT.mkdir(parents=True, exist_ok=True)
        for name, flags in VARIANTS.items():
            b.build(out=OUT / f"libpedoni_{name}.so", extra=flags)
            print("built", name)
    else:
        agents = sys.argv[sys.argv.index("--agents") + 1] if "--agents" in sys.argv else "10000000"
        for name in VARIANTS:
            env = dict(os.environ, PEDONI_CUDA_LIB=str(OUT / f"libpedoni_{name}.so"))
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "10", "--warmup", "3", "--no-e2e",
                                "--no-cpu-baseline", "--relax", "20", "--agents", agents], env=env, capture_output=True,
                               text=True)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                k = d["kernel_ms_per_step"]
                print(f"{name:20s} step {d['ms_per_step']:.4f} ms  " + "  ".join(f"{n} {v:.4f}" for n, v in k.items()),
                      flush=True)
            except Exception:
                print(name, "FAILED", r.stderr[-400:], flush=True)


if __name__ == "__main__":
    main()
