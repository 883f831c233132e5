#!/usr/bin/env python
"""Build tuning variants of the force kernel (launch bounds / tile / list depth) into build/variants/ and,
on a GPU box, time each with bench.py (kernel_ms_per_step.force). Usage:
    python scripts/sweep_force.py build            # here (nvcc cross-compiles)
    python scripts/sweep_force.py run [--agents N] # on the B200, via gpurun
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VARIANTS = {  # name: (threads, min_blocks, tile, list)
    "t128_b9_192_32": (128, 9, 192, 32),
    "t128_b10_192_32": (128, 10, 192, 32),
    "t128_b11_160_32": (128, 11, 160, 32),
    "t192_b6_192_32": (192, 6, 192, 32),
    "t256_b4_192_32": (256, 4, 192, 32),
    "t256_b5_192_32": (256, 5, 192, 32),
    "t256_b5_160_32": (256, 5, 160, 32),
    "t512_b2_192_32": (512, 2, 192, 32),
    "t64_b18_192_32": (64, 18, 192, 32),
}
OUT = ROOT / "build" / "variants"


def main():
    mode = sys.argv[1]
    if mode == "build":
        from pedoni_b200 import build as b
        OUT.mkdir(parents=True, exist_ok=True)
        for name, (t, mb, tile, lst) in VARIANTS.items():
            b.build(out=OUT / f"libpedoni_{name}.so",
                    extra=[f"-DPEDONI_FORCE_THREADS={t}", f"-DPEDONI_FORCE_MIN_BLOCKS={mb}",
                           f"-DPEDONI_TILE_ENTRIES={tile}", f"-DPEDONI_LIST_DEPTH={lst}"])
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
                print(f"{name:20s} force {d['kernel_ms_per_step']['force']:.4f} ms  step {d['ms_per_step']:.4f} ms", flush=True)
            except Exception:
                print(name, "FAILED", r.stderr[-400:], flush=True)


if __name__ == "__main__":
    main()
