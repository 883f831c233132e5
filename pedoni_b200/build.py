"""Build libpedoni_cuda.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m pedoni_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libpedoni_cuda.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # host pass: the field builder / edge precompute must not contract a*b+c (parity with the
    # reference's f32 arithmetic); g++ from the distro, not the image's $CXX wrapper.
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2,-fopenmp",
    "-shared",
]


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp")) + sorted((CSRC / "host").glob("*.cpp"))


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = list(CSRC.rglob("*.cu")) + list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.cpp")) + \
        list(CSRC.rglob("*.hpp")) + list((PKG.parent / "include").glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: Path | None = None, extra=()) -> Path:
    """`out` / `extra` build a tuning variant elsewhere (scripts/sweep_force.py); the default is the product."""
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found; libpedoni_cuda.so cannot be built (there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", str(out), *map(str, sources()), "-lgomp", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    env = dict(os.environ)
    subprocess.run(cmd, check=True, env=env, stdin=subprocess.DEVNULL, timeout=900)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
