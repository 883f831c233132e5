"""bench.py prints ONE JSON line with the keys the driver reads. CPU: the reference arm on a small sample.
GPU: the CUDA arm on a small crowd (same code path as the 10 M default, seconds instead of a minute)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(*args):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--cpu-agents", "20000", "--steps", "2", "--warmup", "1")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "pedestrian-updates/sec" and d["unit"] == "updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_cuda_arm_line():
    d = _run("--agents", "300000", "--steps", "5", "--warmup", "3", "--relax", "5", "--cpu-agents", "50000",
             "--cpu-steps", "2")
    assert BASE | {"roofline", "clocks", "kernel_ms_per_step"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["scaling"] == "strong"
    assert d["value"] > 1e8 and d["gpu_launches"] >= 2 * 5  # force + sort per tick
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 9 * 250000
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 48 * 250000
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_under_torchrun_prints_once():
    """The driver launches the reference arm like our own for N > 1: rank 0 alone runs and prints the line,
    the other ranks exit 0 without output."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "bench.py"), "--impl",
                        "reference", "--gpus", "2", "--cpu-agents", "20000", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
