#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_slabs.py tests/test_gpu_golden.py tests/test_gpu_property.py -m gpu -q --durations=5 ) > gpurun_out/r02g_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02g_pytest.log
timeout 900 python scripts/sweep_force.py run --set sort > gpurun_out/r02g_sort_sweep.log 2>&1
cat gpurun_out/r02g_sort_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sort_cells --launch-skip 30 -c 1 \
    -o gpurun_out/sort_r02g -f python bench.py --steps 3 --warmup 3 --relax 30 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02g_sort.log 2>&1
ls -la gpurun_out/sort_r02g.ncu-rep
