#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_scaled_scenarios.py tests/test_gpu_slabs.py tests/test_bench_contract.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q -s --durations=8 ) > gpurun_out/r02d_pytest.log 2>&1
grep -E "1 M synthetic|x46|x5|passed|failed|FAILED" gpurun_out/r02d_pytest.log
timeout 900 python scripts/sweep_force.py run --set sort > gpurun_out/r02d_sort_sweep.log 2>&1
cat gpurun_out/r02d_sort_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sort_cells --launch-skip 30 -c 1 \
    -o gpurun_out/sort_r02d -f python bench.py --steps 3 --warmup 3 --relax 30 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02d_sort.log 2>&1
ls -la gpurun_out/sort_r02d.ncu-rep
