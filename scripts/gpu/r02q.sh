#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_slabs.py tests/test_gpu_golden.py tests/test_gpu_property.py tests/test_gpu_headline.py -m gpu -q -x --timeout=300 --durations=4 ) > gpurun_out/r02q_pytest.log 2>&1
grep -E "passed|failed|FAILED|s call" gpurun_out/r02q_pytest.log
for MODE in scatter fused; do for A in 10000000 1250000 300000; do
PEDONI_SORT_MODE=$MODE timeout 300 python bench.py --agents $A --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02q_${MODE}_$A.json 2> gpurun_out/bench_r02q_${MODE}_$A.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02q_${MODE}_$A.json'))
print('$MODE', $A, d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
PY
done; done
