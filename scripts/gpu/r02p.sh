#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_slabs.py tests/test_gpu_golden.py tests/test_gpu_property.py tests/test_gpu_headline.py -m gpu -q -x --timeout=300 ) > gpurun_out/r02p_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02p_pytest.log
for P in 1 0; do for A in 10000000 1250000; do
PEDONI_FORCE_PERSISTENT=$P timeout 300 python bench.py --agents $A --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02p_${P}_$A.json 2> gpurun_out/bench_r02p_${P}_$A.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02p_${P}_$A.json'))
print('persistent=$P', $A, d['value'], d['ms_per_step'], d['kernel_ms_per_step']['force'], d['kernel_ms_per_step']['sort'])
PY
done; done
