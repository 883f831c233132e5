#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_slabs.py tests/test_gpu_golden.py tests/test_gpu_property.py tests/test_gpu_scaled_scenarios.py -m gpu -q -x --timeout=300 2>&1 | tail -3
for A in 1250000 2500000 10000000; do
timeout 300 python bench.py --agents $A --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02v_$A.json 2> gpurun_out/bench_r02v_$A.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02v_$A.json'))
print($A, d['value'], d['ms_per_step'], d['kernel_ms_per_step']['force'], d['kernel_ms_per_step']['sort'])
PY
done
PEDONI_CUDA_LIB=$PWD/build/variants/libpedoni_base.so timeout 300 python bench.py --agents 2500000 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old-base 2500000', d['ms_per_step'], d['kernel_ms_per_step']['sort'])"
