#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r02c_pytest.log 2>&1
tail -15 gpurun_out/r02c_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_r02c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02c.json'))
print(d['value'], d['ms_per_step'], d['ms_per_step_without_profiling_events'], d['kernel_ms_per_step'], d['e2e']['value'], d['e2e_blocking']['value'])
PY
