#!/bin/bash
# far-from-walls mask: parity tests, then the same library with and without the mask on the headline workload
set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py tests/test_gpu_slabs.py -m gpu -x -q --timeout=600 ) > gpurun_out/r02wall_pytest.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r02wall_pytest.log | head
for knob in 0 1 0 1; do
  PEDONI_WALL_CUTOFF=$knob timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_wall$knob.json 2> gpurun_out/bench_wall$knob.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_wall$knob.json'))
print("knob $knob", d['value'], d['ms_per_step'], d['kernel_ms_per_step']['force'], d['kernel_ms_per_step']['sort'], d['config'].get('wall_term'), d['roofline_with_field_maps']['frac'])
PY
done
