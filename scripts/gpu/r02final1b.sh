#!/bin/bash
# last check of HEAD on one GPU: the driver's own sequence (pytest -x -m gpu, smoke, bench both arms)
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 ) > gpurun_out/r02zz_pytest.log 2>&1
grep -E "passed|failed|FAILED|real" gpurun_out/r02zz_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02zz.json 2> gpurun_out/bench_r02zz.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02zz.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['value'], d['e2e_blocking']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_with_field_maps']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
