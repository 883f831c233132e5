#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 ) > gpurun_out/r02i_pytest.log 2>&1
grep -E "passed|failed|FAILED|real" gpurun_out/r02i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02i.json'))
print(d['value'], d['ms_per_step'], d['ms_per_step_without_profiling_events'], d['kernel_ms_per_step'], d['e2e']['value'], d['e2e_blocking']['value'], d['config']['active_pedestrians'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02i.csv \
    python bench.py --steps 3 --warmup 3 --relax 10 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02i_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_integrate --launch-skip 30 -c 1 \
    -o gpurun_out/force_r02i -f python bench.py --steps 3 --warmup 3 --relax 30 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02i_force.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sort_cells --launch-skip 30 -c 1 \
    -o gpurun_out/sort_r02i -f python bench.py --steps 3 --warmup 3 --relax 30 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02i_sort.log 2>&1
grep -o '"active_pedestrians": [0-9]*' gpurun_out/ncu_r02i_force.log gpurun_out/ncu_r02i_sort.log
ls -la gpurun_out/*r02i*
