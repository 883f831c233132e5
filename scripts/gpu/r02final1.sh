#!/bin/bash
# final 1-GPU evidence of the round: full GPU suite, bench (both arms), strict math, scenarios, ncu captures
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 --durations=8 ) > gpurun_out/r02z_pytest.log 2>&1
grep -E "passed|failed|FAILED|real" gpurun_out/r02z_pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02z_ref.json 2> gpurun_out/bench_r02z_ref.err
echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02z.json 2> gpurun_out/bench_r02z.err
echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --math strict --no-e2e --no-cpu-baseline > gpurun_out/bench_r02z_strict.json 2> gpurun_out/bench_r02z_strict.err
for RHO in 0.05 0.25 4.0; do
timeout 300 python bench.py --agents 2000000 --density $RHO --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_r02z_rho$RHO.json 2> /dev/null
done
timeout 900 python scripts/bench_scenarios.py > gpurun_out/scenarios_r02z.md 2> gpurun_out/scenarios_r02z.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02z.csv \
    python bench.py --steps 3 --warmup 3 --relax 10 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02z_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_integrate --launch-skip 30 -c 1 \
    -o gpurun_out/force_r02z -f python bench.py --steps 3 --warmup 3 --relax 30 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02z_force.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sort_cells --launch-skip 30 -c 1 \
    -o gpurun_out/sort_r02z -f python bench.py --steps 3 --warmup 3 --relax 30 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02z_sort.log 2>&1
grep -o '"active_pedestrians": [0-9]*' gpurun_out/ncu_r02z_force.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02z*.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d.get('kernel_ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('e2e_blocking') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/scenarios_r02z.md
