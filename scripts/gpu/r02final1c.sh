#!/bin/bash
# HEAD with the far-from-walls mask: the driver's sequence on one GPU, then the ncu evidence of the new force kernel
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 ) > gpurun_out/r02w_pytest.log 2>&1
grep -E "passed|failed|FAILED|real" gpurun_out/r02w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02w.json 2> gpurun_out/bench_r02w.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02w.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['value'], d['e2e_blocking']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_with_field_maps'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:force_integrate|sort_cells|key_kernel|halo|observe|spawn|pack_dest" -c 200 --csv --log-file gpurun_out/launches_r02w.csv \
    python bench.py --steps 5 --warmup 3 --relax 10 --no-cpu-baseline > gpurun_out/ncu_r02w_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_integrate --launch-skip 30 -c 1 \
    -o gpurun_out/force_r02w -f python bench.py --steps 3 --warmup 3 --relax 30 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02w_force.log 2>&1
grep -o '"active_pedestrians": [0-9]*' gpurun_out/ncu_r02w_force.log
ls -la gpurun_out/*.ncu-rep
