#!/bin/bash
# GPU call r02a: new parity tests, baseline bench at HEAD, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_smi.txt
nproc >> gpurun_out/r02a_smi.txt
numactl -H >> gpurun_out/r02a_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02a_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 ) > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02a.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02a.csv \
    python bench.py --steps 3 --warmup 3 --relax 10 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02a_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_integrate --launch-skip 20 -c 1 \
    -o gpurun_out/force_r02a -f python bench.py --steps 3 --warmup 3 --relax 20 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02a_force.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:scan_cells|scatter_kernel|gather_kernel" --launch-skip 60 -c 3 \
    -o gpurun_out/rebuild_r02a -f python bench.py --steps 3 --warmup 3 --relax 20 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02a_rebuild.log 2>&1
ls -la gpurun_out/*r02a*
