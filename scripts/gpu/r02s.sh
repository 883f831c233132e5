#!/bin/bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02s_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:force_integrate|sort_cells|key_kernel|halo|observe|spawn|pack_dest" -c 200 --csv --log-file gpurun_out/launches_r02z.csv \
    python bench.py --steps 5 --warmup 3 --relax 10 --no-cpu-baseline > gpurun_out/ncu_r02s_list.log 2>&1
echo "list rc=$?"
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_smoke.py ) > gpurun_out/r02s_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/r02s_memcheck.log
( timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_smoke.py ) > gpurun_out/r02s_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/r02s_racecheck.log
