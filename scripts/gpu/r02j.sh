#!/bin/bash
# 2 GPUs: NCCL / peer-memory slab tests, bench with timeline + slab parity, e2e order comparison
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02j_topo.txt 2>&1; nproc >> gpurun_out/r02j_topo.txt
( time timeout 900 python -m pytest tests/test_gpu_nccl_slabs.py -m gpu -q --durations=5 ) > gpurun_out/r02j_pytest.log 2>&1
grep -E "passed|failed|FAILED|real" gpurun_out/r02j_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --timeline gpurun_out/timeline_n2.json > gpurun_out/bench_n2_r02j.json 2> gpurun_out/bench_n2_r02j.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n2_r02j.err
PEDONI_BENCH_E2E_ORDER=plain timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-slab-parity > gpurun_out/bench_n2_plain_r02j.json 2> gpurun_out/bench_n2_plain_r02j.err
python - <<'PY'
import json
for f in ['gpurun_out/bench_n2_r02j.json','gpurun_out/bench_n2_plain_r02j.json']:
    try:
        d=json.load(open(f))
        print(f, d['value'], d['ms_per_step'], d['ms_per_step_without_profiling_events'], d.get('slab_parity'), d['kernel_ms_per_step'])
        print('  e2e', {k:v for k,v in d['e2e'].items() if k not in ('api','timer')})
        print('  blocking', d['e2e_blocking']['value'], d['e2e_blocking']['ms_per_step'], d['config']['cpu_affinity'])
    except Exception as e: print(f, 'ERR', e)
PY
