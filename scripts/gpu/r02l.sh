#!/bin/bash
# 8 GPUs: bench at N = 8 (timeline + slab parity) and N = 4
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02l_topo.txt 2>&1; nproc >> gpurun_out/r02l_topo.txt; free -g >> gpurun_out/r02l_topo.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --timeline gpurun_out/timeline_n8.json > gpurun_out/bench_n8_r02l.json 2> gpurun_out/bench_n8_r02l.err
echo "bench8 rc=$?"; grep -E "PedoniError|Error" gpurun_out/bench_n8_r02l.err | head -5
timeout 300 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4_r02l.json 2> gpurun_out/bench_n4_r02l.err
echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/bench_n8_r02l.json','gpurun_out/bench_n4_r02l.json']:
    try:
        d=json.load(open(f))
        print(f, d['value'], d['ms_per_step'], d['ms_per_step_without_profiling_events'], d.get('slab_parity'), d['kernel_ms_per_step'])
        print('  e2e', {k:v for k,v in d['e2e'].items() if k not in ('api','timer')})
        print('  blocking', d['e2e_blocking']['value'], d['e2e_blocking']['ms_per_step'], d['config']['cpu_affinity'])
    except Exception as e: print(f, 'ERR', e)
PY
