#!/bin/bash
# final multi-GPU evidence: bench at N = 8, 4, 2 on one 8-GPU box (timeline at N = 8)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --timeline gpurun_out/timeline_n8_final.json > gpurun_out/bench_n8_r02z.json 2> gpurun_out/bench_n8_r02z.err
echo "bench8 rc=$?"; grep -E "PedoniError|Error" gpurun_out/bench_n8_r02z.err | head -5
timeout 300 $TR --nproc-per-node 4 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4_r02z.json 2> gpurun_out/bench_n4_r02z.err
echo "bench4 rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_r02z.json 2> gpurun_out/bench_n2_r02z.err
echo "bench2 rc=$?"
python - <<'PY'
import json
for n in (8,4,2):
    f=f'gpurun_out/bench_n{n}_r02z.json'
    try:
        d=json.load(open(f))
        print(n, d['value'], d['ms_per_step'], d['ms_per_step_with_profiling_events'], d.get('slab_parity'), d['kernel_ms_per_step'])
        print('  e2e', {k:v for k,v in d['e2e'].items() if k not in ('api','timer')})
        print('  blocking', d['e2e_blocking']['value'], d['e2e_blocking']['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
