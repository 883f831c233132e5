#!/bin/bash
# final multi-GPU numbers after the size-adaptive sort kernel: N = 8 (timeline) and N = 4
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --timeline gpurun_out/timeline_n8_final.json > gpurun_out/bench_n8_r02z.json 2> gpurun_out/bench_n8_r02z.err
echo "bench8 rc=$?"; grep -E "PedoniError|Error" gpurun_out/bench_n8_r02z.err | head -5
timeout 300 $TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4_r02z.json 2> gpurun_out/bench_n4_r02z.err
echo "bench4 rc=$?"
python - <<'PY'
import json
for n in (8,4):
    d=json.load(open(f'gpurun_out/bench_n{n}_r02z.json'))
    print(n, d['value'], d['ms_per_step'], d['ms_per_step_with_profiling_events'], d.get('slab_parity'), d['kernel_ms_per_step'], d['e2e']['value'], d['e2e_blocking']['value'])
PY
