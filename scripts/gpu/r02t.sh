#!/bin/bash
set -x
mkdir -p gpurun_out
PEDONI_CUDA_LIB=$PWD/build/variants/libpedoni_prefetch1.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_headline.py tests/test_gpu_slabs.py -m gpu -q -x --timeout=300 2>&1 | tail -3
PEDONI_CUDA_LIB=$PWD/build/variants/libpedoni_prefetch2.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x --timeout=300 2>&1 | tail -3
timeout 900 python scripts/sweep_force.py run > gpurun_out/r02t_sweep.log 2>&1; cat gpurun_out/r02t_sweep.log
