#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python scripts/diag_headline.py > gpurun_out/r02b_diag.log 2>&1
( time timeout 1500 python -m pytest tests/test_gpu_headline.py tests/test_gpu_scaled_scenarios.py tests/test_gpu_slabs.py -m gpu -q --durations=15 ) > gpurun_out/r02b_pytest_new.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 --deselect tests/test_gpu_headline.py --deselect tests/test_gpu_scaled_scenarios.py --deselect tests/test_gpu_slabs.py ) > gpurun_out/r02b_pytest_rest.log 2>&1
tail -5 gpurun_out/r02b_pytest_new.log gpurun_out/r02b_pytest_rest.log
