#!/bin/bash
# 8 GPUs: the real multi-GPU pytest (world 2, 4, 8; peer-memory and NCCL transports; growing scenario) + two new API tests
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_nccl_slabs.py "tests/test_gpu_parity.py::test_profile_timeline_lists_every_launch_of_a_tick" "tests/test_gpu_parity.py::test_library_pinned_buffers_carry_a_download" -m gpu -v --timeout=280 ) > gpurun_out/r02u_pytest.log 2>&1
grep -E "PASSED|FAILED|ERROR|passed|failed|real" gpurun_out/r02u_pytest.log
