#!/bin/bash
# round-2 call N (2 GPUs): the 2-D multi-GPU transform (ndims = 2) -- parity at 1 and 2 ranks, both exchanges; one large timing
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mgpu_cabi.py -m gpu -x -q -k "2d_through" > gpurun_out/n_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/n_pytest.log
export MGPU_TEST_TIMING_ONLY=1
for fl in 1 0; do timeout 100 tests/cpp/_build/test_mgpu 2 32768 1 32768 $fl 5 2>&1 | grep -v "NCCL version" | tee -a gpurun_out/n_2d_timing.txt; done
