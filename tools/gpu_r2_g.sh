#!/bin/bash
# round-2 call G (2 GPUs): SM-partition sweep (CTA cap of the link-bound launches x SMs the HBM-bound launches leave free)
set -u
mkdir -p gpurun_out
exe=tests/cpp/_build/test_mgpu
timeout 120 $exe 2 256 256 256 1 0 2>&1 | grep -v NCCL | tail -2
timeout 120 $exe 2 256 256 256 0 0 2>&1 | grep -v NCCL | tail -2
export MGPU_TEST_TIMING_ONLY=1
MGPU_TEST_SWEEP="${SWEEP_P2P}" timeout 200 $exe 2 1024 1024 1024 1 10 2>&1 | grep -v NCCL > gpurun_out/g_sweep_p2p_g2.jsonl
MGPU_TEST_SWEEP="${SWEEP_NCCL}" timeout 200 $exe 2 1024 1024 1024 0 10 2>&1 | grep -v NCCL > gpurun_out/g_sweep_nccl_g2.jsonl
cat gpurun_out/g_sweep_p2p_g2.jsonl gpurun_out/g_sweep_nccl_g2.jsonl | cut -c 36-200
