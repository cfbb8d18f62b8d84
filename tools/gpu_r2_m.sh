#!/bin/bash
# round-2 call M (2 GPUs): NCCL exchange with more P2P channels
set -u
exe=tests/cpp/_build/test_mgpu
export MGPU_TEST_TIMING_ONLY=1
for ch in default 16 32; do
  if [ $ch = default ]; then unset NCCL_MIN_P2P_NCHANNELS NCCL_MAX_P2P_NCHANNELS; else export NCCL_MIN_P2P_NCHANNELS=$ch NCCL_MAX_P2P_NCHANNELS=$ch; fi
  echo "== p2p channels: $ch" | tee -a gpurun_out/m_nccl_channels.txt
  MGPU_TEST_SWEEP="1:1:0:0:0:1,2:2:0:0:0:1,2:2:0:0:32:0,4:2:0:0:32:0" timeout 120 $exe 2 1024 1024 1024 0 10 2>&1 | grep -v "NCCL version" | tee -a gpurun_out/m_nccl_channels.txt
done
