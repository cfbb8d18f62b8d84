#!/bin/bash
# round-2 call K (1 GPU): host-pointer pipeline on pinned buffers -- all chunks enqueued up front by one thread (async) against
# the threaded lanes, ramped chunk sizes; correctness of the host paths
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host" 2>&1 | tail -2
for cfg in "0 0 32 4" "1 0 32 4" "1 1 32 4" "1 1 16 4" "1 1 8 4" "1 1 16 6" "1 1 8 8" "1 0 16 4" "1 1 32 6"; do
  set -- $cfg
  KISSFFT_HOST_ASYNC=$1 KISSFFT_CHUNK_RAMP=$2 KISSFFT_CHUNK_MIB=$3 KISSFFT_HOST_LANES=$4 timeout 300 python bench.py --no-configs --steps 5 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('async=$1 ramp=$2 chunk_mib=$3 lanes=$4', 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'], 2), 'ms')" | tee -a gpurun_out/k_e2e_async.txt
done
