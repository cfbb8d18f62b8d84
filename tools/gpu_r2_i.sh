#!/bin/bash
# round-2 call I (1 GPU): in-place column plans (2 / 3 CTAs per SM) against the 2-stage ring on kiss_fftnd 1024^3
set -u
mkdir -p gpurun_out
for v in 0 2 3; do
  KISSFFT_COL_INPLACE=$v timeout 300 python bench.py --workload fftnd1024 --steps 10 --no-configs 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('fftnd1024 inplace=$v', round(d['ms_per_step'], 4), 'ms', round(d['roofline']['frac'], 3), d['kernel_ms'])" | tee -a gpurun_out/i_fftnd.txt
done
for v in 2 3; do
  KISSFFT_COL_INPLACE=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fftnd_permuted or config5_3d_single_gpu_1024 or slab_single_rank" 2>&1 | tail -2 | tee -a gpurun_out/i_fftnd.txt
done
