#!/bin/bash
# round-2 call F (8 GPUs): pipeline-shape sweeps of kiss_fftnd_mgpu_exec at G = 8 and 4 (peer stores and NCCL), the
# bit-exact reference-order mode at G = 8, and the N = 8 bench line the driver will ask for
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/f_topo.txt 2>&1
exe=tests/cpp/_build/test_mgpu
export MGPU_TEST_TIMING_ONLY=1
for G in 8 4; do
  if [ $G = 8 ]; then b=55; else b=64; fi
  MGPU_TEST_SWEEP="4:2:$b:1,4:2:$b:0,4:2:0:1,4:2:$((b/2)):1,4:2:$((b*3/2)):1,4:2:$((b*2)):1,2:2:$b:1,8:2:$b:1,4:4:$b:1,4:1:$b:1,8:4:$b:1,2:4:$b:1,2:1:$b:1,1:1:$b:1,1:1:0:0,4:2:$b:1" \
    timeout 200 $exe $G 1024 1024 1024 1 10 > gpurun_out/f_sweep_p2p_g$G.jsonl 2> gpurun_out/f_sweep_p2p_g$G.err
  echo "p2p sweep G=$G rc=$?"
  MGPU_TEST_SWEEP="4:2:0:1,2:2:0:1,8:2:0:1,4:4:0:1,4:1:0:1,8:4:0:1,2:1:0:1,1:1:0:1,4:2:0:0" \
    timeout 200 $exe $G 1024 1024 1024 0 10 > gpurun_out/f_sweep_nccl_g$G.jsonl 2> gpurun_out/f_sweep_nccl_g$G.err
  echo "nccl sweep G=$G rc=$?"
done
unset MGPU_TEST_TIMING_ONLY
for fl in 2 3; do timeout 120 tests/cpp/_build/test_mgpu-int16_t 8 512 512 512 $fl 5 2>&1 | grep -v NCCL | tail -8 >> gpurun_out/f_reford_g8.txt; done
timeout 120 $exe 8 512 512 512 1 5 2>&1 | tail -8 >> gpurun_out/f_fast_g8_512.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/f_bench_n8.json 2> gpurun_out/f_bench_n8.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/f_bench_n8.json'))
print(d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_pageable']['value'])
for k, v in d['configs'].items(): print(k, v if not isinstance(v, dict) else {a: v.get(a) for a in ('ms', 'parity_ok', 'strong_scaling_efficiency', 'step_vs_bound', 'error', 'exchange')})
PY
for f in gpurun_out/f_sweep_*.jsonl gpurun_out/f_reford_g8.txt gpurun_out/f_fast_g8_512.txt; do echo "== $f"; cat $f | cut -c1-200; done
