#!/bin/bash
# round-2 call E (2 GPUs): multi-rank C-ABI tests incl. the reference-order (bit-exact Q15) mode, and the N = 2 bench line
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/e_topo.txt 2>&1
timeout 900 python -m pytest tests/test_mgpu_cabi.py tests/test_gpu_parity.py -m gpu -x -q -k "multi_rank or mgpu or slab or fftnd_permuted or fourstep" > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/e_bench_n2.json'))
print(d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_pageable']['value'])
for k, v in d['configs'].items(): print(k, v if not isinstance(v, dict) else {a: v.get(a) for a in ('ms', 'parity_ok', 'strong_scaling_efficiency', 'step_vs_bound', 'error', 'exchange')})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/e_bench_ref_n2.json 2> /dev/null
cut -c1-400 gpurun_out/e_bench_ref_n2.json
for t in float int16_t; do
  exe=tests/cpp/_build/test_mgpu; [ $t = float ] || exe=tests/cpp/_build/test_mgpu-$t
  for fl in 2 3; do timeout 120 $exe 2 512 512 512 $fl 5 2>&1 | tail -2 >> gpurun_out/e_reford.txt; done
done
cat gpurun_out/e_reford.txt
du -sh gpurun_out
