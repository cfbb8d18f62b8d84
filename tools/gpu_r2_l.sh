#!/bin/bash
# round-2 call L (2 GPUs): final multi-rank check -- C-ABI tests at 2 ranks, host-path tests, the N = 2 bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mgpu_cabi.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "multi_rank or mgpu or slab or host or dropin or reference_api" > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/l_bench_n2.json 2> gpurun_out/l_bench_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/l_bench_n2.json'))
print(d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_pageable']['value'])
for k, v in d['configs'].items(): print(k, v if not isinstance(v, dict) else {a: v.get(a) for a in ('ms', 'parity_ok', 'strong_scaling_efficiency', 'step_vs_bound', 'error', 'exchange', 'sm_partition')})
PY
