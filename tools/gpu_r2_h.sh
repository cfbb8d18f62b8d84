#!/bin/bash
# round-2 call H (8 GPUs): green-context partition sweep at G = 8 and 4, timeline of the best, N = 8 bench with the best
set -u
mkdir -p gpurun_out
exe=tests/cpp/_build/test_mgpu
export MGPU_TEST_TIMING_ONLY=1
for G in 8 4; do
  if [ $G = 8 ]; then links="0 24 32 40 48 56 64 72"; else links="0 40 48 56 64 72 80"; fi
  sw=""
  for l in $links; do for cp in 4:2 2:2 8:2 4:1 4:4 2:1; do sw="$sw,$cp:-1:0:0:0:$l"; done; done
  MGPU_TEST_SWEEP="${sw#,}" timeout 300 $exe $G 1024 1024 1024 1 10 2>&1 | grep -v NCCL > gpurun_out/h_sweep_p2p_g$G.jsonl
  echo "sweep G=$G rc=$?"
  best=$(python - <<PY
import json
rows=[json.loads(l) for l in open('gpurun_out/h_sweep_p2p_g$G.jsonl') if l.startswith('{')]
b=min(rows,key=lambda r:r['ms'])
print("%d:%d:-1:0:0:1:%d"%(b['chunks'],b['pchunks'],b['link_sms']))
PY
)
  echo "best G=$G: $best"
  MGPU_TEST_SWEEP="$best,$best" timeout 120 $exe $G 1024 1024 1024 1 10 2>&1 | grep -v NCCL > gpurun_out/h_best_timeline_g$G.txt
  if [ $G = 8 ]; then best8=$best; fi
done
unset MGPU_TEST_TIMING_ONLY
IFS=: read c p b pr rs tr l <<< "$best8"
echo "bench with CHUNKS=$c PCHUNKS=$p LINK_SMS=$l"
KISSFFT_MGPU_CHUNKS=$c KISSFFT_MGPU_PCHUNKS=$p KISSFFT_MGPU_LINK_SMS=$l timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/h_bench_n8.json 2> gpurun_out/h_bench_n8.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/h_bench_n8.json'))
print(d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_pageable']['value'])
for k, v in d['configs'].items(): print(k, v if not isinstance(v, dict) else {a: v.get(a) for a in ('ms', 'parity_ok', 'strong_scaling_efficiency', 'step_vs_bound', 'error', 'exchange', 'sm_partition', 'pipeline_chunks')})
PY
for G in 8 4; do python - <<PY
import json
rows=[json.loads(l) for l in open('gpurun_out/h_sweep_p2p_g$G.jsonl') if l.startswith('{')]
rows.sort(key=lambda r:r['ms'])
for r in rows[:12]: print($G, r['chunks'], r['pchunks'], r['link_sms'], r['rest_sms'], r['ms'])
PY
cat gpurun_out/h_best_timeline_g$G.txt | tail -24
done
