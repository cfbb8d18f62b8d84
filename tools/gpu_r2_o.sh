#!/bin/bash
# round-2 call O (1 GPU): last sanity check of the final tree -- smoke(), the single-rank C-ABI mgpu tests (3-D and 2-D), the default bench line
set -u
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_mgpu_cabi.py -m gpu -x -q -k "single_rank_through or 2d_through" 2>&1 | tail -1
timeout 150 python bench.py --steps 10 > gpurun_out/o_bench_n1.json 2> gpurun_out/o_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/o_bench_n1.json'))
print(d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_pageable']['value'], d['cpu_baseline'].get('value'))
for k, v in d['configs'].items(): print(k, v.get('ms'), v.get('frac'), v.get('parity_ok'), v.get('error'), v.get('int_gop_per_s_5NlogN'))
PY
