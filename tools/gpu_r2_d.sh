#!/bin/bash
# round-2 call D: full GPU suite on the fixed Q31 arithmetic, default bench line, ncu of the new Q31 / permuted fftnd kernels,
# column-ring candidates for 512 / 256
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/d_pytest.log
timeout 900 python bench.py > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/d_bench_n1.json'))
print(d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_pageable']['value'], d['cpu_baseline'].get('value'))
for k, v in d['configs'].items(): print(k, v.get('ms'), round(v.get('frac', 0), 3), v.get('parity_ok'), v.get('error'))
PY
for w in q31_2048 fftnd1024; do
  PROF_REPS=1 timeout 600 ncu --set full --clock-control none -k regex:kf_ -o /tmp/ncu_$w -f python tools/prof_launch.py $w > gpurun_out/d_ncu_$w.log 2>&1
  python tools/ncu_summary.py metrics /tmp/ncu_$w.ncu-rep > gpurun_out/d_ncu_${w}_metrics.txt 2>&1
done
TUNE_NCOLS=512 timeout 200 tools/_build/tune_r2e_f32_col512 262144 5 > gpurun_out/d_tune_col512.jsonl 2>&1
python tools/tune_report.py gpurun_out/d_tune_col512.jsonl | head -8
TUNE_NCOLS=256 timeout 200 tools/_build/tune_r2e_f32_col256 262144 5 > gpurun_out/d_tune_col256.jsonl 2>&1
python tools/tune_report.py gpurun_out/d_tune_col256.jsonl | head -8
du -sh gpurun_out
