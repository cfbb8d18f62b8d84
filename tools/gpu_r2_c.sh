#!/bin/bash
# round-2 call C: integer-instruction microbenchmark, Q31 plan variants on the mad.wide arithmetic, the full GPU suite
# (new: reference-order mgpu mode on one rank, permuted 3-D passes), kiss_fftnd timings per strategy
set -u
mkdir -p gpurun_out
timeout 120 tools/_build/intbench > gpurun_out/c_intbench.jsonl 2>&1; cat gpurun_out/c_intbench.jsonl
timeout 300 tools/_build/tune_r2d_q31_2048 65536 10 > gpurun_out/c_tune_q31.jsonl 2>&1
python tools/tune_report.py gpurun_out/c_tune_q31.jsonl | head -14
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/c_pytest.log
for w in fftnd1024 fftnd512 fftnd256; do
  for p in 0 1; do
    KISSFFT_FFTND_PERMUTE=$p timeout 300 python bench.py --workload $w --steps 10 --no-configs 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$w permute=$p', round(d['ms_per_step'], 4), 'ms', round(d['roofline']['frac'], 3))" | tee -a gpurun_out/c_fftnd.txt
  done
done
timeout 300 python bench.py --workload q31_2048 --steps 10 --no-configs > gpurun_out/c_bench_q31.json 2> /dev/null; cut -c1-300 gpurun_out/c_bench_q31.json
du -sh gpurun_out
