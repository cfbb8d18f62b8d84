#!/bin/bash
# One GPU-box command that confirms and times the opt-in paths listed in profiles/r01/NEXT.md.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/validate_next.sh'
# Results land in gpurun_out/next_*.  Every step runs under its own timeout.
set -u
mkdir -p gpurun_out
KISSFFT_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k opt_in > gpurun_out/next_pytest.log 2>&1
tail -3 gpurun_out/next_pytest.log
for t in float double; do
    timeout 120 python tools/sizes_bench.py $t 16384 65536 262144 1048576 > gpurun_out/next_long_${t}_multipass.jsonl 2>&1
    KISSFFT_FOURSTEP=1 timeout 120 python tools/sizes_bench.py $t 16384 65536 262144 1048576 > gpurun_out/next_long_${t}_fourstep.jsonl 2>&1
done
timeout 120 python bench.py --workload fftnd1024 --steps 10 --warmup 3 > gpurun_out/next_fftnd1024_sweeps.json 2> /dev/null
KISSFFT_FFTND_INLAYOUT=1 timeout 120 python bench.py --workload fftnd1024 --steps 10 --warmup 3 > gpurun_out/next_fftnd1024_inlayout.json 2> /dev/null
for b in tools/_build/tune_next_*; do
    [ -x "$b" ] && timeout 120 "$b" 32768 > gpurun_out/$(basename $b).jsonl 2> /dev/null
done
grep -h '"ms"' gpurun_out/next_long_*.jsonl | cut -c1-160
cut -c1-300 gpurun_out/next_fftnd1024_*.json
