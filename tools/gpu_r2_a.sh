#!/bin/bash
# round-2 validation call: full GPU test suite, the default bench line, the launch list, and ncu --set full captures of the
# dominant kernel of every BASELINE config.   gpurun --timeout 2400 -- 'bash tools/gpu_r2_a.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/a_pytest.log
timeout 900 python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/a_bench_n1.json
KISSFFT_FFTND_INLAYOUT=0 timeout 300 python bench.py --workload fftnd1024 --steps 10 --no-configs > gpurun_out/a_fftnd1024_sweeps.json 2> /dev/null
cut -c1-400 gpurun_out/a_fftnd1024_sweeps.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 1 --no-configs > gpurun_out/a_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_ -o gpurun_out/a_ncu_all -f python tools/prof_launch.py > gpurun_out/a_ncu_all.log 2>&1
echo "ncu all rc=$?"
PROF_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_ -o gpurun_out/a_ncu_fftnd1024 -f python tools/prof_launch.py fftnd1024 > gpurun_out/a_ncu_fftnd.log 2>&1
echo "ncu fftnd rc=$?"
KISSFFT_FFTND_INLAYOUT=0 PROF_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_ -o gpurun_out/a_ncu_fftnd1024_sweeps -f python tools/prof_launch.py fftnd1024 > gpurun_out/a_ncu_fftnd_sweeps.log 2>&1
echo "ncu fftnd sweeps rc=$?"
ls -la gpurun_out | head -40
