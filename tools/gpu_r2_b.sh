#!/bin/bash
# round-2 call B: bench line, ncu captures summarised ON THE BOX (the .ncu-rep files are too big to travel), stride experiment
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err
echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 2 --warmup 1 --no-configs > gpurun_out/b_launches_bench.log 2>&1
R=/tmp/ncu_all.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_ -o /tmp/ncu_all -f python tools/prof_launch.py > gpurun_out/b_ncu_all.log 2>&1
echo "ncu all rc=$?"
python tools/ncu_summary.py metrics $R > gpurun_out/b_ncu_all_metrics.txt 2>&1
for w in fftnd1024; do
  PROF_REPS=1 timeout 600 ncu --set full --clock-control none -k regex:kf_ -o /tmp/ncu_$w -f python tools/prof_launch.py $w > gpurun_out/b_ncu_$w.log 2>&1
  python tools/ncu_summary.py metrics /tmp/ncu_$w.ncu-rep > gpurun_out/b_ncu_${w}_metrics.txt 2>&1
done
# hottest SASS of single kernels (one capture each, with source)
for w in q15_2048 q31_2048 c2c1155 z2z1155; do
  PROF_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:kf_ -c 1 -o /tmp/ncu_$w -f python tools/prof_launch.py $w > /dev/null 2>&1
  python tools/ncu_summary.py hot /tmp/ncu_$w.ncu-rep 60 > gpurun_out/b_ncu_${w}_hot.txt 2>&1
done
# stride experiment: is the 8 MiB-stride column pass slow because of address translation or of channel camping?
for nc in 1048576 1049600 1048592 ; do
  echo "== TUNE_NCOLS=$nc (one plane)" >> gpurun_out/b_stride.txt
  KISSFFT_RING_ANY_STRIDE=1 TUNE_NCOLS=$nc timeout 120 tools/_build/tune_r2c_f32_col1024 $nc 5 >> gpurun_out/b_stride.txt 2>&1
  KISSFFT_RING_ANY_STRIDE=1 TUNE_NCOLS=$nc timeout 120 tools/_build/tune_r2c_f32_colcol1024 $nc 5 >> gpurun_out/b_stride.txt 2>&1
done
for nc in 1024 1040 ; do
  b=$((nc * 1008))
  echo "== TUNE_NCOLS=$nc batch $b" >> gpurun_out/b_stride.txt
  TUNE_NCOLS=$nc timeout 120 tools/_build/tune_r2c_f32_col1024 $b 5 >> gpurun_out/b_stride.txt 2>&1
  TUNE_NCOLS=$nc timeout 120 tools/_build/tune_r2c_f32_colcol1024 $b 5 >> gpurun_out/b_stride.txt 2>&1
done
ls -la gpurun_out; du -sh gpurun_out
