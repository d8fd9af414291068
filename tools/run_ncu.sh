#!/bin/bash
# ncu launch list + full captures of the dominant kernels (B200_PROFILING.md recipe). Never a bench number.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_dist_dmma -s 2 -c 1 -f -o gpurun_out/prof_dist $CMD > gpurun_out/ncu_dist.log 2>&1
echo "dist capture rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_frontend -s 2 -c 1 -f -o gpurun_out/prof_front $CMD > gpurun_out/ncu_front.log 2>&1
echo "frontend capture rc=$?"
ls -la gpurun_out
