#!/bin/bash
# Round-2 profiles (B200_PROFILING.md recipe): launch list of the bench command (C3) + one --set full capture per kernel
# that changed.  Every ncu run is preceded by the same command exiting 0 without ncu; numbers printed under ncu are
# never bench values.  One GPU.
mkdir -p gpurun_out
R=r02
NCU="ncu --set full --clock-control none --import-source on"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --core-only --no-e2e --reps 2"
$BENCH > gpurun_out/${R}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv $BENCH > /dev/null 2>&1
echo "launch list rc=$?"
# launches per step: 1 front end + per replicate (umma count, cnt_reduce, dmma, zero_diag, epilogue); skip the 3 warm-up steps
$NCU -k regex:k_dist_dmma -s 6 -c 1 -f -o gpurun_out/${R}_dist_dmma_weighted_c3 $BENCH > gpurun_out/${R}_ncu1.log 2>&1; echo "dmma rc=$?"
$NCU -k regex:k_dist_umma -s 6 -c 1 -f -o gpurun_out/${R}_count_umma_c3 $BENCH > gpurun_out/${R}_ncu2.log 2>&1; echo "count umma rc=$?"
$NCU -k 'regex:k_frontend' -s 3 -c 1 -f -o gpurun_out/${R}_frontend_c3 $BENCH > gpurun_out/${R}_ncu3.log 2>&1; echo "frontend rc=$?"
EM="env N_SITES=20000 python tools/bench_em.py"
$EM > gpurun_out/${R}_plain_em.log 2>&1 &&
$NCU -k regex:k_dist_em3 -s 1 -c 1 -f -o gpurun_out/${R}_dist_em3 $EM > gpurun_out/${R}_ncu5.log 2>&1; echo "em rc=$?"
C4="env N_SITES=50000 PDEL=1 python tools/bench_c4.py"
$C4 > gpurun_out/${R}_plain_c4.log 2>&1 &&
$NCU -k regex:k_dist_umma -s 2 -c 2 -f -o gpurun_out/${R}_dist_umma_c4 $C4 > gpurun_out/${R}_ncu8.log 2>&1; echo "umma rc=$?"
ls -la gpurun_out | grep $R
