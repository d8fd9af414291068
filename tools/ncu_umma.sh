#!/bin/bash
mkdir -p gpurun_out
export N_SITES=${N_SITES:-50000} PDEL=0
timeout 120 python tools/bench_c4.py > gpurun_out/umma_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dist_umma -s 1 -c 1 -f -o gpurun_out/${TAG:-r01}_dist_umma python tools/bench_c4.py > gpurun_out/ncu_umma.log 2>&1
echo "umma capture rc=$?"; cat gpurun_out/umma_plain.log
