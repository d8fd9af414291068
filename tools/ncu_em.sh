#!/bin/bash
# full ncu capture of k_dist_em at a reduced site count (never a bench number)
mkdir -p gpurun_out
export N_SITES=${N_SITES:-20000}
python tools/bench_em.py > gpurun_out/em_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_dist_em -s 1 -c 1 -f -o gpurun_out/${TAG:-r01b}_dist_em python tools/bench_em.py > gpurun_out/ncu_em.log 2>&1
echo "em capture rc=$?"; cat gpurun_out/em_plain.log
