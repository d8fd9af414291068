#!/bin/bash
# full ncu capture of k_dist_imma (and optionally the codes front end) at a reduced site count (never a bench number)
mkdir -p gpurun_out
export N_SITES=${N_SITES:-50000} PDEL=${PDEL:-0}
python tools/bench_c4.py > gpurun_out/c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_dist_imma -s 1 -c 1 -f -o gpurun_out/${TAG:-r01}_dist_imma python tools/bench_c4.py > gpurun_out/ncu_c4.log 2>&1
echo "imma capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_frontend_codes -s 2 -c 1 -f -o gpurun_out/${TAG:-r01}_frontend_codes python tools/bench_c4.py > gpurun_out/ncu_c4b.log 2>&1
echo "frontend capture rc=$?"; cat gpurun_out/c4_plain.log
