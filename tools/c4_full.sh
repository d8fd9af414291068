#!/bin/bash
# full-size C4 (5 000 x 5 000 000 called genotypes) on ONE B200: 6.25 GB of 2-bit codes
mkdir -p gpurun_out
N_SITES=5000000 PDEL=0 python tools/bench_c4.py 2>&1 | tee gpurun_out/c4_full.log
N_SITES=5000000 PDEL=1 python tools/bench_c4.py 2>&1 | tee -a gpurun_out/c4_full.log
PACKED=1 N_SITES=5000000 PDEL=0 python tools/bench_c4.py 2>&1 | tail -1 | tee -a gpurun_out/c4_full.log
