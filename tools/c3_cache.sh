#!/bin/bash
mkdir -p gpurun_out
python tools/bench_c3.py --reps 6 --block-cache --check 2>&1 | tee gpurun_out/c3_block_cache.log
python tools/bench_c3.py --n-ind 500 --n-sites 100000 --reps 5 --block-cache --em --no-pdel 2>&1 | tee gpurun_out/c2_em_block_cache.log
