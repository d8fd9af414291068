#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_int_path.py tests/test_gpu_knife.py -q > gpurun_out/repro.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/repro.log | tail -5
export N_SITES=16000
python tools/bench_em.py > gpurun_out/em_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_dist_em3 -s 1 -c 1 -o gpurun_out/r02_em3 python tools/bench_em.py > gpurun_out/ncu_em.log 2>&1
tail -2 gpurun_out/em_plain.log; tail -3 gpurun_out/ncu_em.log; ls -la gpurun_out/*.ncu-rep
