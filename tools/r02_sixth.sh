#!/bin/bash
mkdir -p gpurun_out
# reproduce the order-dependent failure, then look for uninitialised device reads
timeout 600 python -m pytest tests/test_gpu_int_path.py tests/test_gpu_knife.py -q > gpurun_out/repro.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/repro.log | tail -5
timeout 900 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_gpu_int_path.py tests/test_gpu_knife.py -q -k "genotype_input_many or zero" > gpurun_out/initcheck.log 2>&1
grep -E "Uninitialized|at .*\(|passed|failed|ERROR SUMMARY" gpurun_out/initcheck.log | head -40
