#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_em.py -q -x > gpurun_out/pytest_em.log 2>&1; echo "pytest em rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/pytest_em.log | tail
timeout 120 python tools/bench_em.py > gpurun_out/em_v3.log 2>&1; tail -2 gpurun_out/em_v3.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu5.log | tail -20
