#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/bench_em.py > gpurun_out/em_v3b.log 2>&1; tail -1 gpurun_out/em_v3b.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu8.log | tail -20
