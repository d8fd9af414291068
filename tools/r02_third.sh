#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu3.log | tail -40
