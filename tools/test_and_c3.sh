#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python tools/bench_c3.py --check 2>&1 | tee gpurun_out/c3.log | cut -c1-400
