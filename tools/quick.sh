#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
NGSD_SWEEP_SHORT=1 python tools/sweep.py 2>&1 | tee gpurun_out/sweep.log | head -8
