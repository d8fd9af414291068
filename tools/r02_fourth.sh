#!/bin/bash
mkdir -p gpurun_out
python tools/debug_zero.py > gpurun_out/debug_zero.log 2>&1; cat gpurun_out/debug_zero.log | tail -20
timeout 300 python -m pytest tests/test_gpu_em.py -q -x > gpurun_out/pytest_em.log 2>&1; echo "pytest em rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/pytest_em.log | tail
timeout 120 python tools/bench_em.py > gpurun_out/em_v2.log 2>&1; tail -3 gpurun_out/em_v2.log
NGSD_EM_V1=1 timeout 120 python tools/bench_em.py > gpurun_out/em_v1.log 2>&1; tail -3 gpurun_out/em_v1.log
timeout 1200 python -m pytest tests/test_gpu_knife.py tests/test_gpu_transport.py tests/test_gpu_edges.py tests/test_gpu_configs.py tests/test_gpu_block_cache.py tests/test_gpu_parity.py -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
grep -E "passed|failed|FAILED|rc=|float32" gpurun_out/pytest_gpu4.log | tail -30
