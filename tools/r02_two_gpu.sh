#!/bin/bash
# 2-GPU validation of everything the 8-GPU session will run (small shapes)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_multi.log | tail -8
timeout 300 $TR --nproc-per-node 2 --master-port 29541 tools/bench_c3_full.py --n-sites 200000 --reps 10 --check > gpurun_out/c3full_2.log 2>&1; echo "c3full rc=$?"; grep -E "^\{|property|Error|error" gpurun_out/c3full_2.log | cut -c1-700 | tail -4
timeout 300 python tools/bench_c3_full.py --n-sites 200000 --reps 10 > gpurun_out/c3full_1.log 2>&1; echo "c3full1 rc=$?"; grep -E "^\{" gpurun_out/c3full_1.log | cut -c1-700
timeout 600 $TR --nproc-per-node 2 --master-port 29542 tools/bench_c5.py --n-sites 400000 --check > gpurun_out/c5_2.log 2>&1; echo "c5 rc=$?"; grep -E "^\{|oracle|Error|error" gpurun_out/c5_2.log | cut -c1-900 | tail -4
timeout 120 $TR --nproc-per-node 2 --master-port 29543 tools/h2d_probe.py --bind > gpurun_out/h2d_2.log 2>&1; grep -E "^\{" gpurun_out/h2d_2.log
timeout 900 $TR --nproc-per-node 2 --master-port 29544 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
grep -E "Error|error|Traceback" gpurun_out/bench_n2.err | head -5
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/bench_n2.json'))
    for k in ('value','ms_per_step','step_share_rank0_ms','collective','e2e','c4_tiles','c5_sites'): print(k, json.dumps(d.get(k))[:700])
except Exception as e: print("no bench json", e)
P
