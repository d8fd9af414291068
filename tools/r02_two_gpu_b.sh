#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29542 tools/bench_c5.py --n-sites 800000 --check > gpurun_out/c5_2.log 2>&1; echo "c5 rc=$?"; grep -E "^\{|oracle|rror" gpurun_out/c5_2.log | cut -c1-1200 | tail -4
timeout 300 $TR --nproc-per-node 2 --master-port 29541 tools/bench_c3_full.py --n-sites 200000 --reps 10 > gpurun_out/c3full_2.log 2>&1; echo "c3full rc=$?"; grep -E "^\{" gpurun_out/c3full_2.log | cut -c1-700
timeout 300 python tools/group_check.py 2 > gpurun_out/group_check_2.log 2>&1; echo "group rc=$?"; tail -3 gpurun_out/group_check_2.log
timeout 900 $TR --nproc-per-node 2 --master-port 29544 bench.py --gpus 2 --steps 1 --warmup 3 --no-e2e > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<'P'
import json
for line in open('gpurun_out/bench_n2.json'):
    if line.startswith('{'):
        d=json.loads(line)
        for k in ('value','ms_per_step','c4_tiles','c5_sites'): print(k, json.dumps(d.get(k))[:900])
P
