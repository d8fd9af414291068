#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29541 tools/bench_c3_full.py --n-sites 200000 --reps 10 > gpurun_out/c3full_2.log 2>&1; echo "c3full rc=$?"; grep -E "^\{" gpurun_out/c3full_2.log | cut -c1-800
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
timeout 900 $TR --nproc-per-node 2 --master-port 29544 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<'P'
import json
for line in open('gpurun_out/bench_n2.json'):
    if line.startswith('{'):
        d=json.loads(line); json.dump(d, open('gpurun_out/r02_scale_n2.json','w'), indent=1)
        for k in ('value','ms_per_step','e2e','c4_tiles','c5_sites'): print(k, json.dumps(d.get(k))[:700])
P
