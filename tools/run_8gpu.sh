#!/bin/bash
# 8-GPU session: sharding-mode check, scaling bench (N = 1, 2, 4, 8 back to back), C5-shaped site-sharded run.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node $N --master-port 29501 tools/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; echo "multi check rc=$?"; grep -E "sharding|CHECK" gpurun_out/multi_check_$N.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    else $TR --nproc-per-node $n --master-port $((29510+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
    echo "bench N=$n rc=$?"; python -c "
import json,sys
for l in open('gpurun_out/scale_$n.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=%d value %.4e ms/step %.3f e2e %.4e' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']))"
  fi
done
$TR --nproc-per-node $N --master-port 29530 tools/bench_c5.py --n-ind 20000 --sites-per-gpu 40000 --check > gpurun_out/c5_$N.log 2>&1; echo "c5 rc=$?"; grep -vE "^\*|OMP_NUM|^$" gpurun_out/c5_$N.log | tail -5
