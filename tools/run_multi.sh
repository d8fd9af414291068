#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "multi check rc=$?"; tail -8 gpurun_out/multi_check_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_$N.err; cat gpurun_out/bench_$N.json | cut -c1-700
