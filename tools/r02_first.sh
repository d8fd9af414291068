#!/bin/bash
# round-2 first GPU call: box facts, GPU tests (incl. 2-GPU checks), C3 bench at N=1 and N=2
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; nvidia-smi topo -m; ls /sys/devices/system/node/ | grep node; } > gpurun_out/box.txt 2>&1
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"
tail -c 3000 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
tail -c 3000 gpurun_out/bench_n2.err
cut -c1-1500 gpurun_out/bench_n1.json; cut -c1-1500 gpurun_out/bench_n2.json
