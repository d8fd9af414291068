#!/bin/bash
# 8-GPU session (round 2): the driver-style strong-scaling bench at N = 8 and 4, full C3 x 100 replicates at 8/4/2/1 GPUs,
# full C5 (20 000 x 10 M sites, site-sharded), the host->device ceiling probe, and the sharding checks through the C ABI.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{ nproc; free -g | head -2; nvidia-smi topo -m | head -12; ls /sys/devices/system/node/ | grep node; } > gpurun_out/box8.txt 2>&1
for n in 8 4; do
  timeout 900 $TR --nproc-per-node $n --master-port $((29510+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; echo "bench N=$n rc=$?"
  python - <<P
import json
for l in open('gpurun_out/scale_$n.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=%d value %.4e ms/step %.1f e2e %.4e h2d %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_gbs_per_rank']))
        for k in ('c4_tiles','c5_sites'): print(k, json.dumps(d.get(k))[:600])
P
done
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29520+n)) tools/bench_c3_full.py > gpurun_out/c3full_$n.log 2>&1; echo "c3full N=$n rc=$?"; grep -E "^\{" gpurun_out/c3full_$n.log | cut -c1-900
done
timeout 600 python tools/bench_c3_full.py --check > gpurun_out/c3full_1.log 2>&1; echo "c3full N=1 rc=$?"; grep -E "^\{|property" gpurun_out/c3full_1.log | cut -c1-900
timeout 900 $TR --nproc-per-node $N --master-port 29530 tools/bench_c5.py --check > gpurun_out/c5_$N.log 2>&1; echo "c5 rc=$?"; grep -E "^\{|oracle|rror" gpurun_out/c5_$N.log | cut -c1-1200 | tail -4
for n in 1 4 8; do
  timeout 120 $TR --nproc-per-node $n --master-port $((29540+n)) tools/h2d_probe.py --bind > gpurun_out/h2d_$n.log 2>&1; grep -E "^\{" gpurun_out/h2d_$n.log
done
timeout 120 $TR --nproc-per-node 8 --master-port 29550 tools/h2d_probe.py > gpurun_out/h2d_8_nobind.log 2>&1; grep -E "^\{" gpurun_out/h2d_8_nobind.log
timeout 600 $TR --nproc-per-node $N --master-port 29501 tools/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; echo "multi check rc=$?"; grep -E "sharding|all-gather|CHECK" gpurun_out/multi_check_$N.log | tail -14
timeout 600 python tools/group_check.py $N > gpurun_out/group_check_$N.log 2>&1; echo "group check rc=$?"; tail -9 gpurun_out/group_check_$N.log
