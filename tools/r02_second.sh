#!/bin/bash
# round-2 second GPU call (1 GPU): full GPU test suite, C3 bench, POPC-vs-tensor-core count A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -25 gpurun_out/pytest_gpu2.log
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --core-only --no-cpu-baseline > gpurun_out/bench2_n1.json 2> gpurun_out/bench2_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench2_n1.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench2_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','step_share_rank0_ms')}); print(d['roofline']['frac'], d.get('e2e',{}).get('value'))
P
NGSD_COUNT_POPC=1 timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --core-only --no-cpu-baseline --no-e2e > gpurun_out/bench2_popc.json 2>/dev/null
python - <<'P'
import json
d=json.load(open('gpurun_out/bench2_popc.json'))
print("POPC count:", {k:d[k] for k in ('value','ms_per_step','step_share_rank0_ms')})
P
