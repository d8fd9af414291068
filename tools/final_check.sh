#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_final_n1.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_final_n1.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','gpu_launches','clocks','step_share_rank0_ms','e2e','cpu_baseline','c2','called_path','bootstrap_block_cache'): print(k, json.dumps(d.get(k))[:1100])
print('roofline', d['roofline']['frac'], d['roofline']['traffic'])
P
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/pytest_final.log | tail -8
