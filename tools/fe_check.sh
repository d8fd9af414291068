#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --core-only 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('step %.3f ms  dist %.3f ms  value %.3e  e2e %.3e'%(d['ms_per_step'], r['kernel_ms'], d['value'], d['e2e']['value']))"
python - <<'PY'
import torch, ngsdist_b200 as nb
for planes_kw, label in ((dict(indep_geno=True), "2-plane"), (dict(indep_geno=True, pairwise_del=True), "3-plane")):
    p = nb.Params(n_ind=500, n_sites=100000, in_probs=True, evol_model=2, **planes_kw)
    g = nb.NgsDistB200(p)
    raw = torch.empty((100000, 500, 3), dtype=torch.float64, device="cuda")
    g.synth_raw_device(raw.data_ptr(), 1, 0.0, 0, 100000)
    ts = []
    for _ in range(5):
        g.push_sites_device(raw.data_ptr(), 0, 100000); ts.append(g.timing().frontend_ms)
    b = 5e7 * (64 if label == "2-plane" else 72)
    print("frontend %s: %.3f ms = %.2f TB/s" % (label, min(ts), b / min(ts) * 1e-9))
    g.close()
PY
