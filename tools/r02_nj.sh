#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_nj.py tests/test_cli.py -q -m gpu -k "nj or tree" 2>&1 | tail -5
python - <<'P'
import time, numpy as np, sys
sys.path.insert(0, '.')
import ngsdist_b200 as nb
for n in (2000, 5000):
    rng = np.random.RandomState(n)
    X = rng.rand(n, 8)
    D = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(axis=2))
    with nb.NgsDistB200(nb.Params(n_ind=n, n_sites=64, in_probs=True, indep_geno=True)) as g:
        g.nj_tree(D)
        t0 = time.perf_counter(); s = g.nj_tree(D); dt = time.perf_counter() - t0
    print("NJ n=%d: %.1f ms, newick %d bytes" % (n, dt * 1e3, len(s)))
P
