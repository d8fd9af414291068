"""In-process multi-GPU contexts (ngsd_cfg.n_gpus > 1) against a single-GPU run: ONE process, one host thread per GPU
inside libngsdist_b200.so, NCCL below the C ABI.  Needs >= 2 visible GPUs:  python tools/group_check.py [n_gpus]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ngsdist_b200 as nb
import oracle

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else min(torch.cuda.device_count(), 8)
assert n_gpus >= 2, "needs at least 2 GPUs"
n_ind, n_sites, bs, nrep, seed = 700, 20000, 100, 5, 12345
raw = oracle.synth_raw(99, 0.1, n_ind, n_sites)


def relerr(a, b):
    return float(np.nanmax(np.abs(a - b) / np.abs(b + np.eye(n_ind))))


def run(params, **ctx_kw):
    g = nb.NgsDistB200(params, **ctx_kw)
    for s0 in range(0, n_sites, 4096):                      # chunked host pushes, as the reader does
        g.push_sites(raw[s0:s0 + 4096], s0)
    first = g.distances(want_num=True, want_cnt=True)
    reps = g.run_batched()
    g.close()
    return first, reps


passes = (("FP64 contraction, 3 planes", dict(indep_geno=True, pairwise_del=True)),
          ("FP64 contraction, 2 planes", dict(indep_geno=True, pairwise_del=False)),
          ("int8 contraction (called genotypes)", dict(call_geno=True, pairwise_del=True)),
          ("per pair-site EM", dict(indep_geno=False, pairwise_del=True)))
for label, kw in passes:
    p = nb.Params(n_ind=n_ind, n_sites=n_sites, evol_model=2, n_boot_rep=nrep, boot_block_size=bs, seed=seed, **kw)
    ref_first, ref = run(p)
    for mode, name in ((nb.api.SHARD_REPLICATED, "REPLICATED"), (nb.api.SHARD_SITES, "SITES")):
        first, got = run(p, n_gpus=n_gpus, shard=mode)
        assert np.array_equal(first["cnt"], ref_first["cnt"]), (label, name, "cnt")
        worst = max(relerr(first["dist"], ref_first["dist"]), relerr(first["num"], ref_first["num"]))
        for r, (a, b) in enumerate(zip(got, ref)):
            worst = max(worst, relerr(a, b))
        tol = 1e-9 if not kw.get("indep_geno", True) else 1e-12
        assert worst < tol, (label, name, worst)
        print("%-40s %-10s %d GPUs: %d matrices, cnt exact, max rel diff to one GPU %.1e" % (label, name, n_gpus, len(got), worst), flush=True)
print("GROUP_CHECK_OK n_gpus=%d" % n_gpus)
