"""C3-shaped run (BASELINE.json configs[2]): synthetic GL n_ind x n_sites, 10 % missing, --pairwise_del --indep_geno,
bootstrap replicates with block_size 1000, one B200 (replicates shard across GPUs with no collective, so per-GPU
numbers are the whole story).  Prints per-matrix kernel times, nominal pair-sites/s and executed DMMA TFLOP/s, and
checks the size-independent properties used at full scale (linearity of num/cnt in the block weights, symmetry)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ngsdist_b200 as nb

ap = argparse.ArgumentParser()
ap.add_argument("--n-ind", type=int, default=2000)
ap.add_argument("--n-sites", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--block", type=int, default=1000)
ap.add_argument("--miss", type=float, default=0.10)
ap.add_argument("--chunk", type=int, default=8192)
ap.add_argument("--no-pdel", action="store_true")
ap.add_argument("--check", action="store_true")
ap.add_argument("--block-cache", action="store_true", help="bootstrap block cache (per-block partials contracted once); default: the graded per-replicate weighted contraction")
ap.add_argument("--em", action="store_true", help="default --probs path (per pair-site EM) instead of --indep_geno")
args = ap.parse_args()

n_ind, n_sites = args.n_ind, args.n_sites
p = nb.Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=not args.em, pairwise_del=not args.no_pdel, evol_model=1,
              n_boot_rep=args.reps, boot_block_size=args.block, seed=12345, no_block_cache=not args.block_cache)
t0 = time.time()
g = nb.NgsDistB200(p)
buf = torch.empty((args.chunk, n_ind, 3), dtype=torch.float64, device="cuda")
fe_ms = 0.0
for s0 in range(0, n_sites, args.chunk):
    m = min(args.chunk, n_sites - s0)
    g.synth_raw_device(buf.data_ptr(), 20251018, args.miss, s0, m)
    g.push_sites_device(buf.data_ptr(), s0, m)
    fe_ms += g.timing().frontend_ms
g.frontend()
torch.cuda.synchronize()
print("setup+frontend wall %.2f s; front-end kernels %.1f ms for %.3g ind-sites -> %.1f GB/s algorithmic (72 B/ind-site)" %
      (time.time() - t0, fe_ms, n_ind * n_sites, n_ind * n_sites * 72 / fe_ms / 1e6))
pairs = n_ind * (n_ind - 1) // 2
out = np.empty((n_ind, n_ind))
rows = []
peak = nb.probe_fp64_tflops(0)
for rep in range(args.reps + 1):
    if rep == 0:
        counts, bs, n_eff = None, 1, n_sites
    else:
        counts, bs = g.next_boot_counts()
        n_eff = len(counts) * bs
    for it in range(1 if args.block_cache else 2):   # second pass = warm (with the cache the first weighted call builds it)
        t1 = time.time()
        r = g.distances(counts, bs, out=out)
        wall = time.time() - t1
    t = g.timing()
    nominal = pairs * n_eff
    row = dict(rep=rep, block_cache=t.block_cache, count_ms=t.count_ms, dist_ms=t.dist_ms, epilogue_ms=t.epilogue_ms, total_ms=t.total_ms, wall_ms=wall * 1e3,
               active_sites=t.active_sites, nominal_pair_sites_per_s=nominal / (t.total_ms * 1e-3),
               dmma_tflops=t.dist_dmma * 512 / (max(t.dist_ms, 1e-6) * 1e-3) * 1e-12, dmma_frac=t.dist_dmma * 512 / (max(t.dist_ms, 1e-6) * 1e-3) * 1e-12 / peak,
               useful_tflops=6.0 * nominal / (max(t.dist_ms, 1e-6) * 1e-3) * 1e-12)
    rows.append(row)
    print(json.dumps(row))
print("fp64 dmma peak (probe): %.2f TFLOP/s" % peak)

if args.check:
    # linearity in the block weights: D(c1) + D(c2) == D(c1 + c2) for num (1e-12) and cnt (exact); symmetry; diagonal
    nbk = n_sites // args.block
    rng = np.random.RandomState(1)
    c1 = rng.randint(0, 3, nbk).astype(np.uint32); c2 = rng.randint(0, 3, nbk).astype(np.uint32)
    r1 = g.distances(c1, args.block, want_num=True, want_cnt=True)
    r2 = g.distances(c2, args.block, want_num=True, want_cnt=True)
    r3 = g.distances(c1 + c2, args.block, want_num=True, want_cnt=True)
    assert np.array_equal(r1["cnt"] + r2["cnt"], r3["cnt"]), "cnt not linear"
    rel = np.abs(r1["num"] + r2["num"] - r3["num"]).max() / np.abs(r3["num"]).max()
    assert rel < 1e-12, rel
    assert np.array_equal(r3["dist"], r3["dist"].T) and (np.diag(r3["dist"]) == 0).all()
    ones = np.ones(nbk, dtype=np.uint32)
    r0 = g.distances(None, 1, want_num=True, want_cnt=True)
    r4 = g.distances(ones, args.block, want_num=True, want_cnt=True)
    if n_sites % args.block == 0:
        assert np.array_equal(r0["cnt"], r4["cnt"])
        assert np.abs(r0["num"] - r4["num"]).max() / np.abs(r0["num"]).max() < 1e-12
    print("property checks ok (linearity rel err %.2e)" % rel)
