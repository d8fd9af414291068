"""C4 end to end from host memory on ONE GPU with the push and the contraction overlapped, using nothing but the public
entry points: the sites are split over K contexts (the site-sharding of SURVEY §8e, all on one device); a second host
thread pushes shard k + 1 over PCIe (ngsd_push_packed_genotypes) while the main thread contracts shard k
(ngsd_distances with out == NULL: raw sums stay on the device); the raw sums are added on the device and ngsd_finish
applies the epilogue.  Checked against the single-context result (each shard divides its integer sum by S = 18 before the
addition, so without --pairwise_del the agreement is to rounding, 1e-13, not to the bit)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ngsdist_b200 as nb
from ngsdist_b200 import multi

n_ind, n_sites = int(os.environ.get("N_IND", 5000)), int(os.environ.get("N_SITES", 5000000))
K = int(os.environ.get("SHARDS", 4))
pdel = bool(int(os.environ.get("PDEL", 0)))
stride = (n_ind + 3) // 4
host = torch.empty((n_sites, stride), dtype=torch.uint8).pin_memory()
blk = torch.from_numpy(np.random.RandomState(1).randint(0, 256, size=(min(n_sites, 65536), stride), dtype=np.uint8))
for s0 in range(0, n_sites, blk.shape[0]):
    m = min(blk.shape[0], n_sites - s0)
    host[s0:s0 + m] = torch.roll(blk[:m], s0 // blk.shape[0], dims=1)
hnp = host.numpy()
kw = dict(n_ind=n_ind, in_probs=False, indep_geno=True, pairwise_del=pdel, evol_model=0)
out = torch.empty((n_ind, n_ind), dtype=torch.float64).pin_memory()
pairs = n_ind * (n_ind - 1) // 2

# one context, push then contract
g = nb.NgsDistB200(nb.Params(n_sites=n_sites, **kw))
best1 = None
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g.push_packed_genotypes(hnp); g.frontend(); g.distances_raw(None, 0, 1, out.data_ptr())
    dt = time.perf_counter() - t0
    best1 = dt if it else None
single = out.numpy().copy()
g.close()

shards = multi.site_shards(n_sites, 64, K)
ctxs = [nb.NgsDistB200(nb.Params(n_sites=s1 - s0, **kw)) for s0, s1 in shards]
best = None
for it in range(3):
    pushed = [threading.Event() for _ in range(K)]

    def pusher():
        for k, (s0, s1) in enumerate(shards):
            ctxs[k].push_packed_genotypes(hnp[s0:s1])
            ctxs[k].frontend()
            pushed[k].set()

    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = threading.Thread(target=pusher); th.start()
    for k in range(K):
        pushed[k].wait()
        ctxs[k].partial_sums(None, 1)                    # contraction of shard k while the pusher copies shard k + 1
    th.join()
    _, n0, c0 = ctxs[0].device_results()
    num0 = multi.device_tensor(n0, (n_ind, n_ind), "<f8"); cnt0 = multi.device_tensor(c0, (n_ind, n_ind), "<i8")
    for k in range(1, K):
        _, nk, ck = ctxs[k].device_results()
        num0 += multi.device_tensor(nk, (n_ind, n_ind), "<f8"); cnt0 += multi.device_tensor(ck, (n_ind, n_ind), "<i8")
    torch.cuda.synchronize()
    ctxs[0].finish(out=out.numpy())
    dt = time.perf_counter() - t0
    if it:
        best = dt if best is None else min(best, dt)
same = np.allclose(out.numpy(), single, rtol=1e-13, atol=0, equal_nan=True)
print("packed host input %dx%d pdel=%d: one context push-then-contract %.1f ms; %d site shards, push overlapped with contraction %.1f ms -> %.3e pair-sites/s; agrees to 1e-13: %s"
      % (n_ind, n_sites, pdel, best1 * 1e3, K, best * 1e3, pairs * n_sites / best, same))
sys.exit(0 if same else 1)
