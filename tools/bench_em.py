"""EM-path (default --probs, no --indep_geno) throughput at the C2 shape; development helper."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsdist_b200 as nb
n_ind, n_sites = int(os.environ.get("N_IND", 500)), int(os.environ.get("N_SITES", 100000))
p = nb.Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=False, evol_model=2)
g = nb.NgsDistB200(p)
raw = torch.empty((n_sites, n_ind, 3), dtype=torch.float64, device="cuda")
g.synth_raw_device(raw.data_ptr(), 20251018, 0.0, 0, n_sites)
g.push_sites_device(raw.data_ptr(), 0, n_sites)
out = torch.empty((n_ind, n_ind), dtype=torch.float64).pin_memory()
ts = []
for _ in range(4):
    g.distances_raw(None, 0, 1, out.data_ptr())
    ts.append(g.timing().dist_ms)
ps = n_ind * (n_ind - 1) // 2 * n_sites
print("EM path %dx%d: dist_em %.2f ms (min %.2f) -> %.3e pair-sites/s" % (n_ind, n_sites, statistics.median(ts), min(ts), ps / (min(ts) * 1e-3)))
