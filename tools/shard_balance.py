"""One GPU, every rank's tile shard in turn: the contraction time each of `world` GPUs would see (development helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsdist_b200 as nb
n_ind, n_sites, world = int(os.environ.get("N_IND", 5000)), int(os.environ.get("N_SITES", 500000)), int(os.environ.get("WORLD", 8))
p = nb.Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, call_geno=True, pairwise_del=bool(int(os.environ.get("PDEL", 0))), evol_model=0)
g = nb.NgsDistB200(p)
buf = torch.empty((4096, n_ind, 3), dtype=torch.float64, device="cuda")
for s0 in range(0, n_sites, 4096):
    m = min(4096, n_sites - s0)
    g.synth_raw_device(buf.data_ptr(), 20251018, 0.05, s0, m)
    g.push_sites_device(buf.data_ptr(), s0, m)
g.frontend()
g.distances_raw(None, 0, 1, None)
g.distances_raw(None, 0, 1, None)
full = g.timing().dist_ms
ts = []
for r in range(world):
    g.set_tile_shard(r, world)
    g.distances_raw(None, 0, 1, None)
    g.distances_raw(None, 0, 1, None)
    ts.append(g.timing().dist_ms)
print("%d x %d, world %d: full %.2f ms, ideal %.2f ms, shards %s -> max %.2f ms (%.0f %% of ideal)"
      % (n_ind, n_sites, world, full, full / world, " ".join("%.2f" % t for t in ts), max(ts), 100 * full / world / max(ts)))
