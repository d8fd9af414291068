"""Development sweep: dist-kernel time of the C2 shape under scheduler knobs (NGSD_TUNE / NGSD_NODIAG)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsdist_b200 as nb

n_ind, n_sites = int(os.environ.get("N_IND", 500)), int(os.environ.get("N_SITES", 100000))
p = nb.Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=True, evol_model=2)
g = nb.NgsDistB200(p)
raw = torch.empty((n_sites, n_ind, 3), dtype=torch.float64, device="cuda")
g.synth_raw_device(raw.data_ptr(), 1, 0.0, 0, n_sites)
g.push_sites_device(raw.data_ptr(), 0, n_sites)
print("frontend_ms", g.timing().frontend_ms)
out = torch.empty((n_ind, n_ind), dtype=torch.float64).pin_memory()
configs = [("default", None, None)]
SHORT = bool(os.environ.get("NGSD_SWEEP_SHORT"))
for upc in (() if SHORT else (4, 8, 16)):
    for ratio in (1, 4, 8):
        for frac in (0.7, 0.84, 0.92):
            configs.append(("tune %g,%g,%g" % (upc, ratio, frac), "%g,%g,%g" % (upc, ratio, frac), None))
configs.append(("nodiag default", None, "1"))
configs.append(("nodiag 8,1,0.84", "8,1,0.84", "1"))
for name, tune, nodiag in configs:
    for k, v in (("NGSD_TUNE", tune), ("NGSD_NODIAG", nodiag)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    ts, es = [], []
    for _ in range(6):
        g.distances_raw(None, 0, 1, out.data_ptr())
        t = g.timing()
        ts.append(t.dist_ms); es.append(t.epilogue_ms)
    print("%-24s dist %.3f ms (min %.3f)  epi %.3f  dmma %.2f TF" % (name, statistics.median(ts), min(ts), statistics.median(es),
          t.dist_dmma * 512 / (min(ts) * 1e-3) * 1e-12))
