import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ngsdist_b200 as nb, oracle
def run(tag, mut, n_ind=150, n_sites=700, pushes=((0, 700),)):
    raw = oracle.synth_raw(21, 0.05, n_ind, n_sites)
    mut(raw)
    p = nb.Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=True, pairwise_del=False, evol_model=0)
    with nb.NgsDistB200(p) as g:
        for a, b in pushes:
            g.push_sites(raw[a:b], a)
        r = g.run(want_num=True)[0]
        nh = g.deferred_stats()
    o = oracle.run_job(raw, indep=True, evol_model=0)[0]
    err = np.abs(r["num"] - o["num"])
    iu = np.triu_indices(n_ind, 1)
    bad = np.argwhere(np.triu(err > 1e-9 * np.abs(o["num"]), 1))
    print(tag, "host-evaluated", nh, "max abs err %.3e" % err[iu].max(), "bad pairs", len(bad), "first", bad[:6].tolist())
    if len(bad):
        i, j = bad[0]
        print("   got %.6f want %.6f" % (r["num"][i, j], o["num"][i, j]))
rng = np.random.RandomState(3)
run("none", lambda r: None)
def one(r): r[5, 3] = 0.0
run("one triple (site 5, ind 3)", one)
def site(r): r[5, :] = 0.0
run("whole site 5", site)
def ind(r): r[:, 7] = 0.0
run("whole individual 7", ind)
def ind140(r): r[:, 140] = 0.0
run("whole individual 140", ind140)
def rnd(r):
    for _ in range(400): r[rng.randint(700), rng.randint(150)] = 0.0
run("400 random", rnd)
run("whole site, 3 pushes", site, pushes=((0, 320), (320, 700), (320, 700)))
run("one triple small", one, n_ind=20, n_sites=128)
