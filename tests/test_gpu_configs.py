"""Parity at the shapes BASELINE.json names, sized so the CPU oracle finishes in seconds:
full-matrix or sub-block comparison where the oracle can, size-independent properties at the full sizes."""
import numpy as np
import pytest

import oracle
from test_gpu_parity import assert_close, nb

pytestmark = pytest.mark.gpu


def gpu_ctx_synth(n_ind, n_sites, data_seed, miss, chunk=8192, **pk):
    """Context fed with the device-side synthetic generator (bit-identical to oracle.synth_raw)."""
    import torch
    p = nb().Params(n_ind=n_ind, n_sites=n_sites, **pk)
    g = nb().NgsDistB200(p)
    buf = torch.empty((min(chunk, n_sites), n_ind, 3), dtype=torch.float64, device="cuda")
    for s0 in range(0, n_sites, chunk):
        m = min(chunk, n_sites - s0)
        g.synth_raw_device(buf.data_ptr(), data_seed, miss, s0, m)
        g.push_sites_device(buf.data_ptr(), s0, m)
    g.frontend()
    return g


def test_device_synth_generator_is_bit_identical_to_oracle():
    import torch
    n_ind, n_sites = 37, 300
    g = nb().NgsDistB200(nb().Params(n_ind=n_ind, n_sites=n_sites, indep_geno=True))
    buf = torch.empty((n_sites, n_ind, 3), dtype=torch.float64, device="cuda")
    g.synth_raw_device(buf.data_ptr(), 20251018, 0.1, 0, n_sites)
    assert np.array_equal(buf.cpu().numpy(), oracle.synth_raw(20251018, 0.1, n_ind, n_sites))
    g.synth_raw_device(buf.data_ptr(), 20251018, 0.1, 1000, n_sites)
    assert np.array_equal(buf.cpu().numpy(), oracle.synth_raw(20251018, 0.1, n_ind, n_sites, site0=1000))
    g.close()


def test_c1_shape_em_default():
    """C1: 24 individuals x 10 000 sites, --probs (default EM path), full matrix vs the oracle."""
    raw = oracle.synth_raw(20251018, 0.0, 24, 10000)
    with nb().NgsDistB200(nb().Params(n_ind=24, n_sites=10000, in_probs=True)) as g:
        g.push_sites(raw)
        r = g.run(want_num=True)[0]
    o = oracle.run_job(raw, indep=False)[0]
    assert_close(r["num"], o["num"], "C1 num")
    assert_close(r["dist"], o["dist"], "C1 dist")


def test_c2_full_size_indep_jc69_blocks():
    """C2: 500 x 100 000, --probs --indep_geno --evol_model 2.  GPU on the full problem; the oracle checks three
    64 x 64 pair blocks (diagonal tile, off-diagonal tile, the ragged last tile) over all 100 000 sites."""
    n_ind, n_sites = 500, 100000
    g = gpu_ctx_synth(n_ind, n_sites, 20251018, 0.0, in_probs=True, indep_geno=True, evol_model=2)
    r = g.distances(want_num=True)
    g.close()
    raw = oracle.synth_raw(20251018, 0.0, n_ind, n_sites)
    P = oracle.frontend(raw)
    del raw
    for (r0, c0) in [(0, 0), (100, 300), (436, 436), (0, 436)]:
        o = oracle.distances_block(P, r0, r0 + 64, c0, c0 + 64, indep=True, evol_model=2)
        sub = np.triu(np.ones((64, 64), bool), 1) if r0 == c0 else np.ones((64, 64), bool)
        got_d = r["dist"][r0:r0 + 64, c0:c0 + 64]
        got_n = r["num"][r0:r0 + 64, c0:c0 + 64]
        assert_close(got_n[sub], o["num"][sub], "C2 num block %d,%d" % (r0, c0))
        assert_close(got_d[sub], o["dist"][sub], "C2 dist block %d,%d" % (r0, c0))
    assert np.array_equal(r["dist"], r["dist"].T) and (np.diag(r["dist"]) == 0).all()


def test_c3_geometry_reduced_sites_bootstrap_pairwise_del():
    """C3 geometry (2 000 individuals, 10 % missing, --pairwise_del, block 1000, seed 12345) at 20 000 sites:
    two bootstrap replicates, oracle on pair blocks; counts bit-exact."""
    n_ind, n_sites, bs = 2000, 20000, 1000
    g = gpu_ctx_synth(n_ind, n_sites, 20251018, 0.10, in_probs=True, indep_geno=True, pairwise_del=True, evol_model=1,
                      n_boot_rep=2, boot_block_size=bs, seed=12345)
    res = g.run(want_num=True, want_cnt=True)
    g.close()
    P = oracle.frontend(oracle.synth_raw(20251018, 0.10, n_ind, n_sites))
    rng = oracle.Taus(12345)
    maps = [None] + [rng.boot_map(n_sites // bs, bs) for _ in range(2)]
    for rep, (r, sm) in enumerate(zip(res, maps)):
        for (r0, c0) in [(0, 0), (64, 1900), (1936, 1936)]:
            o = oracle.distances_block(P, r0, r0 + 64, c0, c0 + 64, indep=True, pairwise_del=True, evol_model=1, site_map=sm)
            sub = np.triu(np.ones((64, 64), bool), 1) if r0 == c0 else np.ones((64, 64), bool)
            assert np.array_equal(r["cnt"][r0:r0 + 64, c0:c0 + 64][sub], o["cnt"][sub]), "cnt rep %d" % rep
            assert_close(r["num"][r0:r0 + 64, c0:c0 + 64][sub], o["num"][sub], "num rep %d" % rep)
            assert_close(r["dist"][r0:r0 + 64, c0:c0 + 64][sub], o["dist"][sub], "dist rep %d" % rep)


def test_c3_properties_at_larger_size():
    """Size-independent properties on 2 000 x 200 000 (C3 geometry): linearity of num / cnt in the block weights,
    all-ones weights == replicate 0, symmetry, zero diagonal, counts bounded by the number of sites."""
    n_ind, n_sites, bs = 2000, 200000, 1000
    g = gpu_ctx_synth(n_ind, n_sites, 7, 0.10, in_probs=True, indep_geno=True, pairwise_del=True, evol_model=0)
    nbk = n_sites // bs
    rng = np.random.RandomState(3)
    c1 = rng.randint(0, 3, nbk).astype(np.uint32)
    c2 = rng.randint(0, 4, nbk).astype(np.uint32)
    r1 = g.distances(c1, bs, want_num=True, want_cnt=True)
    r2 = g.distances(c2, bs, want_num=True, want_cnt=True)
    r3 = g.distances(c1 + c2, bs, want_num=True, want_cnt=True)
    r0 = g.distances(None, 1, want_num=True, want_cnt=True)
    r4 = g.distances(np.ones(nbk, dtype=np.uint32), bs, want_num=True, want_cnt=True)
    g.close()
    assert np.array_equal(r1["cnt"] + r2["cnt"], r3["cnt"])
    assert np.abs(r1["num"] + r2["num"] - r3["num"]).max() <= 1e-12 * np.abs(r3["num"]).max()
    assert np.array_equal(r0["cnt"], r4["cnt"])
    assert np.abs(r0["num"] - r4["num"]).max() <= 1e-12 * np.abs(r0["num"]).max()
    assert np.array_equal(r0["dist"], r0["dist"].T) and (np.diag(r0["dist"]) == 0).all()
    off = ~np.eye(n_ind, dtype=bool)
    assert r0["cnt"][off].max() <= n_sites and r0["cnt"][off].min() > 0.7 * n_sites
    # model-0 distances are num / cnt
    assert np.array_equal(r0["dist"][off], (r0["num"][off] / r0["cnt"][off]))


def test_c4_geometry_called_genotypes_reduced():
    """C4 geometry (--call_geno, 5 % missing) at 1 280 x 6 000: called path is bit-exact with --pairwise_del
    (num, cnt, model-0 distances) and within 1e-9 without it."""
    n_ind, n_sites = 1280, 6000
    raw = oracle.synth_raw(11, 0.05, n_ind, n_sites)
    for pdel in (True, False):
        with nb().NgsDistB200(nb().Params(n_ind=n_ind, n_sites=n_sites, call_geno=True, pairwise_del=pdel, evol_model=0)) as g:
            g.push_sites(raw)
            r = g.run(want_num=True, want_cnt=True)[0]
        P = oracle.frontend(raw, call_geno=True)
        for (r0, c0) in [(0, 0), (600, 1216)]:
            o = oracle.distances_block(P, r0, r0 + 64, c0, c0 + 64, indep=True, pairwise_del=pdel, evol_model=0)
            sub = np.triu(np.ones((64, 64), bool), 1) if r0 == c0 else np.ones((64, 64), bool)
            gn, gc, gd = (r[k][r0:r0 + 64, c0:c0 + 64][sub] for k in ("num", "cnt", "dist"))
            assert np.array_equal(gc, o["cnt"][sub])
            if pdel:
                assert np.array_equal(gn, o["num"][sub]) and np.array_equal(gd, o["dist"][sub])
            else:
                assert_close(gn, o["num"][sub]); assert_close(gd, o["dist"][sub])


def test_c5_geometry_avg_nuc_dist_reduced():
    """C5 geometry (--avg_nuc_dist --indep_geno, no missing data) at 3 000 x 4 000: many tiles (276), oracle blocks."""
    n_ind, n_sites = 3000, 4000
    g = gpu_ctx_synth(n_ind, n_sites, 5, 0.0, in_probs=True, indep_geno=True, avg_nuc_dist=True, evol_model=1)
    r = g.distances(want_num=True)
    g.close()
    P = oracle.frontend(oracle.synth_raw(5, 0.0, n_ind, n_sites))
    for (r0, c0) in [(1536, 1536), (10, 2936), (1400, 1700)]:
        o = oracle.distances_block(P, r0, r0 + 64, c0, c0 + 64, indep=True, evol_model=1, score=oracle.score_matrix(True))
        sub = np.triu(np.ones((64, 64), bool), 1) if r0 == c0 else np.ones((64, 64), bool)
        assert_close(r["num"][r0:r0 + 64, c0:c0 + 64][sub], o["num"][sub])
        assert_close(r["dist"][r0:r0 + 64, c0:c0 + 64][sub], o["dist"][sub])
