"""Bootstrap block cache: resampling is by whole blocks (ngsDist.cpp:416-437), so every replicate is a weighted sum of
per-block partial sums that are contracted once.  The cached path must give the reference's matrices (oracle, 1e-9) and
agree with the direct per-replicate weighted contraction (ngsd_cfg.reserved bit 2) to summation-order noise; counts stay
bit-exact.  Covers the three contractions that use it (3-plane and 2-plane DMMA, per pair-site EM)."""
import numpy as np
import pytest

import oracle
from test_gpu_parity import assert_close, nb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,pk,bs", [
    ("dmma3_pdel", dict(indep_geno=True, pairwise_del=True, evol_model=2), 40),
    ("dmma2", dict(indep_geno=True, evol_model=1), 24),
    ("em", dict(indep_geno=False, evol_model=0), 16),
    ("em_pdel", dict(indep_geno=False, pairwise_del=True, evol_model=0), 32),
])
def test_block_cache_matches_oracle_and_direct_path(name, pk, bs):
    n_ind, n_sites, nrep = 150, 2011, 4          # 2011 sites: the bootstrap truncates to whole blocks (ngsDist.cpp:236)
    raw = oracle.synth_raw(123, 0.1, n_ind, n_sites)
    res = {}
    for cache in (True, False):
        p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, n_boot_rep=nrep, boot_block_size=bs, seed=7,
                        no_block_cache=not cache, **pk)
        with nb().NgsDistB200(p) as g:
            g.push_sites(raw)
            g.frontend()
            out, flags = [], []
            for rep in range(nrep + 1):
                if rep == 0:
                    out.append(g.distances(want_num=True, want_cnt=True))
                else:
                    c, b = g.next_boot_counts()
                    out.append(g.distances(c, b, want_num=True, want_cnt=True))
                flags.append(g.timing().block_cache)
            res[cache] = out
        assert flags == ([0, 1, 2, 2, 2] if cache else [0] * 5), flags
    ora = oracle.run_job(raw, indep=pk["indep_geno"], pairwise_del=pk.get("pairwise_del", False), evol_model=pk["evol_model"],
                         n_boot_rep=nrep, boot_block_size=bs, seed=7)
    for rep in range(nrep + 1):
        a, b, o = res[True][rep], res[False][rep], ora[rep]
        assert np.array_equal(a["cnt"], o["cnt"]) and np.array_equal(b["cnt"], o["cnt"])
        assert_close(a["num"], o["num"], "%s rep %d cached vs oracle" % (name, rep))
        assert_close(a["dist"], o["dist"], "%s rep %d cached vs oracle" % (name, rep))
        scale = np.abs(b["num"]).max()
        assert np.abs(a["num"] - b["num"]).max() <= 1e-12 * scale, "%s rep %d cached vs direct" % (name, rep)


def test_block_cache_is_rebuilt_after_new_data_and_new_geometry():
    n_ind, n_sites = 90, 1280
    raw1 = oracle.synth_raw(1, 0.05, n_ind, n_sites)
    raw2 = oracle.synth_raw(2, 0.05, n_ind, n_sites)
    p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=True, pairwise_del=True, evol_model=0)
    counts64 = np.array([2, 0, 1] * 6 + [1, 3], dtype=np.uint32)      # 20 blocks of 64
    counts128 = np.array([1, 0, 2, 1, 0, 3, 1, 1, 0, 1], dtype=np.uint32)
    with nb().NgsDistB200(p) as g:
        g.push_sites(raw1)
        a1 = g.distances(counts64, 64, want_num=True); f1 = g.timing().block_cache
        a2 = g.distances(counts128, 128, want_num=True); f2 = g.timing().block_cache      # new geometry -> rebuilt
        a3 = g.distances(counts64, 64, want_num=True); f3 = g.timing().block_cache
        g.push_sites(raw2)                                                                  # new data -> rebuilt
        b1 = g.distances(counts64, 64, want_num=True); f4 = g.timing().block_cache
    assert (f1, f2, f3, f4) == (1, 1, 1, 1)
    assert np.array_equal(a1["num"], a3["num"])
    P1, P2 = oracle.frontend(raw1), oracle.frontend(raw2)
    for got, P, c, bs in ((a1, P1, counts64, 64), (a2, P1, counts128, 128), (b1, P2, counts64, 64)):
        sm = np.concatenate([np.arange(b * bs, (b + 1) * bs) for b in range(len(c)) for _ in range(int(c[b]))])
        o = oracle.distances(P, indep=True, pairwise_del=True, evol_model=0, site_map=sm)
        assert_close(got["num"], o["num"])


@pytest.mark.parametrize("pdel", [True, False])
def test_block_cache_integer_path_is_exact(pdel):
    """Called genotypes: per-block int32 partials (blocks = whole 64-site words) weighted in int64 -- identical to the
    direct weighted contraction, cnt and (with --pairwise_del) num bit-exact against the oracle."""
    n_ind, n_sites, bs, nrep = 200, 2000, 128, 3
    raw = oracle.synth_raw(9, 0.1, n_ind, n_sites)
    res = {}
    for cache in (True, False):
        p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, call_geno=True, pairwise_del=pdel, evol_model=0,
                        n_boot_rep=nrep, boot_block_size=bs, seed=3, no_block_cache=not cache)
        with nb().NgsDistB200(p) as g:
            g.push_sites(raw)
            g.frontend()
            out, flags = [], []
            for rep in range(nrep + 1):
                if rep == 0:
                    out.append(g.distances(want_num=True, want_cnt=True))
                else:
                    c, b = g.next_boot_counts()
                    out.append(g.distances(c, b, want_num=True, want_cnt=True))
                flags.append(g.timing().block_cache)
            res[cache] = out
        assert flags == ([0, 1, 2, 2] if cache else [0] * 4), flags
    ora = oracle.run_job(raw, indep=True, call_geno=True, pairwise_del=pdel, evol_model=0, n_boot_rep=nrep, boot_block_size=bs, seed=3)
    for rep in range(nrep + 1):
        a, b, o = res[True][rep], res[False][rep], ora[rep]
        assert np.array_equal(a["cnt"], b["cnt"]) and np.array_equal(a["cnt"], o["cnt"])
        assert np.array_equal(a["num"], b["num"]) and np.array_equal(a["dist"], b["dist"], equal_nan=True)
        if pdel:
            assert np.array_equal(a["num"], o["num"])
        else:
            assert_close(a["num"], o["num"])


def test_block_cache_first_call_weighted_with_large_multiplicities():
    """No replicate-0 call before the first weighted one, multiplicities far above one int8 byte: the cached path weights
    the per-block partials in int64, so it needs no weight layers and stays exact."""
    n_ind, n_sites, bs = 130, 1280, 128
    raw = oracle.synth_raw(3, 0.1, n_ind, n_sites)
    c_big = np.array([300, 0, 1, 127, 128, 0, 5, 0, 2, 254], dtype=np.uint32)
    got = {}
    for cache in (True, False):
        p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, call_geno=True, pairwise_del=True, evol_model=0,
                        no_block_cache=not cache)
        with nb().NgsDistB200(p) as g:
            g.push_sites(raw)
            got[cache] = g.distances(c_big, bs, want_num=True, want_cnt=True)
            assert g.timing().block_cache == (1 if cache else 0)
    for k in ("cnt", "num", "dist"):
        assert np.array_equal(got[True][k], got[False][k], equal_nan=True), k
    sm = np.concatenate([np.arange(b * bs, (b + 1) * bs) for b in range(len(c_big)) for _ in range(int(c_big[b]))])
    o = oracle.distances(oracle.frontend(raw, call_geno=True), indep=True, pairwise_del=True, evol_model=0, site_map=sm)
    assert np.array_equal(got[True]["cnt"], o["cnt"]) and np.array_equal(got[True]["num"], o["num"])
