"""Parity of the CUDA hot path (through the C ABI) against the CPU oracle and the reference goldens.

Tolerances (north_star): bit-exact for called-genotype and site-count quantities (codes, masks, cnt, and num / model-0
distances on the called + pairwise_del path); 1e-9 relative for floating-point distances.
"""
import numpy as np
import pytest

import oracle
from util import golden_text, load_bin, manifest, parse_flags, read_text_input

pytestmark = pytest.mark.gpu

RTOL = 1e-9
MAN = manifest()


def nb():
    import ngsdist_b200
    return ngsdist_b200


def params_from(kw, n_ind, n_sites, probs=True, in_text=False):
    p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=probs, in_text=in_text)
    p.in_logscale = kw["in_log"]
    p.call_geno = kw["call_geno"]
    p.N_thresh, p.call_thresh = kw["N_thresh"], kw["call_thresh"]
    p.pairwise_del = kw["pairwise_del"]
    p.avg_nuc_dist = kw["avg_nuc_dist"]
    p.indep_geno = kw["indep"]
    p.tot_sites = kw["tot_sites"]
    p.evol_model = kw["evol_model"]
    p.n_boot_rep = kw["n_boot_rep"]
    p.boot_block_size = kw["boot_block_size"]
    p.seed = kw["seed"]
    return p


def assert_close(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    both_nan = np.isnan(got) & np.isnan(want)
    both_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    ok = both_nan | both_inf | (np.abs(got - want) <= RTOL * np.abs(want))
    assert ok.all(), "%s: %d mismatches, worst rel %.3e" % (
        what, (~ok).sum(), np.nanmax(np.abs(got - want)[~ok] / np.maximum(np.abs(want[~ok]), 1e-300)))


def golden_mats(name, n_ind):
    import os, tempfile
    path = os.path.join(tempfile.mkdtemp(), "g.dist")
    with open(path, "w") as fh:
        fh.write(golden_text(name))
    return [m for _, m in oracle.parse_dist(path, n_ind)]


INDEP_CASES = [c for c in MAN["binary"] if parse_flags(c["flags"])[0]["indep"]]


@pytest.mark.parametrize("case", INDEP_CASES, ids=lambda c: c["name"])
def test_golden_cases_indep(case):
    """CUDA path vs the reference's own .dist output (10 printed decimals) and vs the oracle at full precision."""
    kw, _ = parse_flags(case["flags"])
    raw = load_bin(case["input"], case["n_ind"], case["n_sites"])
    with nb().NgsDistB200(params_from(kw, case["n_ind"], case["n_sites"])) as g:
        g.push_sites(raw)
        res = g.run(want_num=True, want_cnt=True)
    ora = oracle.run_job(raw, **kw)
    gold = golden_mats(case["name"], case["n_ind"])
    assert len(res) == len(ora) == len(gold)
    for rep, (r, o, gm) in enumerate(zip(res, ora, gold)):
        assert_close(r["dist"], o["dist"], "%s rep %d dist" % (case["name"], rep))
        assert_close(r["num"], o["num"], "%s rep %d num" % (case["name"], rep))
        assert np.array_equal(r["cnt"], o["cnt"]), "cnt must be bit-exact"
        fin = np.isfinite(gm)
        assert np.allclose(r["dist"][fin], gm[fin], rtol=0, atol=6e-11)
        assert np.array_equal(np.isnan(r["dist"]), np.isnan(gm))
        assert (np.diag(r["dist"]) == 0).all()
        assert np.array_equal(r["dist"], r["dist"].T, equal_nan=True)


@pytest.mark.parametrize("case", [c for c in MAN["text"]], ids=lambda c: c["name"])
def test_golden_cases_text_inputs(case):
    kw, probs = parse_flags(case["flags"])
    if not kw["indep"]:
        pytest.skip("EM path covered in test_gpu_em.py")
    data, blank = read_text_input(case["input"], case["n_ind"], case["n_sites"], probs, want_blank=True)
    res = run_text_case(kw, case, probs, data, blank)
    ora = oracle.run_job(data, kind=1, blank_sites=blank, genotypes=not probs, **kw)
    gold = golden_mats(case["name"], case["n_ind"])
    assert len(res) == len(gold) == len(ora)
    for r, gm, o in zip(res, gold, ora):
        fin = np.isfinite(gm)
        assert np.allclose(r["dist"][fin], gm[fin], rtol=0, atol=6e-11)
        assert np.array_equal(np.isnan(r["dist"]), np.isnan(gm))
        assert np.array_equal(r["cnt"], o["cnt"]), "cnt must be bit-exact"
        assert_close(r["num"], o["num"], case["name"] + " num")


def run_text_case(kw, case, probs, data, blank):
    """What the drop-in reader does with a text file: values as parsed; an empty line becomes the blank-site marker."""
    with nb().NgsDistB200(params_from(kw, case["n_ind"], case["n_sites"], probs=probs, in_text=True)) as g:
        if probs:
            data = data.copy()
            data[blank] = nb().BLANK_SITE
            g.push_sites(data)
        else:
            data = data.astype(np.int8)
            data[blank] = nb().BLANK_SITE_CODE
            g.push_genotypes(data)
        return g.run(want_num=True, want_cnt=True)


@pytest.mark.parametrize("call,thr", [(False, (0, 0)), (True, (0, 0)), (True, (0.35, 0.9))])
def test_frontend_posteriors_and_masks(call, thr):
    raw = load_bin("g7x53.bin", 7, 53)
    raw2 = oracle.synth_raw(11, 0.2, 150, 1000)
    for r in (raw, raw2):
        n_sites, n_ind, _ = r.shape
        P = oracle.frontend(r, call_geno=call, N_thresh=thr[0], call_thresh=thr[1])
        m = oracle.miss_mask(P)
        p = nb().Params(n_ind=n_ind, n_sites=n_sites, indep_geno=True, call_geno=call, N_thresh=thr[0], call_thresh=thr[1],
                        keep_planes=True)   # all three planes kept so that the posteriors can be read back exactly
        with nb().NgsDistB200(p) as g:
            g.push_sites(r)
            g.frontend()
            Pg, mg = g.posteriors()
        assert np.array_equal(mg, 1 - m), "presence masks must be bit-exact"
        if call and thr == (0, 0):
            assert np.array_equal(Pg, P), "called genotypes must be bit-exact"
        else:
            # un-called path: p = x / sum(x) on the device vs exp(log x - logsum) in the reference; the reference itself
            # carries ~|log p| * 2^-52 relative error from "log(sum) + M" (frontend.cu: posterior_fast)
            assert np.allclose(Pg, P, rtol=1e-12, atol=1e-300)
            assert np.array_equal(Pg == 0, P == 0)   # exact zeros stay exact zeros


@pytest.mark.parametrize("n_ind,n_sites,miss", [(130, 1000, 0.1), (300, 4099, 0.05), (257, 515, 0.3)])
@pytest.mark.parametrize("mode", ["indep", "indep_pdel", "call_pdel", "call"])
def test_multi_tile_shapes_with_bootstrap(n_ind, n_sites, miss, mode):
    raw = oracle.synth_raw(20251018 + n_ind, miss, n_ind, n_sites)
    kw = dict(indep=True, pairwise_del="pdel" in mode, call_geno=mode.startswith("call"), evol_model=2,
              n_boot_rep=2, boot_block_size=10 if n_ind != 257 else 1, seed=777)
    p = nb().Params(n_ind=n_ind, n_sites=n_sites, indep_geno=True, pairwise_del=kw["pairwise_del"], call_geno=kw["call_geno"],
                    evol_model=2, n_boot_rep=2, boot_block_size=kw["boot_block_size"], seed=777)
    with nb().NgsDistB200(p) as g:
        # ragged chunked pushes (multiples of 64 sites, last one partial)
        step = 448
        for s0 in range(0, n_sites, step):
            g.push_sites(raw[s0:s0 + step], s0)
        res = g.run(want_num=True, want_cnt=True)
    ora = oracle.run_job(raw, **kw)
    for rep, (r, o) in enumerate(zip(res, ora)):
        assert np.array_equal(r["cnt"], o["cnt"]), "rep %d cnt" % rep
        if mode == "call_pdel":
            # one-hot posteriors: num is an exact multiple of 0.5 in any summation order
            assert np.array_equal(r["num"], o["num"]), "rep %d num must be bit-exact" % rep
            dg = r["num"] / np.maximum(r["cnt"], 1)
        assert_close(r["num"], o["num"], "rep %d num" % rep)
        assert_close(r["dist"], o["dist"], "rep %d dist" % rep)


def test_two_plane_mode_matches_three_plane_mode():
    """indep_geno && !pairwise_del uses the sum-to-one reduction (2 operand planes + c-vector); keep_planes forces the
    plain 3-plane contraction.  Both must agree to rounding, and the read-back third plane is 1 - p0 - p1."""
    raw = oracle.synth_raw(21, 0.1, 300, 2500)
    res = {}
    for keep in (False, True):
        p = nb().Params(n_ind=300, n_sites=2500, indep_geno=True, evol_model=0, keep_planes=keep, n_boot_rep=2, boot_block_size=7, seed=5)
        with nb().NgsDistB200(p) as g:
            g.push_sites(raw)
            res[keep] = g.run(want_num=True, want_cnt=True)
            if not keep:
                P2, _ = g.posteriors()
    o = oracle.run_job(raw, indep=True, evol_model=0, n_boot_rep=2, boot_block_size=7, seed=5)
    for a, b, c in zip(res[False], res[True], o):
        assert np.array_equal(a["cnt"], b["cnt"]) and np.array_equal(a["cnt"], c["cnt"])
        assert_close(a["num"], b["num"], "2 vs 3 planes")
        assert_close(a["num"], c["num"], "2 planes vs oracle")
        assert_close(a["dist"], c["dist"], "2 planes vs oracle")
    P = oracle.frontend(raw)
    assert np.abs(P2 - P).max() < 1e-14


def test_called_pairwise_del_model0_is_bit_exact():
    raw = oracle.synth_raw(5, 0.1, 200, 3000)
    kw = dict(call_geno=True, pairwise_del=True, evol_model=0, avg_nuc_dist=True, n_boot_rep=2, boot_block_size=500, seed=99)
    p = nb().Params(n_ind=200, n_sites=3000, call_geno=True, pairwise_del=True, evol_model=0, avg_nuc_dist=True, n_boot_rep=2,
                    boot_block_size=500, seed=99)
    with nb().NgsDistB200(p) as g:
        g.push_sites(raw)
        res = g.run(want_num=True, want_cnt=True)
    ora = oracle.run_job(raw, **kw)
    for r, o in zip(res, ora):
        assert np.array_equal(r["cnt"], o["cnt"])
        assert np.array_equal(r["num"], o["num"])
        assert np.array_equal(r["dist"], o["dist"])   # correctly rounded division of exact operands


def test_tot_sites_and_models():
    raw = oracle.synth_raw(8, 0.0, 40, 777)
    for model in (0, 1, 2):
        p = nb().Params(n_ind=40, n_sites=777, indep_geno=True, evol_model=model, tot_sites=5000)
        with nb().NgsDistB200(p) as g:
            g.push_sites(raw)
            r = g.run()[0]
        o = oracle.run_job(raw, indep=True, evol_model=model, tot_sites=5000)[0]
        assert_close(r["dist"], o["dist"], "model %d" % model)


def test_error_behaviour_matches_reference():
    N = nb()
    with pytest.raises(N.NgsDistError, match="missing data threshold"):
        N.NgsDistB200(N.Params(n_ind=4, n_sites=10, N_thresh=0.9, call_thresh=0.5))
    with pytest.raises(N.NgsDistError, match="not yet supported"):
        N.NgsDistB200(N.Params(n_ind=4, n_sites=10, indep_geno=True, evol_model=3))
    with pytest.raises(N.NgsDistError, match="tot_sites"):
        N.NgsDistB200(N.Params(n_ind=4, n_sites=10, indep_geno=True, tot_sites=5, pairwise_del=True))
    raw = np.full((10, 4, 3), 0.2)
    raw[3, 1, 0] = np.nan
    with N.NgsDistB200(N.Params(n_ind=4, n_sites=10, indep_geno=True)) as g:
        g.push_sites(raw)
        with pytest.raises(N.NgsDistError, match="NaN found"):
            g.frontend()
    with N.NgsDistB200(N.Params(n_ind=4, n_sites=10, in_probs=False)) as g:
        codes = np.zeros((10, 4), dtype=np.int8)
        codes[2, 2] = 3
        g.push_genotypes(codes)
        with pytest.raises(N.NgsDistError, match="Genotypes must be coded"):
            g.frontend()
    with N.NgsDistB200(N.Params(n_ind=4, n_sites=200, indep_geno=True)) as g:
        g.push_sites(np.full((64, 4, 3), 0.2), 0)
        with pytest.raises(N.NgsDistError, match="incomplete"):
            g.distances()


def test_empty_overlap_gives_reference_nan():
    # two individuals that never share a site: cnt == 0 -> 0/0 (ngsDist.cpp:376)
    raw = oracle.synth_raw(1, 0.0, 3, 64)
    raw[:32, 0, :] = 1 / 3
    raw[32:, 1, :] = 1 / 3
    for model in (0, 1):
        p = nb().Params(n_ind=3, n_sites=64, indep_geno=True, pairwise_del=True, evol_model=model)
        with nb().NgsDistB200(p) as g:
            g.push_sites(raw)
            r = g.run(want_cnt=True)[0]
        o = oracle.run_job(raw, indep=True, pairwise_del=True, evol_model=model)[0]
        assert r["cnt"][0, 1] == 0 and np.isnan(r["dist"][0, 1]) and np.isnan(o["dist"][0, 1])
        assert_close(r["dist"], o["dist"])
