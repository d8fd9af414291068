"""Default --probs path (no --indep_geno): per pair-site em2 (emOptim2.cpp:69-135, call site ngsDist.cpp:340-353).

The CUDA kernel uses the closed form of SURVEY App. C; the oracle iterates exactly like the reference.  A stopping
index that flips by one at a knife edge would change one site's term by <= 1e-3, i.e. show up far above 1e-9 at these
sizes, so agreement to 1e-9 on every pair is also the T-agreement check.
"""
import numpy as np
import pytest

import oracle
from util import load_bin, manifest, parse_flags, read_text_input
from test_gpu_parity import assert_close, golden_mats, nb, params_from, run_text_case

pytestmark = pytest.mark.gpu
MAN = manifest()
EM_CASES = [c for c in MAN["binary"] if not parse_flags(c["flags"])[0]["indep"]]


@pytest.mark.parametrize("case", EM_CASES, ids=lambda c: c["name"])
def test_golden_cases_em(case):
    kw, _ = parse_flags(case["flags"])
    raw = load_bin(case["input"], case["n_ind"], case["n_sites"])
    with nb().NgsDistB200(params_from(kw, case["n_ind"], case["n_sites"])) as g:
        g.push_sites(raw)
        res = g.run(want_num=True, want_cnt=True)
    ora = oracle.run_job(raw, **kw)
    gold = golden_mats(case["name"], case["n_ind"])
    assert len(res) == len(ora) == len(gold)
    for rep, (r, o, gm) in enumerate(zip(res, ora, gold)):
        assert np.array_equal(r["cnt"], o["cnt"])
        assert_close(r["num"], o["num"], "%s rep %d num" % (case["name"], rep))
        assert_close(r["dist"], o["dist"], "%s rep %d dist" % (case["name"], rep))
        fin = np.isfinite(gm)
        assert np.allclose(r["dist"][fin], gm[fin], rtol=0, atol=6e-11)


@pytest.mark.parametrize("case", [c for c in MAN["text"] if not parse_flags(c["flags"])[0]["indep"]], ids=lambda c: c["name"])
def test_text_probs_em_goldens(case):
    """Text posteriors through the per pair-site EM: plain, on the miss_data boundary under --pairwise_del, and with empty
    lines (every pair NaN without --pairwise_del, the sites skipped with it)."""
    kw, probs = parse_flags(case["flags"])
    data, blank = read_text_input(case["input"], case["n_ind"], case["n_sites"], probs, want_blank=True)
    res = run_text_case(kw, case, probs, data, blank)
    ora = oracle.run_job(data, kind=1, blank_sites=blank, **kw)
    gold = golden_mats(case["name"], case["n_ind"])
    for r, gm, o in zip(res, gold, ora):
        fin = np.isfinite(gm)
        assert np.allclose(r["dist"][fin], gm[fin], rtol=0, atol=6e-11)
        assert np.array_equal(np.isnan(r["dist"]), np.isnan(gm))
        if fin.any():
            assert np.array_equal(r["cnt"], o["cnt"])


@pytest.mark.parametrize("alpha", [0.1, 0.5, 2.0])
@pytest.mark.parametrize("pdel", [False, True])
def test_em_random_dirichlet(alpha, pdel):
    """Soft posteriors of different sharpness (mean EM iterations 6..16 in SURVEY App. C) + missing data + bootstrap."""
    rng = np.random.RandomState(int(alpha * 10) + pdel)
    n_ind, n_sites = 70, 1500
    raw = rng.gamma(alpha, size=(n_sites, n_ind, 3)) + 1e-300
    miss = rng.rand(n_sites, n_ind) < 0.1
    raw[miss] = 1.0 / 3.0
    kw = dict(indep=False, pairwise_del=pdel, evol_model=0, n_boot_rep=1, boot_block_size=50, seed=31)
    p = nb().Params(n_ind=n_ind, n_sites=n_sites, indep_geno=False, pairwise_del=pdel, evol_model=0, n_boot_rep=1, boot_block_size=50, seed=31)
    with nb().NgsDistB200(p) as g:
        g.push_sites(raw)
        res = g.run(want_num=True, want_cnt=True)
    ora = oracle.run_job(raw, **kw)
    for rep, (r, o) in enumerate(zip(res, ora)):
        assert np.array_equal(r["cnt"], o["cnt"])
        assert_close(r["num"], o["num"], "alpha %g rep %d" % (alpha, rep))
        assert_close(r["dist"], o["dist"], "alpha %g rep %d" % (alpha, rep))
