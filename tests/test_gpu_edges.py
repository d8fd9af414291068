"""Edge shapes through the C ABI against the oracle: one individual (no pairs), two individuals, a single site, site counts
around the 8 / 12 / 64-site chunk and word boundaries, an individual that is missing everywhere (cnt = 0 -> nan like the
reference), and all three contractions (FP64 GEMM, per pair-site EM, int8 called path)."""
import numpy as np
import pytest

import oracle
from test_gpu_parity import RTOL, nb

pytestmark = pytest.mark.gpu


def assert_close(got, want, what="", n_sites=1, cond=None):
    """1e-9 relative, with an absolute floor of 1e-15 per site: the two-plane (sum-to-one) evaluation of a site term has
    ~1e-16 ABSOLUTE error, which is a large RELATIVE error only for pairs whose whole sum is ~1e-9 or less (a few sites of
    near-identical sharp posteriors) -- five orders of magnitude below the 10 decimals the writer prints."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    both_nan = np.isnan(got) & np.isnan(want)
    both_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    with np.errstate(invalid="ignore"):
        floor = 1e-15 * n_sites * (1.0 if cond is None else cond)
        ok = both_nan | both_inf | (np.abs(got - want) <= RTOL * np.abs(want) + floor)
    assert ok.all(), "%s: %d mismatches" % (what, (~ok).sum())

MODES = {
    "indep": dict(indep_geno=True),
    "indep_pdel": dict(indep_geno=True, pairwise_del=True),
    "em": dict(indep_geno=False),
    "em_pdel": dict(indep_geno=False, pairwise_del=True),
    "called": dict(call_geno=True),
    "called_pdel": dict(call_geno=True, pairwise_del=True),
}


def ora_kw(pk):
    return dict(indep=pk.get("indep_geno", False) or pk.get("call_geno", False), call_geno=pk.get("call_geno", False),
                pairwise_del=pk.get("pairwise_del", False))


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("n_ind,n_sites", [(1, 10), (2, 1), (2, 7), (3, 8), (5, 12), (9, 63), (9, 64), (9, 65), (130, 3), (129, 129)])
def test_edge_shapes(mode, n_ind, n_sites):
    pk = MODES[mode]
    raw = oracle.synth_raw(1000 + n_ind * 7 + n_sites, 0.2, n_ind, n_sites)
    for model in (0, 2):
        p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, evol_model=model, **pk)
        with nb().NgsDistB200(p) as g:
            g.push_sites(raw)
            r = g.run(want_num=True, want_cnt=True)[0]
        o = oracle.run_job(raw, evol_model=model, **ora_kw(pk))[0]
        assert r["dist"].shape == (n_ind, n_ind)
        assert np.array_equal(r["cnt"], o["cnt"])
        assert_close(r["num"], o["num"], "%s %dx%d num" % (mode, n_ind, n_sites), n_sites)
        # JC69 is -3/4 log(1 - 4p/3): its condition number 1 / |1 - 4p/3| blows up for p -> 3/4, which pairs with one to
        # three sites do hit; the absolute floor is scaled by it (and by 1 / cnt for p = num / cnt)
        with np.errstate(divide="ignore", invalid="ignore"):
            pdist = o["num"] / np.maximum(o["cnt"].astype(np.float64), 1.0)
            cond = 1.0 / np.maximum(o["cnt"].astype(np.float64), 1.0) * (1.0 if model == 0 else 1.0 / np.maximum(np.abs(1.0 - 4.0 * pdist / 3.0), 1e-300))
        assert_close(r["dist"], o["dist"], "%s %dx%d dist m%d" % (mode, n_ind, n_sites, model), n_sites, cond)
        assert (np.diag(r["dist"]) == 0).all()


@pytest.mark.parametrize("mode", ["indep_pdel", "em_pdel", "called_pdel"])
def test_individual_missing_everywhere_gives_nan_like_the_reference(mode):
    pk = MODES[mode]
    n_ind, n_sites = 6, 200
    raw = oracle.synth_raw(5, 0.1, n_ind, n_sites)
    raw[:, 2, :] = 1.0 / 3.0                      # individual 2 has no data at all: every pair with it has cnt = 0
    for model in (0, 1):
        p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, evol_model=model, **pk)
        with nb().NgsDistB200(p) as g:
            g.push_sites(raw)
            r = g.run(want_cnt=True)[0]
        o = oracle.run_job(raw, evol_model=model, **ora_kw(pk))[0]
        assert np.array_equal(r["cnt"], o["cnt"]) and (r["cnt"][2, [0, 1, 3, 4, 5]] == 0).all()
        assert np.array_equal(np.isnan(r["dist"]), np.isnan(o["dist"])) and np.isnan(r["dist"][2, 0])
        assert_close(r["dist"], o["dist"])
