"""Transport tiers of ngsd_push_sites_packed: the same front end from narrower host values.
Fixed point is EXACT for decimal inputs (the double strtod gives "0.333340" is (double) 333340 / 1e6); float32 rounds
the inputs to 24 bits and its effect on the distances is measured here against the 1e-9 budget of north_star."""
import numpy as np
import pytest

import oracle
from test_gpu_parity import nb

pytestmark = pytest.mark.gpu


def decimal_posteriors(n_ind, n_sites, seed=9):
    raw = oracle.synth_raw(seed, 0.1, n_ind, n_sites)
    raw /= raw.sum(axis=2, keepdims=True)
    q = np.rint(raw * 1e6).astype(np.int64)
    vals = np.array([float("%.6f" % (k * 1e-6)) for k in range(q.max() + 1)])   # what a text reader gets
    return q, vals[q]


@pytest.mark.parametrize("mode", ["indep", "indep_pdel", "em", "call"])
@pytest.mark.parametrize("fmt", ["u32", "u20x3"])
def test_fixed_point_transport_is_bit_identical(mode, fmt):
    n_ind, n_sites = 140, 900
    q, raw = decimal_posteriors(n_ind, n_sites)
    kw = {"indep": dict(indep_geno=True), "indep_pdel": dict(indep_geno=True, pairwise_del=True), "em": dict(indep_geno=False),
          "call": dict(call_geno=True)}[mode]
    p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, in_text=True, evol_model=2, n_boot_rep=1, boot_block_size=50, **kw)
    with nb().NgsDistB200(p) as a:
        a.push_sites(raw[:512], 0)
        a.push_sites(raw[512:], 512)
        ra = a.run(want_num=True, want_cnt=True)
    with nb().NgsDistB200(p) as b:
        for s0, s1 in ((0, 512), (512, n_sites)):
            b.push_sites_fixed(q[s0:s1].astype(np.uint32) if fmt == "u32" else nb().pack_u20x3(q[s0:s1]), 1e6, s0)
        rb = b.run(want_num=True, want_cnt=True)
    for x, y in zip(ra, rb):
        for k in ("dist", "num", "cnt"):
            assert np.array_equal(x[k], y[k], equal_nan=True), (mode, fmt, k)


def test_float32_transport_error_budget():
    n_ind, n_sites = 200, 20000
    raw = oracle.synth_raw(4, 0.0, n_ind, n_sites)
    p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=True, evol_model=2)
    with nb().NgsDistB200(p) as a:
        a.push_sites(raw)
        da = a.run()[0]["dist"]
    with nb().NgsDistB200(p) as b:
        b.push_sites_f32(raw.astype(np.float32))
        db = b.run()[0]["dist"]
    off = ~np.eye(n_ind, dtype=bool)
    rel = np.abs(da - db)[off] / np.abs(da)[off]
    print("float32 transport: max rel error of JC69 distances %.2e, median %.2e" % (rel.max(), np.median(rel)))
    assert rel.max() < 5e-8          # each input carries 2^-24 relative error; the sum over 20 000 sites averages it down
    assert rel.max() > 1e-12         # ... but it is NOT inside the 1e-9 budget: a documented, opt-in tier


def test_packed_transport_argument_errors():
    p = nb().Params(n_ind=4, n_sites=64, in_probs=True, indep_geno=True, in_logscale=True)
    with nb().NgsDistB200(p) as g:
        with pytest.raises(nb().NgsDistError):
            g.push_sites_fixed(np.zeros((64, 4, 3), dtype=np.uint32), 1e6)       # fixed point + log scale
    p = nb().Params(n_ind=4, n_sites=64, in_probs=True, indep_geno=True)
    with nb().NgsDistB200(p) as g:
        with pytest.raises(nb().NgsDistError):
            g.push_sites_fixed(np.zeros((64, 4, 3), dtype=np.uint32), 0.0)       # denom
