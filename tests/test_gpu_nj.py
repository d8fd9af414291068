"""N4: neighbour-joining trees from the resident distance matrices (ngsd_nj_tree) against the CPU
restatement of the same published rules (oracle/nj_oracle.py).  The reference has no tree code: parity unpinned by it."""
import numpy as np
import pytest

import oracle
from oracle import nj_oracle
from test_gpu_parity import nb

pytestmark = pytest.mark.gpu


def additive_tree_matrix(n, seed):
    """Distances of a random tree: neighbour joining must recover exactly this tree (up to rounding)."""
    rng = np.random.RandomState(seed)
    parent = [-1] + [int(rng.randint(0, k)) for k in range(1, 2 * n - 2)]
    # simple construction: star-decomposition by random pairwise path sums over a random tree on 2n-2 nodes
    L = rng.uniform(0.01, 0.2, size=2 * n - 2)
    nodes = 2 * n - 2
    D = np.zeros((nodes, nodes))
    for a in range(1, nodes):
        p = parent[a]
        D[a, :a] = D[p, :a] + L[a]
        D[:a, a] = D[a, :a]
    leaves = rng.choice(nodes, n, replace=False)
    return D[np.ix_(leaves, leaves)]


@pytest.mark.parametrize("n", [3, 4, 9, 150, 300])
def test_nj_matches_the_cpu_restatement(n):
    rng = np.random.RandomState(n)
    X = rng.rand(n, 6)
    D = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(axis=2)) + 0.05 * rng.rand(n, n)
    D = (D + D.T) / 2
    np.fill_diagonal(D, 0)
    labels = ["L%d" % k for k in range(n)]
    p = nb().Params(n_ind=n, n_sites=64, in_probs=True, indep_geno=True)
    with nb().NgsDistB200(p) as g:
        got = g.nj_tree(D, labels)
    want, _, _ = nj_oracle.nj(D, labels)
    gt, gl = nj_oracle.newick_lengths(got)
    wt, wl = nj_oracle.newick_lengths(want)
    assert gt == wt, "different topology"
    assert np.allclose(gl, wl, rtol=0, atol=2e-10)


def test_nj_recovers_an_additive_tree_and_runs_on_the_resident_matrix():
    n = 40
    D = additive_tree_matrix(n, 5)
    p = nb().Params(n_ind=n, n_sites=640, in_probs=True, indep_geno=True, evol_model=0)
    with nb().NgsDistB200(p) as g:
        t = g.nj_tree(D)
        want, _, _ = nj_oracle.nj(D)
        # on an additive matrix every cherry has the same Q exactly, so rounding picks the join order (and the Newick text);
        # the unrooted tree -- its set of bipartitions -- is what neighbour joining guarantees
        assert nj_oracle.splits(t) == nj_oracle.splits(want)
        assert len(nj_oracle.splits(t)) > 0
        # ... and straight from the matrix ngsd_distances left on the device
        g.push_sites(oracle.synth_raw(8, 0.0, n, 640))
        d = g.run()[0]["dist"]
        t_dev = g.nj_tree()
        t_host = g.nj_tree(d)
        assert t_dev == t_host
        assert nj_oracle.splits(t_dev) == nj_oracle.splits(nj_oracle.nj(d)[0])


def test_nj_refuses_non_finite_matrices():
    n = 5
    D = np.ones((n, n)) - np.eye(n)
    D[0, 1] = D[1, 0] = np.nan
    p = nb().Params(n_ind=n, n_sites=64, in_probs=True, indep_geno=True)
    with nb().NgsDistB200(p) as g:
        with pytest.raises(nb().NgsDistError):
            g.nj_tree(D)
