"""N4: neighbour-joining trees from the resident distance matrices (ngsd_nj_tree) against the CPU
restatement of the same published rules (oracle/nj_oracle.py).  The reference has no tree code: parity unpinned by it."""
import numpy as np
import pytest

import oracle
from oracle import nj_oracle
from test_gpu_parity import nb

pytestmark = pytest.mark.gpu


def additive_tree_matrix(n, seed):
    """Leaf-to-leaf path lengths of a random unrooted binary tree with positive branch lengths (leaves 0..n-1), and the
    tree's non-trivial bipartitions: neighbour joining must recover exactly this tree."""
    rng = np.random.RandomState(seed)
    # edges as (u, v, length); start from a star on leaves 0, 1, 2 around internal node n
    edges = [(0, n, rng.uniform(0.05, 0.3)), (1, n, rng.uniform(0.05, 0.3)), (2, n, rng.uniform(0.05, 0.3))]
    nxt = n + 1
    for leaf in range(3, n):                      # split a random edge and hang the new leaf there
        u, v, L = edges.pop(int(rng.randint(len(edges))))
        x = rng.uniform(0.2, 0.8) * L
        edges += [(u, nxt, x), (nxt, v, L - x), (leaf, nxt, rng.uniform(0.05, 0.3))]
        nxt += 1
    N = nxt
    adj = [[] for _ in range(N)]
    for u, v, L in edges:
        adj[u].append((v, L)); adj[v].append((u, L))
    D = np.zeros((n, n))
    for a in range(n):
        dist = {a: 0.0}
        stack = [a]
        while stack:
            u = stack.pop()
            for v, L in adj[u]:
                if v not in dist:
                    dist[v] = dist[u] + L
                    stack.append(v)
        D[a] = [dist[b] for b in range(n)]
    true_splits = set()
    for u, v, L in edges:                         # removing an edge splits the leaves
        seen = {u}
        stack = [u]
        while stack:
            w = stack.pop()
            for y, _ in adj[w]:
                if y not in seen and not (w == u and y == v):
                    seen.add(y); stack.append(y)
        side = frozenset("Ind_%d" % k for k in seen if k < n)
        if "Ind_0" in side:
            side = frozenset("Ind_%d" % k for k in range(n)) - side
        if 1 < len(side) < n - 1:
            true_splits.add(side)
    return D, true_splits


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
    D, true_splits = additive_tree_matrix(n, 5)
    p = nb().Params(n_ind=n, n_sites=640, in_probs=True, indep_geno=True, evol_model=0)
    with nb().NgsDistB200(p) as g:
        t = g.nj_tree(D)
        want, _, _ = nj_oracle.nj(D)
        # on an additive matrix every cherry has the same Q exactly, so rounding picks the join order (and the Newick text);
        # the unrooted tree -- its set of bipartitions -- is what neighbour joining guarantees
        assert nj_oracle.splits(t) == nj_oracle.splits(want) == true_splits
        assert len(true_splits) == n - 3
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
