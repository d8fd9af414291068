import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ngsdist_b200 as nb
from oracle import nj_oracle
for n in (10, 12, 16, 24, 33):
    rng = np.random.RandomState(n)
    X = rng.rand(n, 6)
    D = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(axis=2)) + 0.05 * rng.rand(n, n)
    D = (D + D.T) / 2
    np.fill_diagonal(D, 0)
    with nb.NgsDistB200(nb.Params(n_ind=n, n_sites=64, in_probs=True, indep_geno=True)) as g:
        got = g.nj_tree(D)
    want, joins, _ = nj_oracle.nj(D)
    same = nj_oracle.newick_lengths(got)[0] == nj_oracle.newick_lengths(want)[0]
    print(n, "same topology:", same)
    if not same:
        print(" got ", got[:400]); print(" want", want[:400]); print(" oracle joins", [(a, b) for a, b, _, _ in joins][:12])
