"""CPU restatement of the neighbour-joining rules of ngsdist_b200/csrc/nj.cu -- TEST INFRASTRUCTURE ONLY.

The reference has no tree code (its README.md:83-98 hands the .dist file to FastME), so there is nothing of the
reference's to pin this against: PARITY UNPINNED.  It restates the published algorithm (Saitou & Nei 1987 in the
Studier & Keppler 1988 formulation):
    Q(i,j) = (m - 2) d(i,j) - r_i - r_j          the active pair with the smallest Q is joined (ties: smallest i, then j)
    limb_i = d(i,j)/2 + (r_i - r_j) / (2 (m - 2)),  limb_j = d(i,j) - limb_i
    d(u,k) = (d(i,k) + d(j,k) - d(i,j)) / 2       u takes slot i, slot j retires
until three nodes are left (trifurcating root).  Returns Newick with "%.10f" lengths, the join list and the final limbs."""
import numpy as np


def nj(dist, labels=None):
    D = np.array(dist, dtype=np.float64)
    n = D.shape[0]
    assert n >= 3 and np.isfinite(D[~np.eye(n, dtype=bool)]).all()
    np.fill_diagonal(D, 0.0)
    sub = [labels[k] if labels is not None else "Ind_%d" % k for k in range(n)]
    active = np.ones(n, dtype=bool)
    r = D.sum(axis=1)
    joins = []
    m = n
    while m > 3:
        idx = np.flatnonzero(active)
        sub_d = D[np.ix_(idx, idx)]
        Q = (m - 2) * sub_d - r[idx][:, None] - r[idx][None, :]
        Q[np.tril_indices(len(idx))] = np.inf                    # pairs i < j only
        if m == 4:                                               # Q of a pair equals Q of the complementary pair exactly: every
            Q[1:, :] = np.inf                                    # pairing is represented once, by the pair of the smallest index
        a, b = np.unravel_index(np.argmin(Q), Q.shape)           # first minimum in row-major order = smallest i, then j
        i, j = int(idx[a]), int(idx[b])
        dij = D[i, j]
        li = dij / 2 + (r[i] - r[j]) / (2 * (m - 2))
        lj = dij - li
        joins.append((i, j, li, lj))
        ks = [k for k in idx if k != i and k != j]
        duk = (D[i, ks] + D[j, ks] - dij) / 2
        r[ks] += duk - D[i, ks] - D[j, ks]
        D[i, ks] = duk
        D[ks, i] = duk
        r[i] = duk.sum()
        active[j] = False
        sub[i] = "(%s:%.10f,%s:%.10f)" % (sub[i], li, sub[j], lj)
        sub[j] = ""
        m -= 1
    a, b, c = [int(k) for k in np.flatnonzero(active)]
    la = (D[a, b] + D[a, c] - D[b, c]) / 2
    lb = (D[a, b] + D[b, c] - D[a, c]) / 2
    lc = (D[a, c] + D[b, c] - D[a, b]) / 2
    return "(%s:%.10f,%s:%.10f,%s:%.10f);" % (sub[a], la, sub[b], lb, sub[c], lc), joins, (la, lb, lc)


def newick_lengths(s):
    """Topology (string with the lengths cut out) and the list of branch lengths, for tolerant comparisons."""
    import re
    lens = [float(x) for x in re.findall(r":(-?[0-9.]+(?:e-?[0-9]+)?|-?nan|-?inf)", s)]
    return re.sub(r":-?[0-9.a-z]+(?:e-?[0-9]+)?", "", s), lens


def splits(newick):
    """The unrooted topology as a set of leaf bipartitions (each the frozenset of the side without the first leaf)."""
    import re
    tokens = re.findall(r"[(),;]|[^(),;:]+|:[^(),;]+", newick)
    stack, out, leaves = [], set(), []
    cur = None
    for t in tokens:
        if t == "(":
            stack.append(set())
        elif t == ")":
            done = stack.pop()
            out.add(frozenset(done))
            if stack:
                stack[-1] |= done
            cur = done
        elif t in (",", ";") or t.startswith(":"):
            continue
        else:
            leaves.append(t)
            if stack:
                stack[-1].add(t)
    allv = frozenset(leaves)
    first = leaves[0]
    norm = set()
    for sp in out:
        side = sp if first not in sp else allv - sp
        if 1 < len(side) < len(allv) - 1:
            norm.add(frozenset(side))
    return norm
