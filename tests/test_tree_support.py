"""ngsd_tree_support (host code of the library, no GPU): bootstrap support of the main tree's internal edges from the
replicate trees -- the `raxmlHPC -f b` step of the reference's workflow (README.md:83-98).  The reference has no code for
it; the checks restate the definition (a replicate supports an edge when it holds the same bipartition of the leaves)."""
import re

import numpy as np
import pytest

import ngsdist_b200 as nb
from oracle import nj_oracle


def parse(newick):
    """Nested lists of (children | leaf name, node label, length text)."""
    tok = re.findall(r"[(),;]|:[^(),;]+|[^(),;:]+", newick.strip())
    pos = 0

    def node():
        nonlocal pos
        if tok[pos] == "(":
            pos += 1
            kids = [node()]
            while tok[pos] == ",":
                pos += 1
                kids.append(node())
            assert tok[pos] == ")"
            pos += 1
            label = ""
            if pos < len(tok) and tok[pos] not in "(),;" and not tok[pos].startswith(":"):
                label = tok[pos]
                pos += 1
        else:
            kids, label = None, tok[pos]
            pos += 1
        length = ""
        if pos < len(tok) and tok[pos].startswith(":"):
            length = tok[pos]
            pos += 1
        return [kids, label, length]

    root = node()
    assert tok[pos] == ";"
    return root


def leaves(n):
    return [n[1]] if n[0] is None else [x for k in n[0] for x in leaves(k)]


def split_sets(root):
    allv = frozenset(leaves(root))
    out = set()

    def walk(n, is_root):
        if n[0] is None:
            return
        side = frozenset(leaves(n))
        if not is_root and 1 < len(side) < len(allv) - 1:
            out.add(min(side, allv - side, key=sorted))
        for k in n[0]:
            walk(k, False)

    walk(root, True)
    return out, allv


def expected(main, reps, percent):
    root = parse(main)
    _, allv = split_sets(root)
    rep_sets = [split_sets(parse(r))[0] for r in reps if r != "NA"]

    def render(n, is_root):
        if n[0] is None:
            return n[1] + n[2]
        side = frozenset(leaves(n))
        lab = ""
        if not is_root and 1 < len(side) < len(allv) - 1:
            key = min(side, allv - side, key=sorted)
            c = sum(key in s for s in rep_sets)
            lab = str(int(0.5 + 100.0 * c / len(rep_sets))) if percent else str(c)
        return "(" + ",".join(render(k, False) for k in n[0]) + ")" + lab + n[2]

    return render(root, True) + ";"


def test_hand_example_counts_and_percent():
    main = "((A:1,B:1):0.5,(C:1,D:1):0.25,E:1);"
    reps = ["((A:1,B:1):1,(C:1,D:1):1,E:1);",          # both edges
            "((A:1,C:1):1,(B:1,D:1):1,E:1);",          # neither
            "(((B,A),C),(D,E));",                      # rooted, children swapped: AB|CDE yes, CD|ABE no
            "NA"]                                      # skipped
    assert nb.tree_support(main, reps, percent=False) == "((A:1,B:1)2:0.5,(C:1,D:1)1:0.25,E:1);"
    assert nb.tree_support(main, reps, percent=True) == "((A:1,B:1)67:0.5,(C:1,D:1)33:0.25,E:1);"
    assert nb.tree_support(main, [], percent=False) == "((A:1,B:1)0:0.5,(C:1,D:1)0:0.25,E:1);"


def test_rooting_and_child_order_do_not_matter():
    main = "((a:1,b:1):1,((c:1,d:1):1,e:1):1,(f:1,g:1):1);"
    same = ["(g,f,((e,(d,c)),(b,a)));", "((((a,b),(f,g)),e),(c,d));", "(a:2,b:3,((g,f),(e,(c,d))));"]
    out = nb.tree_support(main, same, percent=False)
    assert out == "((a:1,b:1)3:1,((c:1,d:1)3:1,e:1)3:1,(f:1,g:1)3:1);"


@pytest.mark.parametrize("n,seed", [(12, 1), (40, 2), (150, 3)])
def test_noisy_neighbour_joining_replicates_against_the_definition(n, seed):
    rng = np.random.RandomState(seed)
    pts = rng.rand(n, 6)
    base = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    labels = ["s%03d" % k for k in range(n)]
    main = nj_oracle.nj(base, labels)[0]
    reps = []
    for _ in range(25):
        noise = rng.rand(n, n) * 0.15
        d = base * (1 + noise + noise.T)
        np.fill_diagonal(d, 0)
        reps.append(nj_oracle.nj(d, labels)[0])
    reps.insert(7, "NA")
    for percent in (False, True):
        got = nb.tree_support(main, reps, percent=percent)
        assert got == expected(main, reps, percent)
    counts = [int(x) for x in re.findall(r"\)(\d+)", nb.tree_support(main, reps, percent=False))]
    assert len(counts) == n - 3 and 0 < sum(counts) < 25 * (n - 3)      # some edges hold, some do not: the test has teeth


def test_quoted_labels_and_errors():
    main = "(('sample one':1,'it''s':2):1,x:1,y:1,z:2);"
    assert nb.tree_support(main, ["((x,y),z,('it''s','sample one'));"], percent=False) == "((sample one:1,it's:2)1:1,x:1,y:1,z:2);"
    for bad_main, reps in [("((A,B),C", []), ("((A,B),(C,D)));", []), ("((A,A),B,C);", []),
                           ("((A,B),C,D);", ["((A,B),C,E);"]), ("((A,B),C,D);", ["((A,B),C);"]), ("((A,B),C,D);", ["(A,B),C,D;"])]:
        with pytest.raises(ValueError):
            nb.tree_support(bad_main, reps)
