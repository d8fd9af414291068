"""Tile shards on ONE GPU: every rank's share computed in turn must tile the full matrix exactly (what the NCCL SUM of
multi.py assembles across GPUs).  Covers the (split, tile pair) units of dist_umma.cu on shard-local tile lists, where the
tiles of one row block are no longer neighbours of the full list, and the FP64 contraction on the same shards."""
import numpy as np
import pytest

import oracle
from ngsdist_b200 import multi
from test_gpu_parity import nb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("called,pdel", [(True, False), (True, True), (False, True)])
@pytest.mark.parametrize("world", [2, 3])
def test_tile_shards_tile_the_matrix(called, pdel, world):
    n_ind, n_sites = 700, 1500          # 6 row blocks, 21 tiles
    raw = oracle.synth_raw(31, 0.1, n_ind, n_sites)
    p = nb().Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=True, call_geno=called, pairwise_del=pdel, evol_model=0)
    with nb().NgsDistB200(p) as g:
        g.push_sites(raw)
        g.frontend()
        full = g.distances(want_num=True, want_cnt=True)
        num = np.zeros_like(full["num"])
        cnt = np.zeros_like(full["cnt"])
        covered = np.zeros((n_ind, n_ind), dtype=int)
        for rank in range(world):
            g.set_tile_shard(rank, world)
            r = g.distances(want_num=True, want_cnt=True)
            own = multi.tile_owner_mask(n_ind, rank, world)
            assert not r["num"][~own].any() and not r["cnt"][~own].any(), "entries outside the shard must be zero"
            num += r["num"]
            cnt += r["cnt"]
            covered += own
        g.set_tile_shard(0, 1)
        again = g.distances(want_num=True, want_cnt=True)
    off = ~np.eye(n_ind, dtype=bool)
    assert (covered[off] == 1).all()
    assert np.array_equal(cnt, full["cnt"])
    if called:
        assert np.array_equal(num, full["num"])               # integer sums: the split into shards cannot change a bit
    else:
        assert np.allclose(num, full["num"], rtol=1e-12, atol=0)   # K splits are planned per shard: summation order differs
    assert np.array_equal(again["num"], full["num"]) and np.array_equal(again["cnt"], full["cnt"])
