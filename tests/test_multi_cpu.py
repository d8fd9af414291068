"""world_size-2 `gloo` tests of the multi-GPU orchestration (ngsdist_b200/multi.py) on CPU.

The collective wiring, shard plans and host RNG bookkeeping are the product code under test; the per-rank compute
callback is a stand-in built on the CPU oracle (there is no CPU build of the CUDA kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from ngsdist_b200 import multi

N_IND, N_SITES, BS, NREP, SEED = 150, 230, 10, 4, 12345


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    raw = oracle.synth_raw(77, 0.15, N_IND, N_SITES)
    return oracle.frontend(raw)


def _counts_to_map(counts, bs):
    blocks = np.repeat(np.arange(len(counts), dtype=np.uint64), counts)
    return (blocks[:, None] * bs + np.arange(bs, dtype=np.uint64)[None, :]).reshape(-1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = _data()
        kw = dict(indep=True, pairwise_del=True, evol_model=2)
        # ---- replicates ----
        def compute(rep, counts, bs):
            sm = None if counts is None else _counts_to_map(counts, bs)
            return oracle.distances(P, site_map=sm, **kw)["dist"]
        boot = multi.BootStream(N_SITES, BS, SEED)
        mats = multi.run_replicates(NREP, boot, compute, rank, world)
        # ---- tiles ----
        full = oracle.distances(P, **kw)
        own = multi.tile_owner_mask(N_IND, rank, world)
        def compute_owned():
            return {k: np.where(own, v, 0).astype(v.dtype) for k, v in full.items()}
        tiles = multi.run_tiles(compute_owned, rank, world)
        # ---- sites ----
        shards = multi.site_shards(N_SITES, BS, world)
        s0, s1 = shards[rank]
        boot2 = multi.BootStream(N_SITES, BS, SEED)
        counts = boot2.next_counts()
        local = multi.slice_block_counts(counts, shards[rank], BS)
        part = oracle.distances(P[:, s0:s1], site_map=_counts_to_map(local, BS), evol_model=0, indep=True, pairwise_del=True)
        num = torch.from_numpy(part["num"].copy())
        cnt = torch.from_numpy(part["cnt"].astype(np.int64))
        multi.reduce_site_partials(num, cnt)
        if rank == 0:
            q.put(dict(mats=mats, tiles=tiles, site_num=num.numpy(), site_cnt=cnt.numpy(), counts=counts))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_orchestration():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P = _data()
    # replicates: identical to the single-process replicate loop (same RNG stream, same order)
    raw = oracle.synth_raw(77, 0.15, N_IND, N_SITES)
    want = oracle.run_job(raw, indep=True, pairwise_del=True, evol_model=2, n_boot_rep=NREP, boot_block_size=BS, seed=SEED)
    assert len(got["mats"]) == NREP + 1
    for g, w in zip(got["mats"], want):
        assert np.allclose(g, w["dist"], rtol=1e-12, atol=0, equal_nan=True)
    # tiles: the SUM of owned entries is the full matrix, exactly
    full = oracle.distances(P, indep=True, pairwise_del=True, evol_model=2)
    for k in ("dist", "num", "cnt"):
        assert np.array_equal(got["tiles"][k], full[k]), k
    # sites: reduced raw sums equal the unsharded replicate's sums; counts exactly
    sm = oracle.Taus(SEED).boot_map(N_SITES // BS, BS)
    ref = oracle.distances(P, site_map=sm, evol_model=0, indep=True, pairwise_del=True)
    assert np.array_equal(got["site_cnt"].astype(np.uint64), ref["cnt"])
    assert np.allclose(got["site_num"], ref["num"], rtol=1e-12, atol=0)


def test_shard_plans():
    for n_sites, bs, world in [(1000, 10, 3), (1003, 10, 8), (64, 1, 2), (999, 1000, 2), (10 ** 6, 1000, 8)]:
        sh = multi.site_shards(n_sites, bs, world)
        assert sh[0][0] == 0 and sh[-1][1] == n_sites and len(sh) == world
        for (a0, a1), (b0, b1) in zip(sh, sh[1:]):
            assert a1 == b0 and a0 <= a1
        for s0, _ in sh:
            assert s0 % bs == 0
        counts = np.arange(n_sites // bs, dtype=np.uint32)
        back = np.concatenate([multi.slice_block_counts(counts, s, bs) for s in sh]) if n_sites // bs else counts
        assert np.array_equal(back, counts)
    assert multi.replicate_shard(7, 1, 3) == [1, 4]
    # every off-diagonal entry is owned by exactly one rank
    for n, world in [(150, 2), (500, 4), (1300, 8)]:
        tot = sum(multi.tile_owner_mask(n, r, world).astype(int) for r in range(world))
        assert np.array_equal(tot, 1 - np.eye(n, dtype=int))


def test_bootstream_matches_oracle_taus():
    boot = multi.BootStream(103, 10, 4242)
    t = oracle.Taus(4242)
    for _ in range(3):
        c = boot.next_counts()
        sm = t.boot_map(10, 10)
        want = np.bincount((sm[::10] // 10).astype(int), minlength=10)
        assert np.array_equal(c, want)


def test_tile_shards_keep_row_block_pairs_together_and_balanced():
    """Ownership goes by pairs of neighbouring tiles of one row block (what dist_umma.cu contracts with one shared A):
    a rank's tiles split into such pairs again, and the ranks' tile counts differ by at most 2 x ceil-rounding."""
    for n, world in [(5000, 8), (1300, 3), (700, 2), (128, 4)]:
        tiles = multi.tile_list(n)
        owner = np.full(len(tiles), -1)
        for r in range(world):
            m = multi.tile_owner_mask(n, r, world)
            for k, (ti, tj) in enumerate(tiles):
                i, j = ti * 128, tj * 128 + (1 if ti == tj else 0)
                if j < n and m[i, j]:
                    owner[k] = r
        known = owner >= 0                                  # a 1-individual diagonal tile owns no off-diagonal entry
        k = 0
        while k < len(tiles):
            ln = 2 if k + 1 < len(tiles) and tiles[k + 1][0] == tiles[k][0] else 1
            if ln == 2 and known[k] and known[k + 1]:
                assert owner[k] == owner[k + 1]
            k += ln
        cnt = np.bincount(owner[known], minlength=world)
        assert cnt.max() - cnt.min() <= 2 + (len(tiles) % 2), (n, world, cnt)


def test_pack_genotypes_layout():
    """pack_genotypes: individual i of a site in bits 2 (i % 4) of byte i / 4; default fields {0,1,2,-1} -> {0,1,2,3}."""
    from ngsdist_b200 import pack_genotypes, PLINK_BED_CODES
    g = np.array([[0, 1, 2, -1, 2], [-1, -1, 0, 0, 1]], dtype=np.int8)
    p = pack_genotypes(g)
    assert p.shape == (2, 2) and p.dtype == np.uint8
    assert p[0, 0] == (0 | 1 << 2 | 2 << 4 | 3 << 6) and p[0, 1] == 2
    assert p[1, 0] == (3 | 3 << 2) and p[1, 1] == 1
    bed = pack_genotypes(g, field_of_code=[1, 0, 2, 3])     # PLINK: 00 hom A1, 01 missing, 10 het, 11 hom A2
    fields = [(bed[0, i // 4] >> (2 * (i % 4))) & 3 for i in range(5)]
    assert [PLINK_BED_CODES[f] for f in fields] == list(g[0])
