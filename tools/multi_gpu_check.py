"""N-GPU check of the sharding modes against a single-GPU run, through the C ABI only (launch with torchrun, one rank
per GPU).  torch.distributed (gloo, CPU) does nothing but carry the 128-byte NCCL id from rank 0 to the other ranks;
every collective on the data path is issued by libngsdist_b200.so itself (ngsd_comm_*, ngsd_distances_batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import ngsdist_b200 as nb
from ngsdist_b200 import multi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
n_ind, n_sites, bs, nrep, seed = 700, 20000, 100, 5, 12345

MODE = {}   # extra Params of the current pass: {} = soft posteriors (FP64 contraction), call_geno = integer path


def share_id():
    box = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def make(n_sites_ctx, s0, push=None, attach=True, **kw):
    """context for n_sites_ctx sites; push = (first, count) in context coordinates of the sites to push (default all),
    data = the global synthetic stream starting at site s0."""
    kw = dict(MODE, **kw)
    p = nb.Params(n_ind=n_ind, n_sites=n_sites_ctx, indep_geno=True, pairwise_del=kw.pop("pairwise_del", True), evol_model=2, n_boot_rep=nrep,
                  boot_block_size=bs, seed=seed, **kw)
    g = nb.NgsDistB200(p, device=local)
    first, count = push if push else (0, n_sites_ctx)
    if count:
        buf = torch.empty((count, n_ind, 3), dtype=torch.float64, device="cuda")
        g.synth_raw_device(buf.data_ptr(), 99, 0.1, s0 + first, count)     # global site index -> same data as the full run
        g.push_sites_device(buf.data_ptr(), first, count)
    if attach:
        g.comm_attach(share_id(), rank, world)
    return g


def relerr(a, b):
    return float(np.nanmax(np.abs(a - b) / np.abs(b + np.eye(n_ind))))


passes = (("soft posteriors / FP64 contraction (3 planes)", {}),
          ("soft posteriors / FP64 contraction (2 planes)", dict(pairwise_del=False)),
          ("called genotypes / int8 contraction", dict(call_geno=True, in_probs=True)))
for label, mode in passes:
    MODE = mode
    if rank == 0:
        print("== " + label, flush=True)
    full = make(n_sites, 0)
    ref = full.run(want_num=True, want_cnt=True)            # single-GPU reference on every rank

    # 1. replicates: ONE ngsd_distances_batch call, replicate r on rank r % world, gathered on rank 0 by NCCL send/recv
    boot = multi.BootStream(n_sites, bs, seed)
    counts = np.stack([boot.next_counts() for _ in range(nrep)])
    mats = full.distances_batch(counts, bs)
    if rank == 0:
        for r in range(nrep):
            assert np.array_equal(mats[r], ref[r + 1]["dist"], equal_nan=True), "replicate %d differs" % (r + 1)
        print("replicate sharding: %d matrices bit-identical to the single-GPU run (%d bytes over NVLink)" % (nrep, full.comm_stats()[0]), flush=True)

    # 2. tiles: ngsd_set_tile_shard + ngsd_comm_reduce_tiles
    sh = make(n_sites, 0)
    sh.set_tile_shard(rank, world)
    own = multi.tile_owner_mask(n_ind, rank, world)
    r0 = sh.distances(want_num=True, want_cnt=True)
    assert (r0["dist"][~own] == 0).all() and (r0["cnt"][~own] == 0).all(), "entries outside the shard must be 0"
    sh.partial_sums()
    tiles = sh.comm_reduce_tiles(0, want_num_cnt=True)
    if rank == 0:
        assert np.array_equal(tiles["cnt"], ref[0]["cnt"]), "cnt"
        relt = max(relerr(tiles[k], ref[0][k]) for k in ("dist", "num"))   # the K-split plan depends on the tiles owned: FP64 summation order differs
        assert relt < 1e-13, relt
        print("tile sharding: NCCL-assembled matrices == single-GPU matrices (cnt exact, dist/num within %.1e)" % relt, flush=True)
    sh.close()

    # 3. sites: partial sums per rank, ONE reduce of the packed upper triangle (+ cnt under --pairwise_del), epilogue on the root
    shards = multi.site_shards(n_sites, bs, world)
    s0, s1 = shards[rank]
    loc = make(s1 - s0, s0)
    boot = multi.BootStream(n_sites, bs, seed)
    loc.partial_sums(None, 1)
    d0 = loc.comm_reduce_sites(0, n_sites)
    rel = 0.0
    if rank == 0:
        rel = relerr(d0, ref[0]["dist"])
        assert rel < 1e-12, rel
    for rep in range(1, nrep + 1):
        c = boot.next_counts()
        loc.partial_sums(multi.slice_block_counts(c, shards[rank], bs), bs)
        d = loc.comm_reduce_sites(0, len(c) * bs)
        if rank == 0:
            rel = max(rel, relerr(d, ref[rep]["dist"]))
            assert rel < 1e-12, (rep, rel)
    if rank == 0:
        b, ms = loc.comm_stats()
        print("site sharding: %d matrices within %.1e of the single-GPU run; reduce moved %d bytes (full num+cnt matrices: %d)" %
              (nrep + 1, rel, b, n_ind * n_ind * 16), flush=True)
    loc.close()

    # 4. site-sharded front end + NCCL all-gather of the packed operands -> every rank holds every site
    align = 192
    sb = [n_sites // align * r // world * align for r in range(world)] + [n_sites]
    ag = make(n_sites, 0, push=(sb[rank], sb[rank + 1] - sb[rank]))
    ag.comm_allgather_operands(sb)
    ra = ag.distances(want_num=True, want_cnt=True)
    for k in ("dist", "num", "cnt"):
        assert np.array_equal(ra[k], ref[0][k], equal_nan=True), "all-gathered operands give a different " + k
    if rank == 0:
        print("all-gather: matrices from all-gathered operands bit-identical (%d bytes sent+received per rank)" % ag.comm_stats()[0], flush=True)
    ag.close()
    full.close()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTI_GPU_CHECK_OK world=%d" % world)
