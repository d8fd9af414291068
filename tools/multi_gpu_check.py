"""N-GPU check of the three sharding modes against a single-GPU run (launch with torchrun, one rank per GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import ngsdist_b200 as nb
from ngsdist_b200 import multi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_ind, n_sites, bs, nrep, seed = 700, 20000, 100, 5, 12345

MODE = {}   # extra Params of the current pass: {} = soft posteriors (FP64 contraction), call_geno = integer path

def make(n_sites_local, s0, **kw):
    kw = dict(MODE, **kw)
    p = nb.Params(n_ind=n_ind, n_sites=n_sites_local, indep_geno=True, pairwise_del=True, evol_model=2, n_boot_rep=nrep,
                  boot_block_size=bs, seed=seed, **kw)
    g = nb.NgsDistB200(p, device=local)
    buf = torch.empty((n_sites_local, n_ind, 3), dtype=torch.float64, device="cuda")
    g.synth_raw_device(buf.data_ptr(), 99, 0.1, s0, n_sites_local)     # global site index s0.. -> same data as the full run
    g.push_sites_device(buf.data_ptr(), 0, n_sites_local)
    g.frontend()
    return g

for label, mode in (("soft posteriors / FP64 contraction", {}), ("called genotypes / int8 contraction", dict(call_geno=True, in_probs=True))):
    MODE = mode
    if rank == 0:
        print("== " + label)
    full = make(n_sites, 0)
    ref = full.run(want_num=True, want_cnt=True)            # single-GPU reference on every rank

    # 1. replicates
    boot = multi.BootStream(n_sites, bs, seed)
    mats = multi.run_replicates(nrep, boot, lambda rep, c, b: full.distances(c, b)["dist"], rank, world)
    if rank == 0:
        for r, (m, w) in enumerate(zip(mats, ref)):
            assert np.array_equal(m, w["dist"], equal_nan=True), "replicate %d differs" % r
        print("replicate sharding: %d matrices bit-identical to the single-GPU run" % len(mats))

    # 2. tiles
    sh = make(n_sites, 0)
    sh.set_tile_shard(rank, world)
    own = multi.tile_owner_mask(n_ind, rank, world)
    r0 = sh.distances(want_num=True, want_cnt=True)
    assert (r0["dist"][~own] == 0).all() and (r0["cnt"][~own] == 0).all(), "entries outside the shard must be 0"
    tiles = multi.run_tiles(lambda: r0, rank, world)
    assert np.array_equal(tiles["cnt"], ref[0]["cnt"]), "cnt"
    relt = 0.0
    for k in ("dist", "num"):       # the K-split plan depends on the number of owned tiles, so the FP64 summation order differs
        relt = max(relt, np.nanmax(np.abs(tiles[k] - ref[0][k]) / np.abs(ref[0][k] + np.eye(n_ind))))
    assert relt < 1e-13, relt
    if rank == 0:
        print("tile sharding: SUM over ranks == single-GPU matrices (cnt exact, dist/num within %.1e)" % relt)

    # 3. sites (+ NCCL all-reduce of the library's device buffers, epilogue after the reduction)
    shards = multi.site_shards(n_sites, bs, world)
    s0, s1 = shards[rank]
    loc = make(s1 - s0, s0)
    boot = multi.BootStream(n_sites, bs, seed)
    d0 = multi.run_sites_gpu(loc, None, 1)
    rel = np.nanmax(np.abs(d0 - ref[0]["dist"]) / np.abs(ref[0]["dist"] + np.eye(n_ind)))
    assert rel < 1e-12, rel
    for rep in range(1, nrep + 1):
        counts = boot.next_counts()
        d = multi.run_sites_gpu(loc, multi.slice_block_counts(counts, shards[rank], bs), bs)
        rel = max(rel, np.nanmax(np.abs(d - ref[rep]["dist"]) / np.abs(ref[rep]["dist"] + np.eye(n_ind))))
        assert rel < 1e-12, (rep, rel)
    if rank == 0:
        print("site sharding: %d matrices within %.1e of the single-GPU run after the NCCL reduce" % (nrep + 1, rel))
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTI_GPU_CHECK_OK world=%d" % world)
