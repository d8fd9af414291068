"""Full C3 (BASELINE.json configs[2]): 2 000 individuals x 1 000 000 sites, 10 % missing, --pairwise_del --indep_geno,
replicate 0 + 100 bootstrap replicates (block 1000, --seed 12345), each contracted with its block weights (reserved bit 2).
Launch plain (1 GPU) or with torchrun (N ranks): ONE ngsd_distances_batch call shards the replicates (r on rank r % N) and
gathers the matrices on rank 0 by NCCL below the C ABI.  Prints one JSON line; --check verifies size-independent properties."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import ngsdist_b200 as nb
from ngsdist_b200 import multi

ap = argparse.ArgumentParser()
ap.add_argument("--n-ind", type=int, default=2000)
ap.add_argument("--n-sites", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=100)
ap.add_argument("--block", type=int, default=1000)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
n, S, R, B = args.n_ind, args.n_sites, args.reps, args.block
p = nb.Params(n_ind=n, n_sites=S, in_probs=True, indep_geno=True, pairwise_del=True, evol_model=1, n_boot_rep=R, boot_block_size=B, seed=12345,
              no_block_cache=True)
g = nb.NgsDistB200(p, device=local)
if world > 1:
    box = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    g.comm_attach(box[0], rank, world)
t0 = time.time()
buf = torch.empty((16384, n, 3), dtype=torch.float64, device="cuda")
fe_ms = 0.0
for s0 in range(0, S, 16384):
    m = min(16384, S - s0)
    g.synth_raw_device(buf.data_ptr(), 20251018, 0.10, s0, m)
    g.push_sites_device(buf.data_ptr(), s0, m)
    fe_ms += g.timing().frontend_ms
del buf
g.frontend()
setup_s = time.time() - t0
boot = multi.BootStream(S, B, 12345)
counts = np.stack([boot.next_counts() for _ in range(R)])
out = torch.empty((R, n, n), dtype=torch.float64, pin_memory=True) if rank == 0 else None
first = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
# untimed warm-up: one replicate per rank through the same call (first-use allocations of the library and of NCCL)
g._check(nb.lib().ngsd_distances_batch(g._h, counts.ctypes.data, world, counts.shape[1], B, out.data_ptr() if rank == 0 else None))
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t1 = time.time()
if world > 1:                                            # replicate 0: its output-triangle tiles dealt to the ranks, NCCL assembly on rank 0
    g.set_tile_shard(rank, world)
    g.partial_sums()
    g._check(nb.lib().ngsd_comm_reduce_tiles(g._h, 0, 0, first.data_ptr() if rank == 0 else None, None, None))
    g.set_tile_shard(0, 1)
else:
    g.distances_raw(None, 0, 1, first.data_ptr())
t_rep0 = time.time() - t1
t2 = time.time()
g._check(nb.lib().ngsd_distances_batch(g._h, counts.ctypes.data, R, counts.shape[1], B, out.data_ptr() if rank == 0 else None))
if world > 1:
    dist.barrier()
t_batch = time.time() - t2
tim = g.timing()
if rank == 0:
    pairs = n * (n - 1) // 2
    n_eff = counts.shape[1] * B
    rep = dict(workload="C3 full: %d ind x %d sites, 10%% missing, --pairwise_del --indep_geno, 1 + %d matrices, block %d, weighted contraction" % (n, S, R, B),
               world=world, setup_s=setup_s, frontend_kernels_ms=fe_ms, replicate0_s=t_rep0, batch_s=t_batch, per_replicate_ms=t_batch / R * 1e3 * world,
               job_s=t_rep0 + t_batch, pair_sites_per_s=pairs * (S + n_eff * R) / (t_rep0 + t_batch),
               rank0_contraction_ms_sum=tim.dist_ms, rank0_dmma_tflops=tim.dist_dmma * 512 / (tim.dist_ms * 1e-3) * 1e-12,
               nccl_gather_bytes=g.comm_stats()[0] if world > 1 else 0, nccl_gather_ms=g.comm_stats()[1] if world > 1 else 0.0)
    o = out.numpy()
    rep["checksum"] = float(np.nansum(o[:, 0, 1:8]))     # same on every world size (replicates are bit-identical whichever rank runs them)
    rep["finite_offdiag"] = bool(np.isfinite(o[:, ~np.eye(n, dtype=bool)]).all())
    rep["symmetric"] = bool(all(np.array_equal(o[r], o[r].T) for r in range(0, R, max(1, R // 7))))
    print(json.dumps(rep))
    if args.check:
        # linearity of num / cnt in the block weights (size independent): D(c1) + D(c2) == D(c1 + c2)
        nbk = counts.shape[1]
        rng = np.random.RandomState(1)
        c1 = rng.randint(0, 3, nbk).astype(np.uint32); c2 = rng.randint(0, 3, nbk).astype(np.uint32)
        gg = g
        r1 = gg.distances(c1, B, want_num=True, want_cnt=True)
        r2 = gg.distances(c2, B, want_num=True, want_cnt=True)
        r3 = gg.distances(c1 + c2, B, want_num=True, want_cnt=True)
        assert np.array_equal(r1["cnt"] + r2["cnt"], r3["cnt"]), "cnt not linear"
        rel = np.abs(r1["num"] + r2["num"] - r3["num"]).max() / np.abs(r3["num"]).max()
        assert rel < 1e-12, rel
        print("property checks ok: cnt exactly linear in the block weights, num to %.1e" % rel)
g.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
