"""C5-shaped site-sharded run (BASELINE.json configs[4]): n_ind individuals, --avg_nuc_dist --indep_geno, sites sharded
over the ranks (one per GPU), raw sums all-reduced with NCCL over NVLink on the library's own device buffers, epilogue
after the reduction.  Launch with torchrun.  Prints per-phase times and checks a 64 x 64 pair block against the CPU
oracle (which regenerates exactly those individuals' synthetic GLs)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import ngsdist_b200 as nb
from ngsdist_b200 import multi

ap = argparse.ArgumentParser()
ap.add_argument("--n-ind", type=int, default=20000)
ap.add_argument("--sites-per-gpu", type=int, default=40000)
ap.add_argument("--chunk", type=int, default=1024)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_ind = args.n_ind
n_sites = args.sites_per_gpu * world
shards = multi.site_shards(n_sites, 1, world)
s0, s1 = shards[rank]
p = nb.Params(n_ind=n_ind, n_sites=s1 - s0, in_probs=True, indep_geno=True, avg_nuc_dist=True, evol_model=1)
g = nb.NgsDistB200(p, device=local)
buf = torch.empty((args.chunk, n_ind, 3), dtype=torch.float64, device="cuda")
t0 = time.time()
fe_ms = 0.0
for c0 in range(0, s1 - s0, args.chunk):
    m = min(args.chunk, s1 - s0 - c0)
    g.synth_raw_device(buf.data_ptr(), 20251018, 0.0, s0 + c0, m)       # global site index -> same data set on any world size
    g.push_sites_device(buf.data_ptr(), c0, m)
    fe_ms += g.timing().frontend_ms
g.frontend()
torch.cuda.synchronize()
t_fe = time.time() - t0

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

barrier()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
stream = torch.cuda.current_stream()
e[0].record(stream)
g.partial_sums(None, 1)                                  # K2 + K4 on this rank's sites; raw sums stay on the device
tim = g.timing()
_, num_ptr, cnt_ptr = g.device_results()
num = multi.device_tensor(num_ptr, (n_ind, n_ind), "<f8")
cnt = multi.device_tensor(cnt_ptr, (n_ind, n_ind), "<i8")
e[1].record(stream)
if world > 1:
    multi.reduce_site_partials(num, cnt)
e[2].record(stream)
torch.cuda.synchronize()
t1 = time.time()
out_pin = torch.empty((n_ind, n_ind), dtype=torch.float64).pin_memory() if rank == 0 else None
t1 = time.time()
if rank == 0:
    out = g.finish(out=out_pin.numpy())              # epilogue on the reduced sums + D2H of the matrix (pinned)
else:
    nb.api.lib().ngsd_finish(g._h, None)             # other ranks: epilogue only
    out = None
t_fin = time.time() - t1
e[3].record(stream)
barrier()
pairs = n_ind * (n_ind - 1) // 2
ms = [e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])]
tot = torch.tensor([ms[0] + ms[1] + t_fin * 1e3], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
if rank == 0:
    rep = dict(n_ind=n_ind, n_sites=n_sites, world=world, frontend_kernels_ms=fe_ms, setup_wall_s=t_fe, partial_sums_ms=ms[0], dist_kernel_ms=tim.dist_ms,
               epilogue_ms=tim.epilogue_ms, nccl_allreduce_ms=ms[1], allreduce_bytes=2 * n_ind * n_ind * 8, finish_incl_d2h_ms=t_fin * 1e3,
               total_ms_max_over_ranks=float(tot.item()),
               pair_sites_per_s=pairs * n_sites / (float(tot.item()) * 1e-3),
               dmma_tflops_per_gpu=tim.dist_dmma * 512 / (tim.dist_ms * 1e-3) * 1e-12)
    print(json.dumps(rep))
    assert (np.diag(out) == 0).all() and np.array_equal(out[:256, :256], out[:256, :256].T)
    if args.check:
        import oracle
        ids = np.concatenate([np.arange(0, 64), np.arange(n_ind - 64, n_ind)])
        # regenerate only those individuals: the generator is indexed by (site * n_ind + individual)
        raw = np.empty((n_sites, len(ids), 3))
        blk = 4096
        for c0 in range(0, n_sites, blk):
            full = None
            m = min(blk, n_sites - c0)
            # oracle.synth_raw generates all individuals of a site range; at n_ind = 20 000 a 4096-site slab is 2 GB
            full = oracle.synth_raw(20251018, 0.0, n_ind, m, site0=c0)
            raw[c0:c0 + m] = full[:, ids, :]
        P = oracle.frontend(raw)
        o = oracle.distances(P, score=oracle.score_matrix(True), indep=True, evol_model=1)
        got = out[np.ix_(ids, ids)]
        off = ~np.eye(len(ids), dtype=bool)
        rel = np.abs(got[off] - o["dist"][off]).max() / np.abs(o["dist"][off]).max()
        print("oracle check on %d individuals x %d sites: max rel err %.2e" % (len(ids), n_sites, rel))
        assert rel < 1e-9
g.close()
if world > 1:
    dist.destroy_process_group()
