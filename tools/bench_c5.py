"""C5 (BASELINE.json configs[4]): 20 000 individuals x 10 000 000 sites, --avg_nuc_dist --indep_geno, sites sharded over
the ranks (one per GPU, torchrun), raw sums reduced by ONE ncclReduce of the packed upper triangle issued by the library
(ngsd_comm_reduce_sites), epilogue on the root.  4.8 TB of raw GLs never exist: every rank generates its sites chunk by
chunk on the device (SURVEY §8(d) generator).  --check validates the 20 000-individual geometry against the CPU oracle on a
small site range and the size-independent properties of the full matrix."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import ngsdist_b200 as nb
from ngsdist_b200 import multi

ap = argparse.ArgumentParser()
ap.add_argument("--n-ind", type=int, default=20000)
ap.add_argument("--n-sites", type=int, default=10_000_000)
ap.add_argument("--chunk", type=int, default=2048)
ap.add_argument("--slab", type=int, default=98_304, help="sites contracted per ngsd_distances call (operands of a slab stay resident)")
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
n, S = args.n_ind, args.n_sites
shards = multi.site_shards(S, 192, world)
s0, s1 = shards[rank]
# The packed operands of a rank's 1.25 M sites (40 B per individual-site = 1 TB) do not fit either: the shard is walked in
# slabs; each slab is front-ended, contracted (raw sums stay on the device) and added to the running sums by the library's
# own linearity: num is accumulated on the host side of this tool in a device tensor.
slab = min(args.slab, s1 - s0) // 192 * 192
mkp = lambda m: nb.Params(n_ind=n, n_sites=m, in_probs=True, indep_geno=True, avg_nuc_dist=True, evol_model=1)
buf = torch.empty((args.chunk, n, 3), dtype=torch.float64, device="cuda")
acc = torch.zeros((n, n), dtype=torch.float64, device="cuda")
fe_ms = k_ms = 0.0
dmma = 0


def run_slab(ctx, a, m_slab):
    global fe_ms, k_ms, dmma
    for c0 in range(0, m_slab, args.chunk):
        m = min(args.chunk, m_slab - c0)
        ctx.synth_raw_device(buf.data_ptr(), 20251018, 0.0, a + c0, m)   # global site index: the same data set on any world size
        ctx.push_sites_device(buf.data_ptr(), c0, m)
        fe_ms += ctx.timing().frontend_ms
    ctx.partial_sums(None, 1)
    t = ctx.timing()
    k_ms += t.dist_ms
    dmma += t.dist_dmma
    _, num_ptr, _ = ctx.device_results()
    acc.add_(multi.device_tensor(num_ptr, (n, n), "<f8"))
    torch.cuda.synchronize()                              # the library reuses its buffers for the next slab


if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.time()
n_full = (s1 - s0) // slab
tail = (s1 - s0) - n_full * slab
if tail:                                                  # the shorter last slab first, in its own context (freed before the big one exists)
    tc = nb.NgsDistB200(mkp(tail), device=local)
    run_slab(tc, s0 + n_full * slab, tail)
    tc.close()
    torch.cuda.empty_cache()
g = nb.NgsDistB200(mkp(slab), device=local)
t_attach = 0.0
if world > 1:                                             # communicator set-up (ncclCommInitRank + channel connection) is not part of the job
    torch.cuda.synchronize()
    ta = time.time()
    box = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    g.comm_attach(box[0], rank, world)
    t_attach = time.time() - ta
for k in range(n_full):
    run_slab(g, s0 + k * slab, slab)
torch.cuda.synchronize()
t_local = time.time() - t0 - t_attach
# hand the accumulated sums back to the library buffer and reduce: ONE collective for the whole job
_, num_ptr, _ = g.device_results()
multi.device_tensor(num_ptr, (n, n), "<f8").copy_(acc)
torch.cuda.synchronize()
out = torch.empty((n, n), dtype=torch.float64, pin_memory=True) if rank == 0 else None
t1 = time.time()
g._check(nb.lib().ngsd_comm_reduce_sites(g._h, 0, S, out.data_ptr() if rank == 0 else None)) if world > 1 else g.finish(out.numpy())
t_red = time.time() - t1
if world > 1:
    dist.barrier()
t_all = time.time() - t0 - t_attach
if rank == 0:
    pairs = n * (n - 1) // 2
    rb, rms = g.comm_stats() if world > 1 else (0, 0.0)
    rep = dict(workload="C5: %d ind x %d sites, --avg_nuc_dist --indep_geno, sites sharded over %d ranks" % (n, S, world), world=world,
               sites_per_rank=s1 - s0, slab_sites=slab, local_phase_s=t_local, frontend_kernels_ms=fe_ms, contraction_kernels_ms=k_ms,
               dmma_tflops_per_gpu=dmma * 512 / (k_ms * 1e-3) * 1e-12, reduce_and_epilogue_s=t_red, reduce_bytes=rb, reduce_device_ms=rms,
               full_num_cnt_matrices_bytes=n * n * 16, comm_setup_s_excluded=t_attach, job_s=t_all, pair_sites_per_s=pairs * S / t_all)
    o = out.numpy()
    rep["symmetric"] = bool(np.array_equal(o[:512, :512], o[:512, :512].T) and np.array_equal(o[0, :], o[:, 0]))
    rep["diag_zero"] = bool((np.diag(o) == 0).all())
    rep["finite"] = bool(np.isfinite(o).all())
    rep["checksum_row0"] = float(o[0, 1:9].sum())
    print(json.dumps(rep))
    if args.check:
        import oracle
        m = 4096
        gc = nb.NgsDistB200(nb.Params(n_ind=n, n_sites=m, in_probs=True, indep_geno=True, avg_nuc_dist=True, evol_model=1), device=local)
        b2 = torch.empty((m, n, 3), dtype=torch.float64, device="cuda")
        gc.synth_raw_device(b2.data_ptr(), 20251018, 0.0, 0, m)
        gc.push_sites_device(b2.data_ptr(), 0, m)
        small = gc.distances()["dist"]
        gc.close()
        ids = np.concatenate([np.arange(0, 64), np.arange(n - 64, n)])
        raw = oracle.synth_raw(20251018, 0.0, n, m)[:, ids, :]
        ora = oracle.distances(oracle.frontend(raw), score=oracle.score_matrix(True), indep=True, evol_model=1)["dist"]
        got = small[np.ix_(ids, ids)]
        off = ~np.eye(len(ids), dtype=bool)
        rel = np.abs(got[off] - ora[off]).max() / np.abs(ora[off]).max()
        print("oracle check, 20000-individual geometry, 128 individuals x %d sites: max rel err %.2e" % (m, rel))
        assert rel < 1e-9
g.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
