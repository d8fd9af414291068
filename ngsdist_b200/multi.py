"""Shard plans of the multi-GPU hot path, and their CPU-testable mirror.

The multi-GPU orchestration itself lives below the C ABI (csrc/comm.cu: `ngsd_cfg.n_gpus`, `ngsd_comm_attach`,
`ngsd_distances_batch`, `ngsd_comm_reduce_sites`, `ngsd_comm_reduce_tiles`, `ngsd_comm_allgather_operands`, NCCL linked
into the library).  The reference has exactly one parallel strategy (a task per pair through a pthread pool,
ngsDist.cpp:197-269); on a box of B200s the path shards three ways:

  replicates : replicate r on rank r % world; every rank advances the SAME host RNG stream (gsl_rng_taus), so the block
               draws are those of a single-process run.
  tiles      : the 128 x 128 upper-triangle tiles dealt to the ranks (ngsd_set_tile_shard); entries a rank does not own
               are 0, so one SUM assembles the matrix exactly.
  sites      : rank g owns a contiguous block-aligned site range; raw sums num (FP64) and cnt (int64) are reduced and the
               non-linear tail of gen_dist (ngsDist.cpp:372-401) runs after the reduction.

This module holds what a HOST needs around those calls -- the plans (`site_shards`, `slice_block_counts`,
`replicate_shard`, `tile_owner_mask`, `BootStream`) used by bench.py and tools/ -- and a torch.distributed restatement of
the three collectives' wiring (`run_replicates`, `run_tiles`, `reduce_site_partials`) against compute callbacks, so that
the plans and the assembly rules are exercised by world_size-2 `gloo` tests on CPU (tests/test_multi_cpu.py) where
neither CUDA nor NCCL exists.  On GPUs the product path is the C ABI (tools/multi_gpu_check.py, tools/group_check.py).
"""
import numpy as np


# ----------------------------------------------------------------------------------------------- shard plans --

def replicate_shard(n_matrices, rank, world):
    """Indices (0 = full data set, r >= 1 = bootstrap replicate r) computed by `rank`."""
    return [r for r in range(n_matrices) if r % world == rank]


def site_shards(n_sites, block_size, world):
    """Contiguous [s0, s1) ranges, one per rank, aligned to bootstrap blocks (and to 64 sites when block_size allows
    nothing better is needed: every context numbers its own sites from 0).  The last shard takes the remainder,
    including the n_sites % block_size sites that only replicate 0 uses (ngsDist.cpp:236)."""
    n_blocks = n_sites // block_size
    out, b0 = [], 0
    for g in range(world):
        b1 = (n_blocks * (g + 1)) // world
        s0, s1 = b0 * block_size, b1 * block_size
        if g == world - 1:
            s1 = n_sites
        out.append((s0, s1))
        b0 = b1
    return out


def slice_block_counts(counts, shard, block_size):
    """Block multiplicities of the blocks that lie inside `shard` = (s0, s1); s0 is block aligned."""
    s0, s1 = shard
    b0 = s0 // block_size
    b1 = min(len(counts), s1 // block_size)
    return np.ascontiguousarray(counts[b0:b1])


def tile_list(n_ind, band=12):
    """Upper-triangle 128 x 128 tile list in the library's order (csrc/api.cu make_tiles): (ti, tj) with ti <= tj."""
    RB = (n_ind + 127) // 128
    t = []
    for bi in range(0, RB, band):
        for bj in range(bi, RB, band):
            for ti in range(bi, min(bi + band, RB)):
                for tj in range(max(bj, ti), min(bj + band, RB)):
                    t.append((ti, tj))
    return t


def tile_owner_mask(n_ind, rank, world):
    """Boolean n x n mask (upper and lower triangle, no diagonal) of the entries `rank` owns under ngsd_set_tile_shard."""
    m = np.zeros((n_ind, n_ind), dtype=bool)
    tiles = tile_list(n_ind)
    owner, held, t = [], [0] * world, 0
    while t < len(tiles):                       # ownership goes by pairs of neighbouring tiles of one row block,
        ln = 2 if t + 1 < len(tiles) and tiles[t + 1][0] == tiles[t][0] else 1
        to = held.index(min(held))              # each to the rank holding the fewest tiles so far (ngsd_set_tile_shard)
        held[to] += ln
        owner += [to] * ln
        t += ln
    for k, (ti, tj) in enumerate(tiles):
        if owner[k] != rank:
            continue
        i0, i1, j0, j1 = ti * 128, min((ti + 1) * 128, n_ind), tj * 128, min((tj + 1) * 128, n_ind)
        blk = np.zeros((i1 - i0, j1 - j0), dtype=bool)
        ii, jj = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), indexing="ij")
        blk[ii < jj] = True
        m[i0:i1, j0:j1] |= blk
    return m | m.T


class BootStream:
    """The replicate loop's host RNG bookkeeping (ngsDist.cpp:217-238), identical on every rank."""

    def __init__(self, n_sites, block_size, seed):
        from . import api
        self._api = api
        self.n_sites = n_sites
        self.bs = block_size
        self.state = np.zeros(3, dtype=np.uint32)
        api.lib().ngsd_taus_seed(api._ptr(self.state), int(seed) & 0xFFFFFFFF)

    def next_counts(self):
        self.n_sites -= self.n_sites % self.bs
        return self._api.taus_block_counts(self.state, self.n_sites // self.bs)


# ------------------------------------------------------------------------------------------- device tensors --

class _DevBuf:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_tensor(ptr, shape, typestr):
    """A torch CUDA tensor aliasing a library-owned device buffer (no copy): '<f8' or '<i8'."""
    import torch
    return torch.as_tensor(_DevBuf(ptr, shape, typestr), device="cuda")


def _to_tensor(a, backend):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.cuda() if backend == "nccl" else t


# --------------------------------------------------------------------------------------------- orchestrations --

def run_replicates(n_boot_rep, boot: BootStream, compute, rank, world, group=None):
    """compute(rep, counts_or_None, block_size) -> n x n float64 distance matrix.
    Returns the list of all 1 + n_boot_rep matrices on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    mine = {}
    for rep in range(n_boot_rep + 1):
        counts = boot.next_counts() if rep > 0 else None       # every rank draws every replicate: same stream
        if rep % world == rank:
            mine[rep] = np.ascontiguousarray(compute(rep, counts, boot.bs))
    if world == 1:
        return [mine[r] for r in range(n_boot_rep + 1)]
    backend = dist.get_backend(group)
    out = [None] * (n_boot_rep + 1) if rank == 0 else None
    # replicate r lives on rank r % world; rounds of `world` matrices are gathered to rank 0
    n = next(iter(mine.values())).shape[0] if mine else None
    shape = torch.tensor([n if n is not None else 0], dtype=torch.int64)
    shape = shape.cuda() if backend == "nccl" else shape
    dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=group)
    n = int(shape.item())
    for base in range(0, n_boot_rep + 1, world):
        rep = base + rank
        have = rep <= n_boot_rep
        t = _to_tensor(mine[rep] if have else np.zeros((n, n)), backend)
        bucket = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, bucket, dst=0, group=group)
        if rank == 0:
            for g in range(world):
                if base + g <= n_boot_rep:
                    out[base + g] = bucket[g].cpu().numpy()
    return out


def run_tiles(compute_owned, rank, world, group=None):
    """compute_owned() -> dict(dist=, num=, cnt=) with zeros outside the rank's tiles.  All ranks get the full matrices."""
    import torch.distributed as dist
    res = compute_owned()
    if world == 1:
        return res
    backend = dist.get_backend(group)
    out = {}
    for k, v in res.items():
        t = _to_tensor(v.astype(np.int64) if v.dtype == np.uint64 else v, backend)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        a = t.cpu().numpy()
        out[k] = a.astype(np.uint64) if v.dtype == np.uint64 else a
    return out


def reduce_site_partials(num, cnt, group=None):
    """In-place SUM all-reduce of raw sums.  `num`/`cnt` are torch tensors (CUDA tensors aliasing the library's buffers
    with NCCL, CPU tensors with gloo)."""
    import torch.distributed as dist
    dist.all_reduce(num, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return num, cnt
