"""Plain cudaMemcpyAsync host->device probe, one process per GPU (torchrun): what N ranks pushing from pinned memory at
the same time get from this box -- the ceiling of bench.py's end-to-end number (0.55 / 0.45 e2e efficiency
at N = 4 / 8 with every pinned buffer on NUMA node 0).  Each rank copies a 1 GiB pinned buffer 8 times; with --bind the
rank first binds itself and its pages to its GPU's NUMA node (ngsd_bind_host_to_device)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
node = None
if "--bind" in sys.argv:
    import ngsdist_b200 as nb
    node = nb.bind_host_to_device(local)
if world > 1:
    dist.init_process_group("gloo")
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = 8 * n / (e0.elapsed_time(e1) * 1e-3) * 1e-9
if world > 1:
    lst = [None] * world
    dist.all_gather_object(lst, gbs)
else:
    lst = [gbs]
if rank == 0:
    nodes = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")) if os.path.isdir("/sys/devices/system/node") else []
    print(json.dumps({"ranks": world, "bind": "--bind" in sys.argv, "numa_nodes_visible": nodes, "numa_node_rank0": node,
                      "h2d_gbs_per_rank": [round(x, 1) for x in lst], "h2d_gbs_total": round(sum(lst), 1), "cpus": os.cpu_count()}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
