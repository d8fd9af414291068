"""Multi-GPU sharding modes on real GPUs, through the C ABI only (needs >= 2 devices; a 1-GPU box skips them).
The tools print what they verified; logs of the 2- and 8-GPU runs are under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_one_process_per_gpu_nccl_below_the_abi():
    """torchrun, one rank per GPU: ngsd_comm_attach + ngsd_distances_batch / reduce_tiles / reduce_sites / allgather."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_CHECK_OK" in r.stdout, r.stdout[-3000:]


def test_one_process_many_gpus_group_context():
    """ngsd_cfg.n_gpus = 2 in ONE process (REPLICATED and SITES) against a single-GPU context."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "group_check.py"), "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=900)
    assert r.returncode == 0 and "GROUP_CHECK_OK" in r.stdout, r.stdout[-3000:]


def test_group_context_rejects_more_gpus_than_visible():
    import ngsdist_b200 as nb
    p = nb.Params(n_ind=10, n_sites=640, indep_geno=True)
    with pytest.raises(nb.NgsDistError):
        nb.NgsDistB200(p, n_gpus=_n_gpus() + 1)


@pytest.mark.parametrize("shard", ["replicated", "sites"])
@pytest.mark.parametrize("name", ["c1_indep_boot", "c1_thresh_boot_pdel", "c1_call_boot", "c1_em"])
def test_cli_n_gpus_writes_the_reference_file(name, shard, tmp_path):
    """The drop-in command line with --n_gpus 2: same bytes as the reference's .dist (goldens written by oracle/_ref/ngsDist)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from util import GOLDEN, golden_text, manifest
    case = [c for c in manifest()["binary"] if c["name"] == name][0]
    out = str(tmp_path / "o.dist")
    cli = os.path.join(ROOT, "ngsdist_b200", "bin", "ngsDist")
    args = [cli, "--geno", os.path.join(GOLDEN, case["input"]), "--n_ind", str(case["n_ind"]), "--n_sites", str(case["n_sites"]), "--out", out,
            "--n_threads", "4", "--verbose", "0", "--n_gpus", "2", "--shard", shard] + case["flags"]
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(out).read() == golden_text(name)
