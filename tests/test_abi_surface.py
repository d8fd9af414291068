"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every entry point that
include/ngsdist_b200.h declares (and nothing the header does not know), links NCCL itself, and fails loudly -- no CPU
fallback -- when asked to compute without a device."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ngsdist_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"NGSD_API\s+[\w\s\*]+?\b(ngsd_\w+)\s*\(", text)))


def test_library_exports_exactly_the_declared_entry_points():
    import ngsdist_b200 as nb
    L = nb.lib()
    decl = declared_symbols()
    assert len(decl) >= 30, decl
    missing = [s for s in decl if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(nb.ABI_SYMBOLS) == decl, (set(decl) ^ set(nb.ABI_SYMBOLS))
    out = subprocess.check_output(["nm", "-D", "--defined-only", nb.lib_path()], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("ngsd_"))
    assert exported == decl, (set(exported) ^ set(decl))
    assert L.ngsd_abi_version() == nb.api.ABI_VERSION == int(re.search(r"#define NGSD_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_library_links_nccl_directly():
    import ngsdist_b200 as nb
    out = subprocess.check_output(["readelf", "-d", nb.lib_path()], text=True)
    assert "libnccl.so.2" in out


def test_cfg_struct_matches_header_layout():
    import ngsdist_b200 as nb
    # uint64 x3, double[9], int32 x4, double x2, int32 x6, uint64: 8*3 + 72 + 16 + 16 + 24 + 8 = 160 bytes
    assert ctypes.sizeof(nb.api._Cfg) == 160
    cfg = nb.api._Cfg()
    nb.lib().ngsd_default_cfg(ctypes.byref(cfg))
    assert list(cfg.score) == [0, 0.5, 1, 0.5, 0, 0.5, 1, 0.5, 0] and cfg.evol_model == 1 and cfg.n_gpus == 0 and cfg.shard == 0


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ngsdist_b200 as nb
    with pytest.raises(nb.NgsDistError) as e:
        nb.NgsDistB200(nb.Params(n_ind=4, n_sites=64, indep_geno=True))
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(nb.NgsDistError):
        nb.NgsDistB200(nb.Params(n_ind=4, n_sites=64, indep_geno=True), n_gpus=2)


def test_argument_errors_are_reported_before_any_device_work():
    import ngsdist_b200 as nb
    for kw in (dict(n_ind=0, n_sites=64), dict(n_ind=4, n_sites=0), dict(n_ind=4, n_sites=64, evol_model=3),
               dict(n_ind=4, n_sites=64, tot_sites=10, pairwise_del=True), dict(n_ind=4, n_sites=(1 << 32))):
        with pytest.raises(nb.NgsDistError):
            nb.NgsDistB200(nb.Params(indep_geno=True, **kw))


def test_blank_site_marker_and_packing_helpers():
    import numpy as np
    import ngsdist_b200 as nb
    assert np.isnan(nb.BLANK_SITE) and np.array([nb.BLANK_SITE]).view(np.uint64)[0] == 0x7FF84E4753444231
    q = np.array([[[333340, 333330, 333330], [1048575, 0, 7]]])
    w = nb.pack_u20x3(q)
    assert w.shape == (1, 2) and w[0, 0] == 333340 | (333330 << 20) | (333330 << 40) and w[0, 1] == 1048575 | (7 << 40)
    assert float("%.6f" % 0.33334) == 333340 / 1e6          # the identity the fixed-point tier rests on
