#!/usr/bin/env python
"""bench.py -- pair-site evaluations per second of the ngsDist hot path on B200 (BASELINE.json metric).

Workload (every rank, every step): BASELINE.json configs[1] = "C2": synthetic GLs, 500 individuals x 100 000 sites,
`--probs --indep_geno --evol_model 2` (JC69).  One step = one pass of the hot path over that data set:
front end (K1) -> FP64 DMMA contraction (K2) -> split reduction + epilogue (K4) -> 500 x 500 matrix on the host.
`--indep_geno` selects the contraction north_star names; the literal default (per pair-site EM) is reported
next to it under "em_path".

  value : inputs (raw GLs, 1.2 GB) resident in HBM when the timed region starts; CUDA events on the library's stream.
  e2e   : same metric through the C ABI with HOST buffers: pinned raw -> ngsd_push_sites (H2D inside) -> ngsd_distances
          -> matrix in host memory.
  N > 1 : one process per GPU (torchrun), every rank runs the same step on its own data set (job / replicate level
          sharding, no data-path collective: "weak"); value = pair-sites of all ranks / max-over-ranks time.
  --impl reference : the UNMODIFIED reference (oracle/_ref/ngsDist, built from /root/reference with the gsl_rng_taus
          shim) on the host cores, same config, bounded site sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IND = 500
N_SITES = 100_000
SEED = 20251018
METRIC = "pair_site_evals_per_sec"
UNIT = "pair-sites/s"
WORKLOAD = "C2: synthetic GL 500 ind x 100k sites, --probs --indep_geno --evol_model 2 (JC69)"
REF_FLAGS = ["--probs", "--indep_geno", "--evol_model", "2"]


def pairs(n):
    return n * (n - 1) // 2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8:
                self.rows.append(f)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k] == "Active" for r in self.rows)]
        pw = [float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ reference arm --

def run_reference_sample(n_ind, n_sites, threads, workdir, flags=REF_FLAGS):
    """Hot-path seconds of the unmodified reference on n_ind x n_sites synthetic GLs: whole-process wall time minus a
    load-only run of the same file read as 2 individuals (same reader work, 1 pair)."""
    import oracle
    path = os.path.join(workdir, "ref_%dx%d.bin" % (n_ind, n_sites))
    if not os.path.exists(path):
        oracle.synth_raw(SEED, 0.0, n_ind, n_sites).tofile(path)

    def run(ni, ns):
        out = os.path.join(workdir, "ref.dist")
        cmd = [oracle.REF_BIN, "--geno", path, "--n_ind", str(ni), "--n_sites", str(ns), "--out", out, "--n_threads", str(threads),
               "--verbose", "0"] + flags
        t0 = time.perf_counter()
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t0

    wall = run(n_ind, n_sites)
    load = run(2, n_ind * n_sites // 2)
    return max(wall - load, 1e-6), wall, load


def reference_arm(args, rank):
    import oracle
    if rank != 0:
        return
    if not oracle.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ngsDist not built (needs /root/reference at build time)"}))
        return
    threads = os.cpu_count() or 1
    work = tempfile.mkdtemp(prefix="ngsd_bench_ref_")
    # size the per-step sample so that the whole run stays within a few minutes
    hot, wall, _ = run_reference_sample(N_IND, 500, threads, work)
    rate = pairs(N_IND) * 500 / hot
    budget_s = 150.0
    per_step_wall = budget_s / (args.steps + args.warmup)
    sites = int(max(500, min(20000, per_step_wall * 0.6 * rate / pairs(N_IND))) // 100 * 100)
    for _ in range(args.warmup):
        run_reference_sample(N_IND, sites, threads, work)
    hots = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hot, wall, load = run_reference_sample(N_IND, sites, threads, work)
        hots.append(hot)
    total = sum(hots)
    value = pairs(N_IND) * sites * args.steps / total
    sample = "%d ind x %d sites of the C2 data set per step (whole-process wall minus a load-only run of the same file)" % (N_IND, sites)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": {"workload": WORKLOAD, "sample_sites": sites, "threads": threads},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------ our arm --

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--core-only", action="store_true", help="skip the em_path / called_path extras (profiling runs)")
    ap.add_argument("--n-ind", type=int, default=N_IND)
    ap.add_argument("--n-sites", type=int, default=N_SITES)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import ngsdist_b200 as nb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_ind, n_sites = args.n_ind, args.n_sites
    W = max(args.warmup, 3)

    p = nb.Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=True, evol_model=2)
    g = nb.NgsDistB200(p, device=local)
    raw_dev = torch.empty((n_sites, n_ind, 3), dtype=torch.float64, device="cuda")
    g.synth_raw_device(raw_dev.data_ptr(), SEED + rank, 0.0, 0, n_sites)
    out_pin = torch.empty((n_ind, n_ind), dtype=torch.float64).pin_memory()
    stream = torch.cuda.ExternalStream(g.stream(), device=torch.device("cuda", local))

    launches = [0]
    t_dist, t_front, t_epi = [], [], []

    def step_resident():
        g.push_sites_device(raw_dev.data_ptr(), 0, n_sites)
        g.distances_raw(None, 0, 1, out_pin.data_ptr())
        t = g.timing()
        launches[0] += t.launches + 1          # + the front-end launch of the push
        t_dist.append(t.dist_ms); t_epi.append(t.epilogue_ms)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(W):
        step_resident()
    launches[0] = 0
    t_dist.clear(); t_epi.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    n_launch = launches[0]
    dist_ms = statistics.mean(t_dist)
    tim = g.timing()

    # ---- e2e: host buffers through the C ABI ----
    raw_pin = torch.empty((n_sites, n_ind, 3), dtype=torch.float64).pin_memory()
    raw_pin.copy_(raw_dev)

    def step_e2e():
        g.push_sites_ptr(raw_pin.data_ptr(), 0, n_sites)
        g.distances_raw(None, 0, 1, out_pin.data_ptr())

    for _ in range(2):
        step_e2e()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed(step_e2e, e2e_steps)

    # ---- the literal default of the reference for this config (no --indep_geno): per pair-site EM (K2b) ----
    em = None
    if rank == 0 and world == 1 and not args.core_only:
        pe = nb.Params(n_ind=n_ind, n_sites=n_sites, in_probs=True, indep_geno=False, evol_model=2)
        ge = nb.NgsDistB200(pe, device=local)
        ge.push_sites_device(raw_dev.data_ptr(), 0, n_sites)
        ge.distances_raw(None, 0, 1, out_pin.data_ptr())          # warm-up
        em_ms = []
        for _ in range(3):
            ge.distances_raw(None, 0, 1, out_pin.data_ptr())
            em_ms.append(ge.timing().total_ms)
        ge.close()
        em = {"workload": "same data, default --probs (no --indep_geno): per pair-site EM (emOptim2.cpp em2), kernel k_dist_em",
              "ms_per_matrix": statistics.median(em_ms), "value": pairs(n_ind) * n_sites / (statistics.median(em_ms) * 1e-3), "unit": UNIT}

    # ---- called genotypes (BASELINE configs[3] geometry, reduced site count): exact int8 tensor-core contraction (K2c) ----
    called = None
    if rank == 0 and world == 1 and (n_ind, n_sites) == (N_IND, N_SITES) and not args.core_only:
        cn, cs = 5000, 100_000
        pc = nb.Params(n_ind=cn, n_sites=cs, in_probs=True, call_geno=True, pairwise_del=False, evol_model=0)
        gc = nb.NgsDistB200(pc, device=local)
        chunk = 4096
        buf = torch.empty((chunk, cn, 3), dtype=torch.float64, device="cuda")
        fe_ms = 0.0
        for s0 in range(0, cs, chunk):
            m = min(chunk, cs - s0)
            gc.synth_raw_device(buf.data_ptr(), SEED, 0.05, s0, m)
            gc.push_sites_device(buf.data_ptr(), s0, m)
            fe_ms += gc.timing().frontend_ms
        del buf
        gc.frontend()
        out_c = torch.empty((cn, cn), dtype=torch.float64).pin_memory()
        gc.distances_raw(None, 0, 1, out_c.data_ptr())          # warm-up
        c_ms = []
        for _ in range(3):
            gc.distances_raw(None, 0, 1, out_c.data_ptr())
            tc = gc.timing()
            c_ms.append(tc.dist_ms)
        umma = not os.environ.get("NGSD_IMMA_SYNC")
        imma_peak = nb.probe_umma_tmacs(local) if umma else nb.probe_int8_tmacs(local)
        k_ms = statistics.median(c_ms)
        exe = tc.dist_imma * 4096.0 / (k_ms * 1e-3) * 1e-12
        called = {"workload": "C4 geometry at 1/50 of the sites: %d ind x %d sites, 5 %% missing, --call_geno (integer path, bit-exact), kernel %s" % (cn, cs, "k_dist_umma" if umma else "k_dist_imma"),
                  "kernel_ms": k_ms, "value": pairs(cn) * cs / (k_ms * 1e-3), "unit": UNIT,
                  "frontend_ms": fe_ms, "frontend_gbs_raw": cn * cs * 24 / (fe_ms * 1e-3) * 1e-9,
                  "roofline": {"bound": "tensor", "kernel": "k_dist_umma (tcgen05.mma kind::i8, UTCIMMA)" if umma else "k_dist_imma (mma.sync int8 IMMA.16832)",
                               "achieved": exe, "peak": imma_peak, "unit": "TMAC/s", "frac": exe / imma_peak,
                               "peak_source": "live back-to-back %s issue-rate probe in this run" % ("tcgen05.mma 128x128x32" if umma else "IMMA.16832"),
                               "note": "4 int8 MAC per pair-site (one-hot code x table column); the operands are expanded on chip from 2-bit codes: "
                                       "the tcgen05 path is bound by the on-chip expansion (A rows into TMEM, B rows into shared memory; ncu: tensor pipe 66 %, LSU wavefronts 67 %, integer ALU 62 %), not by the tensor pipe"}}
        gc.close()
        # the same shape end to end from HOST memory: 2-bit packed genotypes (ngsd_push_packed_genotypes, 0.25 B per
        # individual-site over PCIe) -> front end -> contraction -> D2H of the matrix
        try:
            import numpy as np
            stride = (cn + 3) // 4
            host = torch.from_numpy(np.random.RandomState(SEED & 0xFFFF).randint(0, 256, size=(cs, stride), dtype=np.uint8)).pin_memory()
            pg = nb.Params(n_ind=cn, n_sites=cs, in_probs=False, indep_geno=True, pairwise_del=False, evol_model=0)
            best = None
            gp = nb.NgsDistB200(pg, device=local)
            for it in range(4):                                   # first pass allocates the staging and result buffers
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                gp.push_packed_genotypes(host.numpy())
                gp.frontend()
                gp.distances_raw(None, 0, 1, out_c.data_ptr())
                dt = time.perf_counter() - t0
                if it > 0:
                    best = dt if best is None else min(best, dt)
            gp.close()
            called["e2e_packed"] = {"workload": "%d ind x %d sites as 2-bit genotypes in pinned host memory (25 %% missing), push + front end + contraction + D2H" % (cn, cs),
                                    "ms": best * 1e3, "value": pairs(cn) * cs / best, "unit": UNIT, "h2d_bytes": int(host.numel()), "d2h_bytes": cn * cn * 8}
        except Exception as ex:                                   # never lose the main line over an extra
            called["e2e_packed"] = {"error": str(ex)[:200]}
        del out_c

    # ---- bootstrap replicates (BASELINE configs[2] geometry at 1/10 of the sites): the graded per-replicate weighted
    #      contraction (diag(w_r) GEMMs) and, next to it, the block cache (per-block partials contracted once) ----
    boot = None
    if rank == 0 and world == 1 and (n_ind, n_sites) == (N_IND, N_SITES) and not args.core_only:
        bn, bsites, bblock = 2000, 100_000, 1000
        rows = {}
        for mode, nocache in (("weighted_contraction", True), ("block_cache", False)):
            pb = nb.Params(n_ind=bn, n_sites=bsites, in_probs=True, indep_geno=True, pairwise_del=True, evol_model=1, n_boot_rep=4,
                           boot_block_size=bblock, seed=12345, no_block_cache=nocache)
            gb = nb.NgsDistB200(pb, device=local)
            chunk = 8192
            buf = torch.empty((chunk, bn, 3), dtype=torch.float64, device="cuda")
            for s0 in range(0, bsites, chunk):
                m = min(chunk, bsites - s0)
                gb.synth_raw_device(buf.data_ptr(), SEED, 0.10, s0, m)
                gb.push_sites_device(buf.data_ptr(), s0, m)
            del buf
            gb.frontend()
            out_b = torch.empty((bn, bn), dtype=torch.float64).pin_memory()
            gb.distances_raw(None, 0, 1, out_b.data_ptr())
            ms = []
            for rep in range(4):
                counts, bs_ = gb.next_boot_counts()
                gb.distances_raw(counts.ctypes.data, len(counts), bs_, out_b.data_ptr())
                tb = gb.timing()
                ms.append((tb.total_ms, tb.block_cache))
            gb.close()
            del out_b
            steady = [m for m, flag in ms if flag != 1]
            rows[mode] = {"ms_per_replicate": statistics.median(steady), "first_call_ms": ms[0][0],
                          "value": pairs(bn) * bsites / (statistics.median(steady) * 1e-3), "unit": UNIT}
        boot = {"workload": "C3 geometry at 1/10 of the sites: %d ind x %d sites, 10 %% missing, --pairwise_del, block %d; nominal pair-sites per replicate / device time"
                            % (bn, bsites, bblock),
                "weighted_contraction": rows["weighted_contraction"], "block_cache": rows["block_cache"],
                "note": "block_cache is effective throughput (per-block partials are contracted once by the first replicate); the roofline above is the weighted contraction"}

    units_step = pairs(n_ind) * n_sites                # nominal pair-site evaluations per step per rank
    value = units_step * world * args.steps / (ms_total * 1e-3)
    e2e_value = units_step * world * e2e_steps / (ms_e2e * 1e-3)

    if rank == 0:
        peak = nb.probe_fp64_tflops(local)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tr = json.load(fh)["k_dist_dmma_c2"]
            if (n_ind, n_sites) == (N_IND, N_SITES):
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except Exception:
            pass
        # SURVEY §8(d) accounting rule: (i) roofline.achieved = FP64 tensor FLOP/s EXECUTED by the kernel (DMMA.8x8x4 count
        # x 512, including diagonal-tile waste) against the measured DMMA peak; (ii) `useful` = nominal pair-sites/s x 6 FLOP
        # (3 FMA per pair-site, ngsDist.cpp:351-353) against the same peak.  Without --pairwise_del the kernel uses the
        # sum-to-one reduction (2 FMA per pair-site), so (ii) may exceed (i) and even 1.0; it is never the roofline fraction.
        alg_flops = 6.0 * units_step
        useful = alg_flops / (dist_ms * 1e-3) * 1e-12
        executed = tim.dist_dmma * 512.0 / (dist_ms * 1e-3) * 1e-12
        achieved = executed
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD if (n_ind, n_sites) == (N_IND, N_SITES) else "synthetic GL %d x %d --indep_geno -m 2" % (n_ind, n_sites),
                       "n_ind": n_ind, "n_sites": n_sites, "pairs": pairs(n_ind), "l2": "inputs_larger_than_l2 (1.2 GB raw, 1.6 GB packed operands + 0.4 GB B2 plane)",
                       "sharding": "one independent job per GPU, no collective", "timed": "front end + contraction + epilogue + D2H of the matrix"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_sites * n_ind * 24, "d2h_bytes_per_step": n_ind * n_ind * 8,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": n_launch,
            "roofline": {"bound": "tensor", "kernel": "k_dist_dmma (FP64 DMMA.8x8x4 contraction)", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_note": "DRAM bytes read+written by one k_dist_dmma launch (ncu --set full, profiles/r01_ncu_summary.md); operands are 1.6 GB",
                         "peak_source": "live DMMA.8x8x4 issue-rate probe in this run (MEASURED_PEAKS.json has no FP64 figure; cuBLAS Dgemm measured 35.5)",
                         "executed_flops_per_launch": tim.dist_dmma * 512.0, "kernel_ms": dist_ms,
                         "useful_tflops_at_6flop_per_pair_site": useful, "useful_frac": useful / peak,
                         "algorithmic_flops_per_launch": alg_flops,
                         "note": "achieved = executed DMMA FLOP/s; the 2-plane (sum-to-one) contraction does 4 FLOP per pair-site",
                         "step_share": {"dist_ms": dist_ms, "epilogue_ms": statistics.mean(t_epi), "step_ms": ms_total / args.steps}},
        }
        if em is not None:
            line["em_path"] = em
        if called is not None:
            line["called_path"] = called
        if boot is not None:
            line["bootstrap_path"] = boot
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            if oracle.have_ref():
                threads = os.cpu_count() or 1
                work = tempfile.mkdtemp(prefix="ngsd_bench_cpu_")
                sites = 8000
                hot, wall, load = run_reference_sample(N_IND, sites, threads, work)
                line["cpu_baseline"] = {"value": pairs(N_IND) * sites / hot, "unit": UNIT, "cores": threads, "kind": "reference",
                                        "sample": "oracle/_ref/ngsDist --n_threads %d on %d ind x %d sites of the C2 data set: %.1f s wall - %.1f s load-only"
                                                  % (threads, N_IND, sites, wall, load)}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref/ngsDist missing"}
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
