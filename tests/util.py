"""Shared helpers for the parity tests (host-side only; no product code here)."""
import gzip
import json
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


def load_bin(name, n_ind, n_sites):
    return np.fromfile(os.path.join(GOLDEN, name), dtype=np.float64).reshape(n_sites, n_ind, 3)


def golden_text(name):
    with open(os.path.join(GOLDEN, name + ".dist")) as fh:
        return fh.read()


def fmt_value(v):
    """glibc "%.10f" including its spelling of nan / -nan / inf."""
    if math.isnan(v):
        return "-nan" if math.copysign(1.0, v) < 0 else "nan"
    if math.isinf(v):
        return "-inf" if v < 0 else "inf"
    return "%.10f" % v


def format_dist(mats, labels=None):
    """The writer of ngsDist.cpp:282-287: '\\n<n>\\n' then 'label\\tv0\\t...\\n' per row, per matrix."""
    out = []
    for m in mats:
        n = m.shape[0]
        lab = labels or ["Ind_%d" % i for i in range(n)]
        out.append("\n%d\n" % n)
        for i in range(n):
            out.append(lab[i] + "\t" + "\t".join(fmt_value(float(v)) for v in m[i]) + "\n")
    return "".join(out)


def parse_flags(flags):
    """CLI tokens of a manifest case -> keyword arguments shared by oracle.run_job and the product API."""
    kw = dict(in_log=False, call_geno=False, N_thresh=0.0, call_thresh=0.0, avg_nuc_dist=False, indep=False,
              pairwise_del=False, tot_sites=0, evol_model=1, n_boot_rep=0, boot_block_size=1, seed=12345)
    probs = False
    k = 0
    while k < len(flags):
        f = flags[k]
        if f == "--probs":
            probs = True
        elif f == "--log_scale":
            kw["in_log"] = True; probs = True
        elif f == "--indep_geno":
            kw["indep"] = True
        elif f == "--call_geno":
            kw["call_geno"] = True
        elif f == "--pairwise_del":
            kw["pairwise_del"] = True
        elif f == "--avg_nuc_dist":
            kw["avg_nuc_dist"] = True
        elif f in ("--N_thresh", "--call_thresh"):
            kw[f[2:]] = float(flags[k + 1]); kw["call_geno"] = True; k += 1
        elif f in ("--evol_model", "--tot_sites", "--n_boot_rep", "--boot_block_size", "--seed"):
            kw[f[2:]] = int(flags[k + 1]); k += 1
        else:
            raise ValueError(f)
        k += 1
    if not probs or kw["call_geno"]:
        kw["indep"] = True  # ngsDist.cpp:55-62
    return kw, probs


def read_text_input(name, n_ind, n_sites, probs, want_blank=False):
    """The text reader's semantics (read_data.cpp:48-103) for the well-formed golden inputs: header skipped when it has
    too few numeric fields, last n_ind*n_geno numeric columns used; an empty line consumes a site (:58-59) -- its row is a
    placeholder here and its index is returned in `blank` (the caller applies the reference's outcome for such a site)."""
    n_geno = 3 if probs else 1
    rows, blank = [], []
    with gzip.open(os.path.join(GOLDEN, name), "rt") as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line == "":
                if len(rows) < n_sites:
                    blank.append(len(rows))
                    rows.append([1.0 / 3] * (n_ind * n_geno) if probs else [-1.0] * n_ind)
                continue
            nums = []
            for tok in line.replace(" ", "\t").split("\t"):
                if tok == "":
                    continue
                try:
                    nums.append(float(tok))
                except ValueError:
                    pass
            if len(nums) < n_ind * n_geno:
                continue  # header
            rows.append(nums[-n_ind * n_geno:])
    a = np.array(rows[:n_sites], dtype=np.float64)
    a = a.reshape(n_sites, n_ind, 3) if probs else a.reshape(n_sites, n_ind).astype(np.int32)
    return (a, blank) if want_blank else a
