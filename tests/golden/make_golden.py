#!/usr/bin/env python
"""Generate the golden fixtures in this directory with the UNMODIFIED reference binary.

Run in the build container (needs /root/reference to build oracle/_ref/ngsDist):
    python tests/golden/make_golden.py
Writes the small inputs (*.bin: binary [site][ind][3] doubles; *.txt.gz: text), one reference `.dist`
per case and manifest.json.  The reference's own goldens (examples/test.md5) are unreachable because
their inputs are not shipped; these fixtures replace them.  The flag matrix mirrors
examples/test.sh (no bootstrap / --n_boot_rep 5 / + --boot_block_size 10 / + --call_geno /
+ --N_thresh 0.3 --call_thresh 0.9, --seed 12345) and adds the flags test.sh never covers.
"""
import gzip
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def make_inputs():
    inputs = {}
    # 7 x 53 with hand-made edge cases
    raw = oracle.synth_raw(20251018, 0.10, 7, 53)
    raw[3, 2] = (0.5, 0.5, 0.0)          # tie between two maxima + an exact zero
    raw[4, 5] = (0.0, 0.0, 1.0)          # already one-hot
    raw[5, :, :] = 1.0 / 3.0             # an all-missing site
    raw[10:30, 6, :] = 0.25              # an individual with a long missing stretch (un-normalised equal triple)
    raw[7, 1] = (0.900000, 0.05, 0.05)   # sits near --call_thresh 0.9 (App. E-7)
    raw[8, 1] = (0.2, 0.5, 0.3)          # below call_thresh, above N_thresh -> stays soft
    raw[9, 1] = (0.34, 0.33, 0.33)       # max below a 0.35 N_thresh
    inputs["g7x53"] = raw
    inputs["g24x400"] = oracle.synth_raw(20251018, 0.10, 24, 400)
    inputs["g5x23_nomiss"] = oracle.synth_raw(7, 0.0, 5, 23)
    lg = np.log(oracle.synth_raw(99, 0.0, 6, 31))
    lg[4, 3] = (-np.inf, 0.0, -np.inf)   # -inf survives a binary --log_scale read un-clamped (read_data.cpp:37-38)
    inputs["g6x31_log"] = lg
    # all-zero likelihood triples (binary, normal scale): exp(-1.125) three times in the reference, i.e. a triple that
    # does not sum to one -- as row and as column individual, next to ordinary data
    z = oracle.synth_raw(11, 0.05, 6, 40)
    z[3, 0] = 0.0; z[3, 4] = 0.0; z[17, 0] = 0.0; z[18, 2] = 0.0; z[39, 5] = 0.0; z[20, 1] = 0.0; z[21, 1] = 0.0
    inputs["g6x40_zero"] = z
    inputs["grid12x96"] = decimal_grid()
    return inputs


def decimal_grid(n_ind=12, n_sites=96):
    """6-decimal posteriors (what ANGSD -doGeno 8 prints) sitting ON the miss_data boundary |p0-p1| = |p1-p2| = 1e-5
    (gen_func.cpp:862-868, EPSILON = 1e-5) and one micro-unit either side of it: the outcome of the reference's
    comparison depends on the last bit of exp(log(x) - logsum), so cnt under --pairwise_del pins the front end's rounding."""
    rng = np.random.RandomState(20261018)
    g = np.empty((n_sites, n_ind, 3))
    for s in range(n_sites):
        for i in range(n_ind):
            kind = rng.randint(0, 4)
            if kind == 0:        # ordinary soft triple, 6 decimals
                x = rng.dirichlet((0.7, 0.7, 0.7))
                a, b = int(round(x[0] * 1e6)), int(round(x[1] * 1e6))
            else:                # near-uniform, differences of 8..12 micro-units in both comparisons
                d01 = int(rng.choice([-12, -11, -10, -9, -8, 8, 9, 10, 11, 12]))
                d12 = int(rng.choice([-11, -10, -9, 0, 3, 9, 10, 11]))
                # a - b = d01, b - c = d12, a + b + c = 1e6  ->  3 b = 1e6 - d01 + d12
                b = (1000000 - d01 + d12) // 3
                a = b + d01
            c = 1000000 - a - b
            g[s, i] = (float("%.6f" % (a * 1e-6)), float("%.6f" % (b * 1e-6)), float("%.6f" % (c * 1e-6)))
    return g


CASES = [
    # name, input, flags
    ("em_default", "g7x53", ["--probs"]),
    ("em_m0", "g7x53", ["--probs", "--evol_model", "0"]),
    ("indep_m0", "g7x53", ["--probs", "--indep_geno", "--evol_model", "0"]),
    ("indep_m1", "g7x53", ["--probs", "--indep_geno"]),
    ("indep_m2_avg", "g7x53", ["--probs", "--indep_geno", "--evol_model", "2", "--avg_nuc_dist"]),
    ("indep_pdel", "g7x53", ["--probs", "--indep_geno", "--pairwise_del", "--evol_model", "0"]),
    ("indep_tot", "g7x53", ["--probs", "--indep_geno", "--tot_sites", "100"]),
    ("call", "g7x53", ["--probs", "--call_geno"]),
    ("call_pdel_m2", "g7x53", ["--probs", "--call_geno", "--pairwise_del", "--evol_model", "2"]),
    ("call_thresh", "g7x53", ["--probs", "--N_thresh", "0.35", "--call_thresh", "0.9"]),
    ("em_boot", "g7x53", ["--probs", "--n_boot_rep", "5", "--boot_block_size", "10", "--seed", "12345"]),
    ("indep_boot_pdel", "g7x53", ["--probs", "--indep_geno", "--pairwise_del", "--n_boot_rep", "5", "--boot_block_size", "10", "--seed", "12345"]),
    ("call_boot_b7", "g7x53", ["--probs", "--call_geno", "--n_boot_rep", "5", "--boot_block_size", "7", "--seed", "12345"]),
    ("indep_boot_b1", "g7x53", ["--probs", "--indep_geno", "--n_boot_rep", "3", "--seed", "4242", "--evol_model", "0"]),
    ("c1_em", "g24x400", ["--probs"]),
    ("c1_indep_boot", "g24x400", ["--probs", "--indep_geno", "--n_boot_rep", "5", "--boot_block_size", "10", "--seed", "12345"]),
    ("c1_call_boot", "g24x400", ["--probs", "--call_geno", "--n_boot_rep", "5", "--boot_block_size", "10", "--seed", "12345"]),
    ("c1_thresh_boot_pdel", "g24x400", ["--probs", "--N_thresh", "0.3", "--call_thresh", "0.9", "--pairwise_del", "--n_boot_rep", "5", "--boot_block_size", "10", "--seed", "12345"]),
    ("c1_indep_pdel_m2", "g24x400", ["--probs", "--indep_geno", "--pairwise_del", "--evol_model", "2"]),
    ("nomiss_indep_m2", "g5x23_nomiss", ["--probs", "--indep_geno", "--evol_model", "2"]),
    ("log_indep", "g6x31_log", ["--probs", "--log_scale", "--indep_geno", "--evol_model", "0"]),
    ("log_em", "g6x31_log", ["--probs", "--log_scale"]),
    ("zero_indep", "g6x40_zero", ["--probs", "--indep_geno", "--evol_model", "0"]),
    ("zero_indep_boot", "g6x40_zero", ["--probs", "--indep_geno", "--n_boot_rep", "3", "--boot_block_size", "5", "--seed", "12345"]),
    ("zero_pdel", "g6x40_zero", ["--probs", "--indep_geno", "--pairwise_del", "--evol_model", "2"]),
    ("zero_em", "g6x40_zero", ["--probs", "--evol_model", "0"]),
    ("grid_bin_pdel", "grid12x96", ["--probs", "--indep_geno", "--pairwise_del", "--evol_model", "0"]),
    ("grid_bin_em_pdel", "grid12x96", ["--probs", "--pairwise_del"]),
    ("grid_bin_thresh", "grid12x96", ["--probs", "--N_thresh", "0.33334", "--call_thresh", "0.9", "--pairwise_del", "--evol_model", "0"]),
    ("boot_bigblock", "g5x23_nomiss", ["--probs", "--indep_geno", "--n_boot_rep", "2", "--boot_block_size", "50", "--seed", "12345"]),
]

TEXT_CASES = [
    # name, text input, n_ind, n_sites, flags
    ("txt_geno", "geno9x40.txt.gz", 9, 40, []),
    ("txt_geno_pdel_boot", "geno9x40.txt.gz", 9, 40, ["--pairwise_del", "--n_boot_rep", "2", "--boot_block_size", "5", "--seed", "12345", "--evol_model", "0"]),
    ("txt_probs_em", "probs9x40.txt.gz", 9, 40, ["--probs"]),
    ("txt_probs_call", "probs9x40.txt.gz", 9, 40, ["--probs", "--call_geno", "--evol_model", "2"]),
    # the decimal grid through the text reader (log() without the -inf clamp, read_data.cpp:83-99)
    ("txt_grid_pdel", "grid12x96.txt.gz", 12, 96, ["--probs", "--indep_geno", "--pairwise_del", "--evol_model", "0"]),
    ("txt_grid_em_pdel", "grid12x96.txt.gz", 12, 96, ["--probs", "--pairwise_del", "--evol_model", "2"]),
    # an empty line consumes a site and leaves (0,0,0) for every individual (read_data.cpp:58-59)
    ("txt_blank_indep", "probsblank9x40.txt.gz", 9, 40, ["--probs", "--indep_geno", "--evol_model", "0"]),
    ("txt_blank_pdel_boot", "probsblank9x40.txt.gz", 9, 40, ["--probs", "--indep_geno", "--pairwise_del", "--n_boot_rep", "2", "--boot_block_size", "5", "--seed", "12345"]),
    ("txt_blank_call", "probsblank9x40.txt.gz", 9, 40, ["--probs", "--call_geno"]),
    ("txt_blank_em", "probsblank9x40.txt.gz", 9, 40, ["--probs"]),
    ("txt_blank_em_pdel", "probsblank9x40.txt.gz", 9, 40, ["--probs", "--pairwise_del"]),
    ("txt_blank_geno", "genoblank9x40.txt.gz", 9, 40, ["--evol_model", "0"]),
    ("txt_blank_geno_pdel_boot", "genoblank9x40.txt.gz", 9, 40, ["--pairwise_del", "--n_boot_rep", "2", "--boot_block_size", "5", "--seed", "12345"]),
]


def main():
    oracle.build()
    assert oracle.have_ref(), "oracle/_ref/ngsDist missing"
    manifest = {"binary": [], "text": []}
    inputs = make_inputs()
    for name, raw in inputs.items():
        raw.tofile(os.path.join(HERE, name + ".bin"))
    for name, inp, flags in CASES:
        raw = inputs[inp]
        n_sites, n_ind, _ = raw.shape
        _, text = oracle.run_reference(None, flags, geno_path=os.path.join(HERE, inp + ".bin"), n_ind=n_ind, n_sites=n_sites)
        with open(os.path.join(HERE, name + ".dist"), "w") as fh:
            fh.write(text)
        manifest["binary"].append(dict(name=name, input=inp + ".bin", n_ind=n_ind, n_sites=n_sites, flags=flags))
    # text inputs: called genotypes (with -1 = missing) and 6-decimal posteriors with 3 leading label columns + header
    rng = np.random.RandomState(5)
    codes = rng.randint(-1, 3, size=(40, 9))
    with gzip.open(os.path.join(HERE, "geno9x40.txt.gz"), "wt") as fh:
        for s in range(40):
            fh.write("chr1\t%d\t" % (s + 1) + "\t".join(str(int(c)) for c in codes[s]) + "\n")
    pr = oracle.synth_raw(3, 0.1, 9, 40)
    pr /= pr.sum(axis=2, keepdims=True)
    with gzip.open(os.path.join(HERE, "probs9x40.txt.gz"), "wt") as fh:
        fh.write("marker\tallele1\tallele2\t" + "\t".join("Ind%d\tInd%d\tInd%d" % (i, i, i) for i in range(9)) + "\n")
        for s in range(40):
            fh.write("chr1_%d\tA\tC\t" % (s + 1) + "\t".join("%.6f" % v for v in pr[s].reshape(-1)) + "\n")
    grid = inputs["grid12x96"]
    with gzip.open(os.path.join(HERE, "grid12x96.txt.gz"), "wt") as fh:
        for srow in grid:
            fh.write("\t".join("%.6f" % v for v in srow.reshape(-1)) + "\n")
    # the same two files with sites 7 and 23 replaced by empty lines
    for src, dst in (("probs9x40.txt.gz", "probsblank9x40.txt.gz"), ("geno9x40.txt.gz", "genoblank9x40.txt.gz")):
        lines = gzip.open(os.path.join(HERE, src), "rt").read().split("\n")
        first = 1 if src.startswith("probs") else 0          # header line
        for s_blank in (7, 23):
            lines[first + s_blank] = ""
        with gzip.open(os.path.join(HERE, dst), "wt") as fh:
            fh.write("\n".join(lines))
    for name, inp, n_ind, n_sites, flags in TEXT_CASES:
        _, text = oracle.run_reference(None, flags, geno_path=os.path.join(HERE, inp), n_ind=n_ind, n_sites=n_sites)
        with open(os.path.join(HERE, name + ".dist"), "w") as fh:
            fh.write(text)
        manifest["text"].append(dict(name=name, input=inp, n_ind=n_ind, n_sites=n_sites, flags=flags))
    manifest["reference_version"] = "1.0.10 (ngsDist.cpp:25)"
    manifest["built_with"] = subprocess.check_output(["/usr/bin/g++", "--version"]).decode().split("\n")[0]
    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print("wrote %d binary + %d text cases" % (len(CASES), len(TEXT_CASES)))


if __name__ == "__main__":
    main()
