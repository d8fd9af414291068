"""The drop-in command line (ngsdist_b200/bin/ngsDist): same flags, inputs and .dist layout as the reference.

CPU part: argument validation reproduces parse_args.cpp:203-220 (message + exit status 255) before any CUDA call.
GPU part: every golden case is run through the binary and compared with the reference's own .dist text.
"""
import os
import subprocess

import numpy as np
import pytest

import oracle
from util import GOLDEN, golden_text, manifest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ngsdist_b200", "bin", "ngsDist")
MAN = manifest()


def run_cli(args, **kw):
    return subprocess.run([CLI] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)


def test_cli_is_built():
    assert os.path.exists(CLI), "run __graft_entry__.build()"


@pytest.mark.parametrize("args,msg", [
    ([], "genotype input file (--geno) missing!"),
    (["--geno", "x.bin"], "number of individuals (--n_ind) missing!"),
    (["--geno", "x.bin", "--n_ind", "3"], "number of sites (--n_sites) missing!"),
    (["--geno", "x.bin", "--n_ind", "3", "--n_sites", "4", "--tot_sites", "9", "--pairwise_del", "--out", "o"], "cannot specify total number of sites"),
    (["--geno", "x.gz", "--n_ind", "3", "--n_sites", "4", "--call_geno", "--out", "o"], "can only call genotypes from likelihoods/probabilities!"),
    (["--geno", "x.bin", "--n_ind", "3", "--n_sites", "4", "--evol_model", "9", "--out", "o"], "invalid correction method specified!"),
    (["--geno", "x.bin", "--n_ind", "3", "--n_sites", "4", "--evol_model", "3", "--out", "o"], "requires position information!"),
    (["--geno", "x.bin", "--n_ind", "3", "--n_sites", "4"], "output prefix (--out) missing!"),
    (["--geno", "x.bin", "--n_ind", "3", "--n_sites", "4", "--out", "o", "--n_threads", "0"], "number of threads cannot be less than 1!"),
    # single-dash long options and unique prefixes are accepted (getopt_long_only), then the size check fires
    (["-geno", os.path.join(GOLDEN, "g7x53.bin"), "-n_ind", "7", "-n_sites", "50", "-out", "o", "-verb", "0"], "invalid/corrupt genotype input file!"),
])
def test_cli_argument_errors(args, msg):
    r = run_cli(args + ([] if "-verb" in args else ["--verbose", "0"]))
    assert r.returncode == 255
    assert "ERROR: [" in r.stderr and msg in r.stderr


def test_cli_io_selftest_exact_formatter_and_parser():
    """The writer's exact fast "%.10f" and the reader's fast number parser against libc on 2e6 random values (+ specials,
    decimal half-way points, non-numbers): must be byte / bit identical."""
    r = run_cli(["--selftest_io", "2000000"])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "0 mismatches" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in MAN["binary"] + MAN["text"] if c["n_sites"] > 64][:6], ids=lambda c: c["name"])
def test_cli_chunked_threaded_reader(case, tmp_path):
    """Many 64-site chunks through the double-buffered reader thread (NGSD_CLI_CHUNK) give the same file as one chunk."""
    outs = []
    for chunk in ("64", "1000000"):
        out = str(tmp_path / ("out%s.dist" % chunk))
        args = ["--geno", os.path.join(GOLDEN, case["input"]), "--n_ind", str(case["n_ind"]), "--n_sites", str(case["n_sites"]),
                "--out", out, "--n_threads", "3", "--verbose", "0"] + case["flags"]
        r = run_cli(args, env=dict(os.environ, NGSD_CLI_CHUNK=chunk))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(open(out).read())
    assert outs[0] == outs[1]


@pytest.mark.gpu
def test_cli_text_reader_many_chunks_matches_binary_input(tmp_path):
    """C1-style text input (header + 3 leading label columns, gz) read in 64-site chunks by the threaded parser gives
    the same .dist as the same values read from the binary file."""
    import gzip
    from util import load_bin
    n_ind, n_sites = 24, 400
    raw = load_bin("g24x400.bin", n_ind, n_sites)
    txt = str(tmp_path / "g.txt.gz")
    with gzip.open(txt, "wt") as fh:
        fh.write("marker\tallele1\tallele2\t" + "\t".join("Ind%d" % (i // 3) for i in range(n_ind * 3)) + "\n")
        for s in range(n_sites):
            fh.write("chr1_%d\tA\tC\t" % s + "\t".join(repr(float(v)) for v in raw[s].reshape(-1)) + "\n")
    outs = []
    for geno, chunk in ((os.path.join(GOLDEN, "g24x400.bin"), "1000000"), (txt, "64")):
        out = str(tmp_path / ("o%s.dist" % chunk))
        r = run_cli(["--geno", geno, "--probs", "--n_ind", str(n_ind), "--n_sites", str(n_sites), "--out", out, "--n_threads", "4",
                     "--verbose", "0", "--indep_geno", "--n_boot_rep", "2", "--boot_block_size", "10", "--seed", "5"],
                    env=dict(os.environ, NGSD_CLI_CHUNK=chunk))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(open(out).read())
    assert outs[0] == outs[1]


@pytest.mark.gpu
@pytest.mark.parametrize("case", MAN["binary"] + MAN["text"], ids=lambda c: c["name"])
def test_cli_reproduces_reference_dist_files(case, tmp_path):
    out = str(tmp_path / "out.dist")
    args = ["--geno", os.path.join(GOLDEN, case["input"]), "--n_ind", str(case["n_ind"]), "--n_sites", str(case["n_sites"]),
            "--out", out, "--n_threads", "4", "--verbose", "0"] + case["flags"]
    r = run_cli(args)
    assert r.returncode == 0, r.stderr[-2000:]
    got_text = open(out).read()
    want_text = golden_text(case["name"])
    got = [m for _, m in oracle.parse_dist(out, case["n_ind"])]
    ref_path = str(tmp_path / "ref.dist")
    open(ref_path, "w").write(want_text)
    want = [m for _, m in oracle.parse_dist(ref_path, case["n_ind"])]
    assert len(got) == len(want)
    # identical layout: same lines, labels, column counts
    assert [l.split("\t")[0] for l in got_text.split("\n")] == [l.split("\t")[0] for l in want_text.split("\n")]
    for g, w in zip(got, want):
        fin = np.isfinite(w)
        assert np.array_equal(np.isnan(g), np.isnan(w)) and np.array_equal(np.isinf(g), np.isinf(w))
        assert np.abs(g[fin] - w[fin]).max() <= 1.1e-10      # at most one unit in the 10th printed decimal
    if case["name"] in ("txt_geno_pdel_boot", "call", "txt_geno"):
        # called / genotype input: exact sums (multiples of 0.5 or the same u = 1/3 terms) -> byte-identical text expected
        if case["name"] == "txt_geno_pdel_boot":
            assert got_text == want_text


@pytest.mark.gpu
def test_cli_labels_stdin_and_nan_spelling(tmp_path):
    n_ind, n_sites = 5, 64
    raw = oracle.synth_raw(3, 0.0, n_ind, n_sites)
    raw[:32, 0, :] = 1 / 3
    raw[32:, 1, :] = 1 / 3                      # individuals 0 and 1 never share a site -> cnt == 0 -> -nan / nan
    geno = tmp_path / "in.bin"
    raw.tofile(geno)
    labels = tmp_path / "labels.txt"
    labels.write_text("# comment\nheader\nA\tx\nB\nC\nD\nE\n")
    out = tmp_path / "o.dist"
    flags = ["--probs", "--indep_geno", "--pairwise_del", "--evol_model", "0"]
    r = run_cli(["--geno", str(geno), "--n_ind", "5", "--n_sites", "64", "--labelsH", str(labels), "--out", str(out), "--verbose", "0"] + flags)
    assert r.returncode == 0, r.stderr
    ref, ref_text = oracle.run_reference(None, flags + ["--labelsH", str(labels)], geno_path=str(geno), n_ind=n_ind, n_sites=n_sites) \
        if oracle.have_ref() else (None, None)
    text = out.read_text()
    assert text.split("\n")[2].startswith("A\t0.0000000000\t-nan\t")
    if ref_text is not None:
        assert [l.split("\t")[0] for l in text.split("\n")] == [l.split("\t")[0] for l in ref_text.split("\n")]
        assert text.count("nan") == ref_text.count("nan") and text.count("-nan") == ref_text.count("-nan")
    # binary from stdin ("-")
    out2 = tmp_path / "o2.dist"
    with open(geno, "rb") as fh:
        r = subprocess.run([CLI, "--geno", "-", "--probs", "--n_ind", "5", "--n_sites", "64", "--labelsH", str(labels), "--out", str(out2),
                            "--verbose", "0"] + flags[1:], stdin=fh, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert out2.read_text() == text
