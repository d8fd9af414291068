"""Whole-process comparison on one file: the drop-in command line (ngsdist_b200/bin/ngsDist) vs the unmodified reference
(oracle/_ref/ngsDist) -- same flags, same binary input, wall-clock of the complete run (read + compute + write), and the
two .dist files compared value by value.  Default: the C2 data set (500 x 100 000, --probs --indep_geno --evol_model 2)."""
import argparse, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--n-ind", type=int, default=500)
ap.add_argument("--n-sites", type=int, default=100000)
ap.add_argument("--flags", default="--probs --indep_geno --evol_model 2")
ap.add_argument("--ref-sites", type=int, default=0, help="run the reference on the first N sites only (0 = all)")
args = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cli = os.path.join(root, "ngsdist_b200", "bin", "ngsDist")
work = tempfile.mkdtemp(prefix="ngsd_cli_e2e_")
path = os.path.join(work, "in.bin")
t0 = time.time()
step = 10000
with open(path, "wb") as fh:
    for s0 in range(0, args.n_sites, step):
        oracle.synth_raw(20251018, 0.0, args.n_ind, min(step, args.n_sites - s0), site0=s0).tofile(fh)
print("wrote %s (%.2f GB) in %.1f s" % (path, os.path.getsize(path) / 1e9, time.time() - t0))
threads = os.cpu_count() or 1
flags = args.flags.split()

last_stamps = [""]

def run(binary, n_sites, out):
    cmd = [binary, "--geno", path, "--n_ind", str(args.n_ind), "--n_sites", str(n_sites), "--out", out, "--n_threads", str(threads),
           "--verbose", "0"] + flags
    t = time.perf_counter()
    r = subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=dict(os.environ, NGSD_CLI_TIMING="1"))
    dt = time.perf_counter() - t
    last_stamps[0] = "".join(l + "\n" for l in r.stderr.splitlines() if l.startswith("[timing]"))
    return dt

ours = [run(cli, args.n_sites, os.path.join(work, "ours.dist")) for _ in range(3)]
print("ngsdist_b200/bin/ngsDist: %s s wall (3 runs; every run pays CUDA initialisation + context creation, which depends on the box: see the stamps)" % ", ".join("%.2f" % t for t in ours))
print("phases of the last run:\n" + last_stamps[0], end="")
pairs = args.n_ind * (args.n_ind - 1) // 2
if oracle.have_ref():
    ref_sites = args.ref_sites or args.n_sites
    if ref_sites != args.n_sites:
        # the reference checks the file size against n_sites: give it its own truncated file
        sub = os.path.join(work, "ref.bin")
        with open(path, "rb") as fi, open(sub, "wb") as fo:
            fo.write(fi.read(ref_sites * args.n_ind * 24))
        path_full, path = path, sub
    tr = run(oracle.REF_BIN, ref_sites, os.path.join(work, "ref.dist"))
    print("reference ngsDist --n_threads %d on %d sites: %.2f s wall" % (threads, ref_sites, tr))
    if ref_sites == args.n_sites:
        a = [m for _, m in oracle.parse_dist(os.path.join(work, "ours.dist"), args.n_ind)]
        b = [m for _, m in oracle.parse_dist(os.path.join(work, "ref.dist"), args.n_ind)]
        worst = max(float(np.nanmax(np.abs(x - y))) for x, y in zip(a, b))
        same = open(os.path.join(work, "ours.dist")).read() == open(os.path.join(work, "ref.dist")).read()
        print("matrices: %d, max |difference| of printed values %.3g, files byte-identical: %s" % (len(a), worst, same))
    print("whole-process speed-up (reference wall scaled to %d sites / best of ours): %.0fx" %
          (args.n_sites, tr * args.n_sites / ref_sites / min(ours)))
