"""Development helper: split the SASS-level samples of an .ncu-rep by barrier-delimited phases + opcode histogram + stalls."""
import csv, io, subprocess, sys
from collections import Counter
path = sys.argv[1]
def page(p):
    out = subprocess.run(["ncu", "-i", path, "--page", p, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
rows = page("source")
hdr, data = rows[1], rows[2:]
iS = hdr.index("Warp Stall Sampling (All Samples)"); iI = hdr.index("Instructions Executed")
tot = sum(int(r[iS]) for r in data); toti = sum(int(r[iI]) for r in data)
print("total samples", tot, "total warp inst", toti)
seg_s = seg_i = start = 0
for k, r in enumerate(data):
    seg_s += int(r[iS]); seg_i += int(r[iI])
    if 'BAR.SYNC' in r[1] or 'EXIT' in r[1]:
        print("rows %d-%d samples %.1f%% inst %.1f%% : %s" % (start, k, 100 * seg_s / tot, 100 * seg_i / toti, r[1].strip()))
        seg_s = seg_i = 0; start = k + 1
c = Counter(); s = Counter()
for r in data:
    f = r[1].split()
    op = (f[1] if f[0].startswith('@') else f[0]).split('.')[0]
    c[op] += int(r[iI]); s[op] += int(r[iS])
print("opcode: warp-inst share, sample share")
for op, n in c.most_common(18): print("  %-8s %5.1f%% %5.1f%%" % (op, 100 * n / toti, 100 * s[op] / tot))
raw = page("raw"); h = raw[0]; v = raw[2]
for k, x in zip(h, v):
    if 'smsp__average_warps_issue_stalled' in k and 'per_issue_active' in k and 'not_issued' not in k:
        if float(x) > 0.05: print("stall", k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), x)
    if k in ("gpu__time_duration.sum", "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"):
        print(k, x)
