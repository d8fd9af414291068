"""Per-kernel share of one resident bench step from an ncu launch list (gpu__time_duration.sum per launch)."""
import csv, collections, re, sys

def main(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        seq.append((k, float(row["Metric Value"].replace(",", "")) / 1e3, row["Grid Size"]))
    steps, cur = [], []
    for k, us, grid in seq:
        base = re.sub(r"<.*>", "", k)
        if base == "k_frontend" and not cur:
            cur = [(k, us)]
        elif cur and base in ("k_dist_dmma", "k_cvec", "k_zero_diag", "k_epilogue"):
            cur.append((k, us))
            if base == "k_epilogue":
                steps.append(cur); cur = []
        else:
            cur = []
    agg = collections.OrderedDict()
    for s in steps:
        for k, us in s:
            agg.setdefault(k, []).append(us)
    tot = sum(sum(v) / len(v) for v in agg.values())
    print("| kernel | launches seen | mean µs (ncu: cold cache, serialised) | share of the resident step |\n|---|---|---|---|")
    for k, v in agg.items():
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v), sum(v) / len(v) / tot * 100))
    print("| sum | | %.1f | 100 %% |" % tot)

if __name__ == "__main__":
    main(sys.argv[1])
