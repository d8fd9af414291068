#!/bin/bash
# Turn the captures of tools/profile_r02.sh (gpurun_out/r02_*.ncu-rep) into the tracked summaries under profiles/.
R=r02
O=profiles/${R}_ncu_summary.md
{
echo "# Round 02 — ncu evidence"
echo
echo "Captured with \`tools/profile_r02.sh\` on one B200 (each ncu run preceded by the same command exiting 0 without ncu)."
echo "Raw pages: \`${R}_*_raw.csv\`; launch list: \`${R}_launches_bench.csv\`. Bench numbers are NOT taken from these runs."
echo
echo "## Launch list of \`python bench.py --steps 1 --warmup 3 --no-cpu-baseline --core-only --no-e2e --reps 2\` (C3: 2 000 x 1 000 000, weighted replicates)"
echo
python - <<'PY'
import csv, collections, re
lines = [l for l in open("gpurun_out/r02_launches_bench.csv") if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    seq.append((k, float(row["Metric Value"].replace(",", "")) / 1e3))
# one step = k_frontend ... up to the next k_frontend; keep the LAST complete step (after the warm-ups)
starts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_frontend")]
step = seq[starts[-1]:] if len(starts) >= 1 else seq
# the synthetic-data generator launches precede the first front end; drop trailing probe kernels
step = [(k, us) for k, us in step if not k.startswith("k_dmma_peak") and not k.startswith("k_synth")]
agg = collections.OrderedDict()
for k, us in step:
    agg.setdefault(k, []).append(us)
tot = sum(sum(v) for v in agg.values())
print("| kernel | launches in the step | total µs (ncu: cold cache, serialised) | share of the resident step |\n|---|---|---|---|")
for k, v in agg.items():
    print("| `%s` | %d | %.1f | %.2f %% |" % (k, len(v), sum(v), sum(v) / tot * 100))
print("| sum | | %.1f | 100 %% |" % tot)
PY
echo
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_plain_bench.log") if l.startswith("{")][-1])
r, sh = d["roofline"], d["step_share_rank0_ms"]
print("Live CUDA-event figures of the same command without ncu (`r02_bench_plain.log`): step %.1f ms for 2 replicates; `k_dist_dmma` %.1f ms per launch = %.1f %% of the step"
      " (front end %.1f ms, tensor-core count %.1f ms, epilogue %.2f ms; the rest is host bookkeeping, D2H of the matrices and launch gaps, which the serialised ncu list does not see) -- shares agree." %
      (d["ms_per_step"], r["kernel_ms"], 100 * sh["contraction"] / d["ms_per_step"], sh["frontend"], sh["mask_count_overlapped"], sh["epilogue"]))
PY
echo
echo "## Full captures (\`ncu --set full --clock-control none --import-source on\`)"
echo
for f in dist_dmma_weighted_c3 count_umma_c3 frontend_c3 dist_em3 dist_umma_c4; do
  if [ -f gpurun_out/${R}_$f.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/${R}_$f.ncu-rep
    ncu -i gpurun_out/${R}_$f.ncu-rep --page raw --csv > profiles/${R}_${f}_raw.csv 2>/dev/null
  fi
done
} > $O
cp gpurun_out/${R}_launches_bench.csv profiles/
grep "^{" gpurun_out/${R}_plain_bench.log > profiles/${R}_bench_plain.log
cp gpurun_out/${R}_plain_em.log profiles/${R}_plain_em.log; cp gpurun_out/${R}_plain_c4.log profiles/${R}_plain_c4.log
python - <<'PY'
import csv, json
def tr(path, row=2):
    rows = list(csv.reader(open(path)))
    d = dict(zip(rows[0], rows[row])); u = dict(zip(rows[0], rows[1]))
    def b(k):
        v = float(d[k].replace(",", "")); un = u[k]
        return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[un])
    return {"dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"), "source": path + " (ncu --set full, one launch)"}
out = json.load(open("profiles/traffic.json"))
out["k_dist_dmma_c3_weighted"] = tr("profiles/r02_dist_dmma_weighted_c3_raw.csv")
out["k_count_umma_c3"] = tr("profiles/r02_count_umma_c3_raw.csv")
out["k_frontend_c3"] = tr("profiles/r02_frontend_c3_raw.csv")
out["k_dist_em3"] = tr("profiles/r02_dist_em3_raw.csv")
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps({k: out[k] for k in ("k_dist_dmma_c3_weighted", "k_count_umma_c3", "k_frontend_c3")}, indent=1))
PY
