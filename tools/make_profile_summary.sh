#!/bin/bash
# Turn the captures of tools/profile_all.sh (gpurun_out/<R>_*.ncu-rep) into the tracked summaries under profiles/.
R=${1:-r01}
O=profiles/${R}_ncu_summary.md
{
echo "# Round ${R#r} — ncu evidence"
echo
echo "Captured with \`tools/profile_all.sh $R\` on one B200 (each ncu run preceded by the same command exiting 0 without ncu)."
echo "Raw pages: \`${R}_*_raw.csv\`; launch list: \`${R}_launches_bench.csv\`. Bench numbers are NOT taken from these runs."
echo
echo "## Launch list of \`python bench.py --steps 2 --warmup 3 --no-cpu-baseline --core-only\` (C2), one resident step"
echo
python tools/launch_share.py gpurun_out/${R}_launches_bench.csv
echo
python - <<PY
import json
d = json.loads(open("gpurun_out/${R}_plain_bench.log").read().strip().splitlines()[-1])
r = d["roofline"]
print("Live CUDA-event figures of the same command without ncu (\`${R}_bench_plain.log\`): step %.3f ms, \`k_dist_dmma\` %.3f ms = %.1f %% of the step"
      " (the step also holds the 2 MB D2H of the matrix and host launch gaps, which the serialised ncu list does not see) -- shares agree." %
      (d["ms_per_step"], r["kernel_ms"], 100 * r["kernel_ms"] / d["ms_per_step"]))
PY
echo
echo "## Full captures (\`ncu --set full --clock-control none --import-source on\`)"
echo
for f in dist_dmma_c2 frontend_c2 dist_dmma_weighted_c3 mask_count_c3 dist_em dist_umma dist_imma frontend_codes; do
  if [ -f gpurun_out/${R}_$f.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/${R}_$f.ncu-rep
    ncu -i gpurun_out/${R}_$f.ncu-rep --page raw --csv > profiles/${R}_${f}_raw.csv 2>/dev/null
  fi
done
} > $O
cp gpurun_out/${R}_launches_bench.csv profiles/
cp gpurun_out/${R}_plain_bench.log profiles/${R}_bench_plain.log
for f in c3 c4 em; do [ -f gpurun_out/${R}_plain_$f.log ] && cp gpurun_out/${R}_plain_$f.log profiles/${R}_plain_$f.log; done
[ -f gpurun_out/cli_e2e.log ] && cp gpurun_out/cli_e2e.log profiles/${R}_cli_e2e.log
python - <<PY
import csv, json
def tr(path):
    rows = list(csv.reader(open(path)))
    d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
    def b(k):
        v = float(d[k].replace(",", "")); un = u[k]
        return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[un])
    return {"dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"), "source": path + " (ncu --set full, one launch)"}
out = {"k_dist_dmma_c2": tr("profiles/${R}_dist_dmma_c2_raw.csv"), "k_frontend_c2": tr("profiles/${R}_frontend_c2_raw.csv"),
       "k_dist_imma": tr("profiles/${R}_dist_imma_raw.csv"), "k_dist_em": tr("profiles/${R}_dist_em_raw.csv")}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
