"""Summarise .ncu-rep captures (read here, no GPU) into a markdown table for profiles/."""
import csv, subprocess, sys, io, os

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe % (active)"),
    ("sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "IMMA pipe % (active)"),
    ("sm__ops_path_tensor_src_fp64.sum.pct_of_peak_sustained_elapsed", "FP64 tensor ops % of peak (elapsed)"),
    ("sm__ops_path_tensor_src_fp64.sum.per_second", "FP64 tensor FMA/ns"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 (DFMA) pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 inst %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
]

def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        res.append(d)
    return res

def main():
    for path in sys.argv[1:]:
        for d in load(path):
            name = d.get("Kernel Name", ("?", ""))[0]
            print("### `%s` — %s\n" % (os.path.basename(path), name.split("(")[0]))
            print("| metric | value |\n|---|---|")
            for k, label in KEYS:
                if k in d and d[k][0] != "":
                    print("| %s (`%s`) | %s %s |" % (label, k, d[k][0], d[k][1]))
            print()

if __name__ == "__main__":
    main()
