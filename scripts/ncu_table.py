"""Formats the headline metrics of every kernel in an .ncu-rep as a text table (for profiles/).
usage: python scripts/ncu_table.py report.ncu-rep "label 1|label 2|..." > profiles/xxx.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
labels = sys.argv[2].split("|") if len(sys.argv) > 2 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(name):
    return hdr.index(name)
M = [("time_us", "gpu__time_duration.sum", 1.0), ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0), ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("dramR_MB", "dram__bytes_read.sum", 1.0),
     ("dramW_MB", "dram__bytes_write.sum", 1.0), ("regs", "launch__registers_per_thread", 1.0), ("grid", "launch__grid_size", 1.0),
     ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0)]
print(f"{'kernel':44s} " + " ".join(f"{m[0]:>9s}" for m in M) + "  what")
units = rows[1]
for i, r in enumerate(rows[2:]):
    name = r[col("Kernel Name")]
    name = name.replace("void mmb::<unnamed>::", "").replace("mmb::", "").split("(CUtensorMap")[0][:44]
    vals = []
    for short, full, _ in M:
        v = float(r[col(full)].replace(",", ""))
        u = units[col(full)]
        if short.endswith("_MB"):
            v = v / 1e6 if u == "byte" else v / 1e3 if u == "Kbyte" else v * 1e3 if u == "Gbyte" else v
        if short == "time_us":
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        vals.append(v)
    print(f"{name:44s} " + " ".join(f"{v:9.1f}" for v in vals) + "  " + (labels[i] if i < len(labels) else ""))
