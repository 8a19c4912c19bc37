"""Reads an .ncu-rep here (no GPU): headline metrics per kernel and the top stalled SASS instructions.
usage: python scripts/ncu_top.py report.ncu-rep [kernel-regex] [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k[:70]:70s} {units[i]:8s}", [r[i][:44] for r in rows[2:]])
for i, h in enumerate(hdr):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        print(f"{h[27:70]:43s}", [r[i] for r in rows[2:]])
if kre:
    import re
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = src.split('"Kernel Name"')
    seen = set()
    for blk in blocks[1:]:
        name = blk.split("\n")[0]
        if not re.search(kre, name) or name in seen:
            continue
        seen.add(name)
        lines = blk.split("\n")
        print("== kernel", lines[0][:120])
        rr = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        h = rr[0]
        isrc, isamp, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        data = [(int(r[isamp]), n, r[isrc].strip(), int(r[iex])) for n, r in enumerate(rr[1:]) if len(r) > isamp and r[isamp].isdigit()]
        tot = sum(d[0] for d in data) or 1
        for s, n, text, ex in sorted(data, reverse=True)[:ntop]:
            print(f"{100 * s / tot:5.1f}%  #{n:4d} ex={ex:9d}  {text[:100]}")
