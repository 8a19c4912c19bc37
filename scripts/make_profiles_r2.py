"""Turns the files one `scripts/gpu_profile_r2.sh <tag>` call left in gpurun_out/ into the tracked summaries of profiles/:
    python scripts/make_profiles_r2.py gpurun_out r2p
"""
import collections, csv, io, json, os, re, shutil, subprocess, sys

src, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(root, "profiles")
WL = ("mosei_unaligned_b64", "mosi_aligned_b64", "ur_funny_b64")


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void mmb::<unnamed>::", "").replace("void mmb::", "").replace("mmb::", "")[:58]


def launches(path):
    """ncu --csv launch list with several metrics -> [(kernel, {metric: value})] in launch order (ns / bytes)."""
    lines = [l for l in open(path, newline="") if l.startswith('"')]
    rows = collections.OrderedDict()
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        scale = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "byte": 1.0, "Kbyte": 1e3,
                 "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        rows.setdefault(r["ID"], [r["Kernel Name"], {}])[1][r["Metric Name"]] = v * scale
    return list(rows.values())


traffic = {}
for w in WL:
    for kind in ("bench", "step_table"):
        ext = "json" if kind == "bench" else "txt"
        p = os.path.join(src, f"{tag}_{kind}_{w}.{ext}")
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(dst, f"r2_{kind}_{w}.{ext}"))
    p = os.path.join(src, f"{tag}_launches_{w}.csv")
    if not os.path.exists(p):
        continue
    L = launches(p)
    tot = collections.OrderedDict()
    for name, m in L:
        a = tot.setdefault(short(name), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in tot.values())
    with open(os.path.join(dst, f"r2_launches_{w}.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                f"--profile-from-start off python scripts/one_step.py {w}\n"
                f"# every kernel of ONE training step (forward, backward, AdamW, gradient zeroing): {len(L)} launches, "
                f"{total / 1e6:.3f} ms serialised under ncu (cold caches: shares, not absolute times, carry over to the bench)\n")
        f.write(f"{'kernel':58s} {'n':>4s} {'total_ms':>9s} {'share':>7s} {'avg_us':>9s} {'dram_GB':>8s} {'GB/s':>8s}\n")
        for name, (c, ns, by) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name:58s} {c:4d} {ns / 1e6:9.3f} {100 * ns / total:6.1f}% {ns / c / 1e3:9.1f} {by / 1e9:8.2f} {by / ns:8.0f}\n")
    g = [(n, m) for n, m in L if "gemm_tcgen05" in n]
    if g:
        rd = sum(m.get("dram__bytes_read.sum", 0.0) for _, m in g)
        wr = sum(m.get("dram__bytes_write.sum", 0.0) for _, m in g)
        traffic[w] = {"gemm_launches_captured": len(g), "dram_read_bytes": rd, "dram_write_bytes": wr,
                      "traffic_bytes_per_launch": (rd + wr) / len(g),
                      "sum_duration_us_under_ncu": sum(m["gpu__time_duration.sum"] for _, m in g) / 1e3}
if traffic:
    traffic["_how"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                       "--profile-from-start off python scripts/one_step.py <workload>: every GEMM launch of ONE training step")
    json.dump(traffic, open(os.path.join(dst, "r2_gemm_traffic.json"), "w"), indent=1)
for f in ("rowops.txt", "bench_reference_arm.json", "pytest.txt"):
    p = os.path.join(src, f"{tag}_{f}")
    if os.path.exists(p) and os.path.getsize(p) > 0:
        if f == "pytest.txt":
            open(os.path.join(dst, "r2_gpu_pytest_summary.txt"), "w").write("".join(open(p).readlines()[-4:]))
        else:
            shutil.copy(p, os.path.join(dst, "r2_" + f))

M = [("time_us", "gpu__time_duration.sum"), ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
     ("l1tex%", "l1tex__throughput.avg.pct_of_peak_sustained_active"), ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("dramR_MB", "dram__bytes_read.sum"),
     ("dramW_MB", "dram__bytes_write.sum"), ("GB/s", None), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size")]


def full_table(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    out.write(f"\n## {title}\n{'kernel':50s} " + " ".join(f"{m[0]:>9s}" for m in M) + "\n")
    for r in rows[2:]:
        vals = {}
        for sh, full in M:
            if full is None:
                continue
            v = float(r[hdr.index(full)].replace(",", ""))
            u = units[hdr.index(full)]
            if sh.endswith("_MB"):
                v = {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(u, v)
            if sh == "time_us":
                v = {"ns": v / 1e3, "nsecond": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3}.get(u, v)
            vals[sh] = v
        vals["GB/s"] = (vals["dramR_MB"] + vals["dramW_MB"]) / vals["time_us"] * 1e3
        out.write(f"{short(r[hdr.index('Kernel Name')])[:50]:50s} " + " ".join(f"{vals[m[0]]:9.1f}" for m in M) + "\n")


reps = [(os.path.join(src, f"{tag}_full_fwd.ncu-rep"), "forward: first encoder layer (QKV GEMM, attention, Wo GEMM, LN, FFN1 GEMM (GELU + gelu'), FFN2 GEMM, LN, next QKV)"),
        (os.path.join(src, f"{tag}_full_bwd.ncu-rep"), "backward: one encoder layer (LN bwd, FFN2 dgrad (x gelu', + bias-gradient column sums), FFN2 wgrad, FFN1 dgrad, FFN1 wgrad, LN bwd, Wo dgrad, Wo wgrad, attention bwd prep / dK,dV / dQ, colsum(dqkv), QKV dgrad, QKV wgrad)")]
if any(os.path.exists(r) for r, _ in reps):
    with open(os.path.join(dst, "r2_ncu_full_mosei_layer.txt"), "w") as out:
        out.write("# round 2 — ncu --set full --clock-control none --import-source on --profile-from-start off, MOSEI-unaligned B=64 "
                  "(73 600 packed rows), bert-base; python scripts/one_step.py mosei_unaligned_b64\n"
                  "# tensor% = sm__pipe_tensor_cycles_active; issue% = smsp__issue_active; xu% = MUFU pipe; l1tex% / lts% / dram% = unit "
                  "throughput of peak; DRAM MB per launch; GB/s = DRAM bytes / duration (numbers under a profiler are never bench values)\n")
        for r, t in reps:
            if os.path.exists(r):
                full_table(r, out, t)
print("profiles/ updated:", sorted(f for f in os.listdir(dst) if f.startswith("r2_")))
