"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals / shares."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
n = 0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:60]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    n += 1
total = sum(a[1] for a in tot.values())
print(f"# launches captured: {n}  total: {total:.3f} ms")
print(f"{'kernel':56s} {'n':>4s} {'total_ms':>9s} {'share':>7s} {'avg_us':>9s}")
for name, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:56s} {c:4d} {ms:9.3f} {100 * ms / total:6.1f}% {1000 * ms / c:9.1f}")
