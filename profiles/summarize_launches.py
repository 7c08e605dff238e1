"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv "<command line that was profiled>" > profiles/rNN_ncu_launches_summary.txt"""
import csv
import re
import sys

path, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
col = {h: i for i, h in enumerate(hdr)}
agg = {}
total = 0.0
n = 0
for r in rd:
    if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    total += ms
    n += 1
print("ncu --metrics gpu__time_duration.sum --clock-control none :", what)
print("launches %d, total kernel time %.3f ms (cold-cache, serialised: shares are meaningful, absolutes are not)" % (n, total))
for name, (k, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-92s n=%5d %10.3f ms %6.1f%%" % (name[:92], k, ms, 100.0 * ms / total))
