"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share of the step.  usage: summarize_launches.py launches.csv [out.md]"""
import csv, io, re, sys
from collections import defaultdict
rows = []
lines = [l for l in open(sys.argv[1], errors="ignore") if l.startswith('"')]
rd = csv.reader(io.StringIO("".join(lines)))
hdr = next(rd)
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^.*::", "", name)
    agg[name][0] += 1
    agg[name][1] += us
    total += us
out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {c} | {t:.1f} | {100*t/total:.1f}% |")
out.append(f"| **sum** | {sum(c for c, _ in agg.values())} | {total:.1f} | 100% |")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
