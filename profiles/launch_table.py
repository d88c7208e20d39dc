"""Aggregates an ncu launch list (csv of gpu__time_duration.sum) into a per-kernel table.
    python profiles/launch_table.py <launches.csv> <steps in the window> [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
nsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
hdr, data = None, []
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    k = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("advmil::", "")[:70]
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1000 if d["Metric Unit"] == "ns" else v * 1000 if d["Metric Unit"] == "ms" else v
    agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{len(data)} launches, {tot / nsteps:.1f} us of kernel time per step")
print("| kernel | launches/step | us/launch | us/step | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
    print(f"| `{k}` | {len(v) / nsteps:.1f} | {sum(v) / len(v):.1f} | {sum(v) / nsteps:.1f} | {100 * sum(v) / tot:.1f}% |")
