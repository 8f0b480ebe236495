"""Developer tool: aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/summarize_launches.py launches.csv [runs_in_file] > profiles/xxx.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
runs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
per = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    unit = row["Metric Unit"]
    t = t / 1e3 if unit == "ns" else (t * 1e3 if unit == "ms" else t)
    name = row["Kernel Name"]
    m = re.search(r"(?:d3::\(anonymous namespace\)::|d3::|at::native::|at::)?([A-Za-z_0-9]+)\s*(<[^>]*>)?\s*\(", name)
    key = (m.group(1) + (m.group(2) or "")) if m else name[:60]
    per[key][0] += 1
    per[key][1] += t
tot = sum(v[1] for v in per.values())
print(f"{path}: per-kernel totals of the ncu launch list (gpu__time_duration.sum, --clock-control none), divided by {runs:g} runs")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:36]:
    print(f"{v[1] / runs / 1e3:9.2f} ms {v[0] / runs:8.1f} launches {100 * v[1] / tot:5.1f}%  {k[:110]}")
print(f"total {tot / runs / 1e3:.1f} ms")
