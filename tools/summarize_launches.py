"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST `n` d3 launches
(one eager UNet forward = 172 launches).  usage: python tools/summarize_launches.py file.csv [n_last]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 172
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
MINE = re.compile(r"igemm|gn_apply|attention_hd8|conv_out|conv_in_direct|im2col|time_embed|temb_project|upsample2x|d3::")
rows = [r for r in csv.DictReader(lines) if MINE.search(r["Kernel Name"])]
rows = rows[-n_last:]
agg = collections.defaultdict(lambda: [0, 0.0])
detail = collections.defaultdict(float)
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("d3::<unnamed>::", "").replace("unnamed>::", "")
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r["Metric Unit"]]
    agg[name][0] += 1
    agg[name][1] += v
    detail[(name, r["Grid Size"])] += v
tot = sum(v[1] for v in agg.values())
print(f"# {len(rows)} launches, total {tot:.3f} ms (cold-cache, serialised under ncu: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%  x{v[0]:4d}  {k}")
