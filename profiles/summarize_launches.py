"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time, share and count.
Usage: python profiles/summarize_launches.py gpurun_out/launches.csv [header text] > profiles/rX_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    name = re.sub(r"\(.*$", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    rows.append((name, ms))
agg = defaultdict(lambda: [0.0, 0])
for n, ms in rows:
    agg[n][0] += ms
    agg[n][1] += 1
total = sum(ms for _, ms in rows)
if len(sys.argv) > 2:
    print(sys.argv[2])
print(f"# total {total:.3f} ms over {len(rows)} launches; compare SHARES with bench.py's kernel_families_ms")
for n, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:9.3f} ms {100 * ms / total:5.1f}%  x{c:4d}  {n}")
