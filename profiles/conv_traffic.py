"""Summarise an ncu capture of dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum over the tensor-core
launches of one training step into the JSON bench.py reads for `roofline.traffic`.
Usage: python profiles/conv_traffic.py gpurun_out/conv_dram_<tag>.csv > profiles/<tag>_conv_traffic.json"""
import csv
import json
import re
import sys
from collections import defaultdict

with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
per = defaultdict(dict)
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1.0)
    per[int(r["ID"])][r["Metric Name"]] = v * scale
    per[int(r["ID"])]["name"] = re.sub(r"^void |\(.*$", "", r["Kernel Name"])
# the capture may span more than one step: keep exactly one period of the (kernel name, grid) sequence
ids = sorted(per)
seq = [per[i]["name"] for i in ids]
period = len(seq)
for P in range(100, len(seq) // 2 + 1):
    if all(seq[i] == seq[i + P] for i in range(len(seq) - P)):
        period = P
        break
per = {i: per[i] for i in ids[:period]}
rd = sum(p.get("dram__bytes_read.sum", 0.0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0.0) for p in per.values())
ms = sum(p.get("gpu__time_duration.sum", 0.0) for p in per.values())
by = defaultdict(lambda: [0, 0.0, 0.0])
for p in per.values():
    e = by[p["name"]]
    e[0] += 1; e[1] += p.get("dram__bytes_read.sum", 0.0) + p.get("dram__bytes_write.sum", 0.0); e[2] += p.get("gpu__time_duration.sum", 0.0)
print(json.dumps({
    "source": sys.argv[1], "launches": len(per), "dram_read_bytes": rd, "dram_write_bytes": wr,
    "dram_bytes_per_launch": (rd + wr) / max(len(per), 1), "kernel_ms_serialised_cold": ms,
    "per_kernel": {k: {"launches": v[0], "dram_bytes": v[1], "ms": round(v[2], 4)} for k, v in sorted(by.items())},
    "note": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over "
            "the conv_gemm / wgrad launches of one training step (B=256 pairs, 256x256); serialised, cold cache"}, indent=1))
