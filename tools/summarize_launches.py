"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) launches, total, mean, share.
Usage: python tools/summarize_launches.py launches.csv > summary.md"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"]
    name = re.sub(r"\((?:[A-Za-z_:]|\(anonymous namespace\))[^)]*\)\s*$", "", name)  # drop the parameter list
    name = name.split("::")[-1].strip()
    name = re.sub(r"^void\s+", "", name)
    rows.append((name, r.get("Grid Size", ""), us))
agg = OrderedDict()
for name, grid, us in rows:
    a = agg.setdefault((name, grid), [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
print(f"Total launches profiled: {len(rows)}, total {total / 1e3:.1f} ms\n")
print("| kernel | grid | launches | total us | avg us | share |")
print("|---|---|---:|---:|---:|---:|")
for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | {grid} | {n} | {t:.1f} | {t / n:.2f} | {100 * t / total:.1f}% |")
DECODE_KERNELS = ("step_gemm_kernel", "step_finalize_kernel", "attn_decode_kernel", "sample_kernel", "frame_end_kernel",
                  "step_inc_kernel", "rows_kernel")
dec = {k: v for k, v in agg.items()
       if k[0].startswith(DECODE_KERNELS) or (k[0].startswith("embed_kernel") and k[1].startswith("(32,"))}
dt = sum(v[1] for v in dec.values())
if dt > 0:
    print(f"\n## Decode frames only (the kernels of the per-frame CUDA graph, batch 32): {dt / 1e3:.2f} ms\n")
    print("| kernel | grid | launches | avg us | share of decode |")
    print("|---|---|---:|---:|---:|")
    for (name, grid), (n, t) in sorted(dec.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {grid} | {n} | {t / n:.2f} | {100 * t / dt:.1f}% |")
