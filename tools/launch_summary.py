#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-step table (markdown).
usage: launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        rows.append((r["Kernel Name"], v))
# a step starts at the forward affine pass (fp32 map in); the backward one (bf16 gradient in) is mid-step
starts = [i for i, (n, _) in enumerate(rows) if "affine_tile_kernel<float" in n or "affine_transpose_kernel<float" in n]
a, b = starts[-2], starts[-1]
tot = sum(v for _, v in rows[a:b])
ours = sum(v for n, v in rows[a:b] if "b200::" in n)
print("| # | kernel | us | share of step |\n|---|---|---:|---:|")
for i, (n, v) in enumerate(rows[a:b]):
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"^void ", "", n)[:90]
    print("| %d | `%s` | %.1f | %.1f%% |" % (i, n, v, 100 * v / tot))
print("\nstep total %.1f us over %d launches; hand-written b200:: kernels %.1f us (%.1f%%), library/torch kernels %.1f us" %
      (tot, b - a, ours, 100 * ours / tot, tot - ours))
