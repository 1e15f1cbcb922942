#!/usr/bin/env python
"""Hottest CUDA source lines of one kernel from an .ncu-rep (compile with -lineinfo).
usage: ncu_src.py <report> <kernel-regex> [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rd = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rd) if r and r[0] == "Line No")
hdr = rd[hi]
ci, ii, ai = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
rows = []
for r in rd[hi + 1:]:
    if len(r) <= ii or r[ai] != "-":      # keep the per-source-line aggregate rows only
        continue
    try:
        rows.append((float(r[ci] or 0), float(r[ii] or 0), r[0], r[1].strip()[:110]))
    except ValueError:
        pass
tot = sum(r[0] for r in rows) or 1
toti = sum(r[1] for r in rows) or 1
print("== %s: %d samples, %.4g warp-instructions" % (rx, tot, toti))
for v, n, ln, src in sorted(rows, reverse=True)[:top]:
    print("%6.0f %5.1f%% | inst %5.1f%% | L%-4s %s" % (v, 100 * v / tot, 100 * n / toti, ln, src))
