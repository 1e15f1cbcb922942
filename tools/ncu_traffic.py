#!/usr/bin/env python
"""`ncu --set full` report -> per-kernel table (markdown) + DRAM traffic per launch for bench.py's `roofline.traffic`.

usage: ncu_traffic.py <report.ncu-rep | raw-page.csv> <config-key> [--md profiles/rNN_ncu_full_step_kernels.md] [--json profiles/r02_ncu_traffic.json]

The report holds ONE step's worth of the headline kernels (tools/gpu_ncu_top.sh: the window of consecutive matching
launches is exactly as long as a step's launch count, so every launch of the step is in it once).  Written into the JSON under
<config-key> (bench.py's `<mode>_b<images>_p<proposals>_k<classes>`):
  gemm           sum over the step's tensor-core GEMM launches of dram__bytes_read.sum + dram__bytes_write.sum (the roofline
                 entry is the sum of those launches too), plus the launch count and the per-launch mean;
  roi_align_fwd  the same for the ROIAlign forward launch;  roi_align_bwd for the backward gather.
"""
import argparse
import csv
import json
import os
import re
import subprocess

METRICS = {
    "us": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "tensor": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_op": "sm__inst_executed_pipe_tensor.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "issue": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "inst": "smsp__inst_executed.sum",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def rows_of(rep):
    if rep.endswith(".csv"):                 # `ncu -i report --page raw --csv` already exported (on the GPU box)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rd) if "Kernel Name" in r)
    hdr, units = rd[hi], rd[hi + 1]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rd[hi + 2:]:
        if len(r) < len(hdr):
            continue
        d = {"name": r[col["Kernel Name"]]}
        for k, m in METRICS.items():
            if m not in col:
                d[k] = None
                continue
            try:
                v = float(r[col[m]].replace(",", ""))
            except ValueError:
                d[k] = None
                continue
            d[k] = v * UNIT_SCALE.get(units[col[m]], 1.0)
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("key")
    ap.add_argument("--md")
    ap.add_argument("--json", default=os.path.join(os.path.dirname(__file__), "..", "profiles", "r02_ncu_traffic.json"))
    ap.add_argument("--command", default="tools/gpu_ncu_top.sh")
    a = ap.parse_args()
    rows = rows_of(a.report)
    short = lambda n: re.sub(r"^b200::", "", re.sub(r"\(.*", "", n))
    groups = {"gemm": r"gemm2_pair_kernel|gemm_bf16_tcgen05_kernel", "roi_align_fwd": r"roi_align_fwd", "roi_align_bwd": r"roi_bwd_csr_gather",
              "sgd": r"sgd_momentum"}
    entry = {}
    for g, rx in groups.items():
        sel = [r for r in rows if re.search(rx, r["name"]) and r["rd"] is not None]
        if not sel:
            continue
        tot = sum(r["rd"] + r["wr"] for r in sel)
        entry[g] = tot if g == "gemm" else tot / len(sel)
        entry[g + "_launches"] = len(sel)
        entry[g + "_read_bytes"] = sum(r["rd"] for r in sel)
        entry[g + "_write_bytes"] = sum(r["wr"] for r in sel)
        entry[g + "_us_under_ncu"] = sum(r["us"] for r in sel)
    entry["source"] = "ncu --set full --clock-control none, %s, report summarised in %s" % (a.command, os.path.basename(a.md or ""))
    try:
        allj = json.load(open(a.json))
    except Exception:  # noqa: BLE001
        allj = {}
    allj[a.key] = entry
    json.dump(allj, open(a.json, "w"), indent=1, sort_keys=True)
    if a.md:
        f = lambda v, s="%.1f": "-" if v is None else s % v
        with open(a.md, "w") as fh:
            fh.write("# ncu --set full: one step's launches of the headline kernels (%s)\n\n" % a.key)
            fh.write("Command: `%s` (after the same bench command exited 0 without ncu).  Per-launch times are cold-cache and serialised; "
                     "shares, not absolutes, compare with the bench.\n\n" % a.command)
            fh.write("| # | kernel | us | grid | dram rd MB | dram wr MB | tensor pipe % | dram % | L2 % | L1 % | issue % | regs | warp instr |\n")
            fh.write("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
            for i, r in enumerate(rows):
                fh.write("| %d | `%s` | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |\n" % (
                    i, short(r["name"])[:60], f(r["us"]), f(r["grid"], "%d"), f(r["rd"] and r["rd"] / 1e6), f(r["wr"] and r["wr"] / 1e6),
                    f(r["tensor"]), f(r["dram_pct"]), f(r["l2_pct"]), f(r["l1_pct"]), f(r["issue"]), f(r["regs"], "%d"), f(r["inst"], "%d")))
            fh.write("\nTotals per group (bytes are dram__bytes_read.sum + dram__bytes_write.sum):\n\n")
            for g in groups:
                if g in entry:
                    fh.write("* %s: %d launches, %.1f MB read + %.1f MB written, %.1f us under ncu\n" % (
                        g, entry[g + "_launches"], entry[g + "_read_bytes"] / 1e6, entry[g + "_write_bytes"] / 1e6, entry[g + "_us_under_ncu"]))
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
