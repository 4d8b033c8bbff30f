#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full --import-source on) into the few numbers the roofline
arguments in DESIGN.md rest on.  Usage: python tools/ncu_summary.py report.ncu-rep [> profiles/rNN/x.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def ncu(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(path):
    rows = ncu(path, "raw")
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for n, vals in enumerate(rows[2:]):
        print(f"== launch {n}: {vals[ik][:110]}")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print(f"  {h:85s} {v} {u}")
    rows = ncu(path, "source")
    # one table per kernel in the source page; only the first is summarised
    try:
        h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    except StopIteration:
        return
    hdr = rows[h]
    ix = {k: i for i, k in enumerate(hdr)}
    data = []
    for r in rows[h + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        data.append(r)
    cat, ex = {}, {}
    reasons = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = dict.fromkeys(reasons, 0)
    prev = None
    for r in data:
        src = r[ix["Source"]].split()
        op = (src[1] if src and src[0].startswith("@") else (src[0] if src else "")).split(".")[0]
        key = "NOP(after DMMA)" if op == "NOP" and prev == "DMMA" else op
        n = int(r[ix["# Samples"]] or 0)
        cat[key] = cat.get(key, 0) + n
        ex[key] = ex.get(key, 0) + int(r[ix["Instructions Executed"]] or 0)
        for k in reasons:
            agg[k] += int(r[ix[k]] or 0)
        prev = op
    tot = max(1, sum(cat.values()))
    print("== warp-state samples by opcode (first kernel): share, warp-instructions executed")
    for k, v in sorted(cat.items(), key=lambda kv: -kv[1])[:12]:
        print(f"  {k:18s} {100 * v / tot:5.1f}%  {ex[k]}")
    print("== warp-state samples by stall reason")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {k:22s} {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
