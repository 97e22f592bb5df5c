#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into one line per launch: python tools/ncu_summary.py file.ncu-rep"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct","smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]][:40]
    parts = [name, "grid=" + r[idx.get("Grid Size", 0)], "blk=" + r[idx.get("Block Size", 0)]]
    for k in KEYS:
        if k in idx:
            parts.append(f"{k.split('.')[0].replace('__', ':')}={r[idx[k]]}{units[idx[k]]}")
    print(" | ".join(parts))
