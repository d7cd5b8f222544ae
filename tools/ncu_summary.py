#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics the DESIGN/profiles notes quote.
usage: tools/ncu_summary.py <report.ncu-rep> [extra-regex]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
pat = re.compile(
    r"^(Kernel Name|gpu__time_duration\.sum|launch__(grid_size|block_size|registers_per_thread|occupancy_limit_\w+|waves_per_multiprocessor)"
    r"|sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum$|sm__inst_executed\.avg\.per_cycle_active"
    r"|sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|fp64|xu|lsu|uniform|cbu|adu)\.sum$"
    r"|sm__inst_executed_pipe_\w+\.avg\.pct_of_peak_sustained_active"
    r"|sm__pipe_\w+_cycles_active\.avg\.pct_of_peak_sustained_active"
    r"|dram__bytes_(read|write)\.sum$|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|lts__t_bytes\.sum$|lts__t_sector_hit_rate\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum"
    r"|smsp__average_warps?_issue_stalled_\w+_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_\w+\.ratio"
    r"|smsp__warp_issue_stalled_\w+_per_warp_active\.pct|sm__cycles_elapsed\.max|smsp__cycles_active\.avg$)")
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if pat.search(h) or (extra and re.search(extra, h)):
            v = r[i]
            if v in ("", "0", "0.000000") and "Kernel" not in h:
                continue
            print(f"{h:95s} {units[i]:>12s} {v}")
