#!/usr/bin/env python
"""Print the handful of ncu metrics the design decisions rest on, per kernel launch, from a
`ncu --page raw --csv` export:  python benchmarks/ncu_pick.py export.csv [kernel-substring]"""
import csv
import sys

PICK = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
]
rows = list(csv.reader(open(sys.argv[1], newline="")))
h = next(r for r in rows if "Kernel Name" in r)
u = rows[rows.index(h) + 1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
for r in rows[rows.index(h) + 2 :]:
    name = r[h.index("Kernel Name")]
    if want not in name:
        continue
    print("##", name[:110], "id", r[0])
    for m in PICK:
        if m in h:
            print(f"   {m:90s} {r[h.index(m)]:>14s} {u[h.index(m)]}")
