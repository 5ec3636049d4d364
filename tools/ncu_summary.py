#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into one line per kernel launch (the metrics the
roofline discussion in DESIGN.md uses).  Usage: python tools/ncu_summary.py raw.csv [> profiles/x.txt]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
col = {n: i for i, n in enumerate(hdr)}
want = [("ms", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
        ("warps_act%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("fp64pipe%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("issue%", "sm__inst_executed.avg.pct_of_peak_sustained_active") if "sm__inst_executed.avg.pct_of_peak_sustained_active" in col else ("ipc", "sm__inst_executed.avg.per_cycle_active"),
        ("L1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), ("L2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
        ("st_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("st_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("st_noinst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("st_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("st_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
        ("st_dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
        ("st_notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
        ("smem_wave", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), ("inst", "smsp__inst_executed.sum"),
        ("local_ld", "smsp__inst_executed_op_local_ld.sum"), ("local_st", "smsp__inst_executed_op_local_st.sum")]
want = [(a, b) for a, b in want if b in col]
for r in data:
    name = r[col["Kernel Name"]]
    m = re.search(r"(\w+)<(.*)>", name)
    short = ("%s<%s>" % (m.group(1), m.group(2).replace("(int)", "").replace(" ", ""))) if m else name[:40]
    out = []
    for a, b in want:
        v = r[col[b]].replace(",", "")
        try:
            f = float(v)
            u = units[col[b]] if b in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum") else ""
            out.append("%s=%s%s" % (a.replace("_MB", ""), ("%.3g" % f), u))
        except ValueError:
            out.append("%s=%s" % (a, v))
    print(short, " ".join(out))
