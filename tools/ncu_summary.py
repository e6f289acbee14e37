#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) per kernel: duration, occupancy, pipe use, stall reasons, memory."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.per_cycle_active", "sm__cycles_active.avg", "gpc__cycles_elapsed.max",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("=" * 100)
    print(r[idx["Kernel Name"]][:90])
    for k in keys:
        if k in idx:
            print("  %-75s %s %s" % (k, r[idx[k]], units[idx[k]]))
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  stalls (warps per issue-active cycle):", ", ".join("%s=%.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
