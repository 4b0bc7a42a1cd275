#!/usr/bin/env python
"""Print the handful of ncu raw metrics the roofline discussion needs.  usage: ncu -i rep --page raw --csv | this"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    if w in h:
        i = h.index(w)
        print(f"{w:70s}", [r[i] for r in rows[1:]][-3:])
