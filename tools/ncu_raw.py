#!/usr/bin/env python3
"""Prints selected raw metrics of every kernel in an ncu report.  usage: ncu_raw.py report.ncu-rep [extra-substring ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
names = [r[hdr.index("Kernel Name")].split("(")[0][-28:] for r in rows[2:]]
print("metric".ljust(78), " | ".join(n.rjust(18) for n in names))
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h) or any(e in h for e in extra):
        vals = [r[i] for r in rows[2:]]
        try:
            if all(float(v) < 0.05 for v in vals) and "stalled" in h: continue
        except ValueError: pass
        print((h + " [" + units[i] + "]").ljust(78)[:78], " | ".join(v[:18].rjust(18) for v in vals))
