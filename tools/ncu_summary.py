#!/usr/bin/env python3
"""Summarise an .ncu-rep: per captured kernel the duration, launch shape, pipe use, DRAM traffic and
the top stall reasons. Usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in keys:
        if k in d:
            print("%-70s %s" % (k, d[k][:90]))
    st = {}
    for h, v in d.items():
        if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued"):
            try:
                st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v.replace(",", ""))
            except ValueError:
                pass
    tot = sum(st.values()) or 1.0
    print("stalls: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
