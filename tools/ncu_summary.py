#!/usr/bin/env python
"""Raw-page metrics of the first kernel in an .ncu-rep as `metric,unit,value` rows (what profiles/*_ncu_summary.csv hold).
usage: python tools/ncu_summary.py report.ncu-rep > summary.csv"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
print("# kernel: %s" % name)
keep = ("dram__", "gpu__time", "l1tex__data_bank", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "l1tex__t_sector_hit", "lts__t_sector_hit", "lts__t_bytes", "launch__", "sm__throughput", "sm__warps_active",
        "sm__inst_executed", "smsp__inst_executed.sum", "smsp__issue_active", "smsp__pcsamp_warps_issue_stalled", "smsp__thread_inst_executed_per_inst", "sm__cycles_active.avg", "smsp__warps_eligible")
for h, u, v in zip(hdr, units, vals):
    if h.startswith(keep):
        print("%s,%s,%s" % (h, u, v))
