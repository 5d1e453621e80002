#!/usr/bin/env python3
"""Condense an `ncu --set full` report into the few lines kept under profiles/ (reads `ncu -i <rep> --page raw --csv`).
usage: scripts/summarize_ncu.py gpurun_out/prof_<tag>.ncu-rep "<header line>" > profiles/<tag>_fill_systolic_summary.txt"""
import csv, io, re, subprocess, sys

KEEP = re.compile(r"^(Kernel Name|dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|launch__(block_size|grid_size|occupancy_limit|registers_per_thread|"
                  r"shared_mem_per_block_dynamic)|sm__cycles_elapsed\.max|sm__inst_executed_pipe_(alu|lsu)\.avg\.pct|"
                  r"sm__pipe_fma_cycles_active\.avg\.pct|sm__warps_active\.avg\.pct|smsp__average_warps_issue_stalled_.*_per_issue_active|"
                  r"smsp__cycles_active\.avg$|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct)")
rep, header = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
names, units, vals = rows[0], rows[1], rows[2]
print("# " + header)
for n, u, v in sorted(zip(names, units, vals)):
    if KEEP.match(n):
        print(f"{n} [{u}] = {v}")
