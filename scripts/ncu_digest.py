#!/usr/bin/env python3
"""Digest of one `ncu --set full` capture (CSV of `ncu -i rep --page raw --csv`): the counters DESIGN.md argues with.
usage: scripts/ncu_digest.py <raw.csv> [warp_iterations]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
d = dict(zip(rows[0], rows[2]))
def f(k):
    try: return float(d[k].replace(",", ""))
    except Exception: return float("nan")
keys = [("kernel", "Kernel Name"), ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
        ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"), ("occ_limit_smem", "launch__occupancy_limit_shared_mem"),
        ("occ_limit_regs", "launch__occupancy_limit_registers"), ("time_ms", "gpu__time_duration.sum"),
        ("cycles_elapsed_max", "sm__cycles_elapsed.max"), ("cycles_active_avg", "smsp__cycles_active.avg"),
        ("warp_inst", "smsp__inst_executed.sum"), ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("l1tex_data_pipe_pct_active", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        ("l1tex_wavefronts_pct_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), ("smem_ld_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
        ("smem_st_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"),
        ("smem_ld_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
        ("smem_st_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
        ("global_st_wavefronts", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum"), ("global_st_requests", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"),
        ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active")]
for name, k in keys:
    print(f"{name} = {d.get(k)}")
for k in sorted(d):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        v = f(k)
        if v >= 0.05: print(f"stall {k.split('issue_stalled_')[1].split('_per_issue')[0]} = {v:.2f} per issue")
if len(sys.argv) > 2:
    wi = float(sys.argv[2])
    print(f"per warp-iteration: instructions {f('smsp__inst_executed.sum') / wi:.1f}, shared wavefronts {f('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / wi:.1f} "
          f"(ld {f('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum') / wi:.1f}, st {f('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum') / wi:.1f}), "
          f"global store wavefronts {f('l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum') / wi:.1f}")
