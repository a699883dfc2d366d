#!/usr/bin/env python
"""Key figures of an ncu report (first kernel row): python scripts/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, d = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
for k in want:
    if k in hdr:
        i = hdr.index(k); print(f"{k:70s} {d[i]:>16s} {units[i]}")
for i, k in enumerate(hdr):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[i] or 0) >= 0.3:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {float(d[i]):6.2f}")
