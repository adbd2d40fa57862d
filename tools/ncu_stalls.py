#!/usr/bin/env python3
"""tools/ncu_stalls.py <report.ncu-rep> — per kernel: warp stall reasons (> 1.5 % of warp-active) and the instruction mix."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
MIX = ("smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
       "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum", "smsp__inst_executed_op_global_atom.sum",
       "smsp__inst_executed_op_global_red.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
       "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum", "smsp__inst_executed_op_branch.sum",
       "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_op_bar.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
       "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum")
for r in rows[2:]:
    print("=" * 90)
    print(r[idx["Kernel Name"]], r[idx["gpu__time_duration.sum"]], "us" if "gpu__time_duration.sum" in idx else "")
    for h in hdr:
        if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct") and "not_issued" not in h:
            try:
                v = float(r[idx[h]].replace(",", "") or 0)
            except ValueError:
                continue
            if v > 1.5:
                name = h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")
                print(f"  stall {name:28s} {v:8.1f}")
    for h in MIX:
        if h in idx:
            print(f"  {h:50s} {r[idx[h]]}")
