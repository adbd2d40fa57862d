#!/usr/bin/env python3
"""tools/ncu_brief.py <report.ncu-rep> — per kernel launch: duration, DRAM traffic, issue/LSU utilisation, shared-memory
wavefronts and bank conflicts, instruction counts, warp stall samples (the numbers the profiles/*_summary.txt files quote)."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_inst_executed_op_shared_st.sum", "smsp__inst_executed_op_branch.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sector_hit_rate.pct",
]
for r in rows[2:]:
    print("=" * 100)
    print(r[idx["Kernel Name"]], "grid", r[idx.get("Grid Size", 0)], "block", r[idx.get("Block Size", 0)])
    for k in KEYS:
        if k in idx:
            print(f"  {k:86s} {r[idx[k]]:>18s} {units[idx[k]]}")
    st = []
    for h in hdr:
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                st.append((float(r[idx[h]]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print("  warp stall samples: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:9]))
