#!/usr/bin/env python
"""Compact per-kernel table from an .ncu-rep (run where ncu is installed):
   python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep > profiles/X_summary.txt"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
     ("lts__t_bytes.sum", "l2_bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
     ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
     ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dsmem"),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
     ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
     ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
     ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
     ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
     ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleep"),
     ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
     ]
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:90], "| id", r[idx["ID"]] if "ID" in idx else "")
    out = []
    for key, short in M:
        if key in idx and r[idx[key]] != "":
            v = r[idx[key]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            out.append(f"{short}={v}{units[idx[key]] if units[idx[key]] not in ('%','') else ''}")
    print("   " + "  ".join(out))
