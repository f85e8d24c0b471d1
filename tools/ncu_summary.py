#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep) into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [pattern ...]

Prints, per profiled launch, every raw metric whose name contains one of the patterns (default set below).
"""
import csv
import io
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy",
           "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "op_dfma_pred_on.sum", "op_dmul_pred_on.sum", "op_dadd_pred_on.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second", "local_load", "local_store",
           "smsp__average_warp", "smsp__warps_issue_stalled", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma",
           "sm__inst_executed_pipe_xu", "sm__cycles_active.avg", "smsp__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== launch", r[0], r[hdr.index("Kernel Name")][:70], "block", r[hdr.index("Block Size")], "grid", r[hdr.index("Grid Size")])
        for i, h in enumerate(hdr):
            if any(p in h for p in pats):
                print("  %-95s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()
