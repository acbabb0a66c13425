#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the text block kept under profiles/: one section per captured kernel
with the metrics DESIGN.md / bench.py cite.   python tools/ncu_summary.py report.ncu-rep [title] > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum")


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(title)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("----")
        print("Kernel Name [] =", d.get("Kernel Name"))
        for i, k in enumerate(hdr):
            if k in KEEP or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
                print(f"{k} [{units[i]}] = {r[i]}")


if __name__ == "__main__":
    main()
