#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i … --page raw --csv`) into the handful of numbers the
roofline discussion needs.  Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [substr…]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_fmaheavy.sum",
    "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_xu.sum",
    "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum",
    "smsp__inst_executed_pipe_uniform.sum", "smsp__inst_executed_pipe_cbu.sum",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            if h in KEYS or any(e in h for e in extra):
                print(f"  {h} = {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
