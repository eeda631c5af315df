#!/usr/bin/env python
"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics DESIGN.md
and profiles/ quote.  Usage: python tools/ncu_summary.py gpurun_out/prof_solve.ncu-rep [substring ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "smsp__inst_executed_pipe_fp64",
        "sm__inst_executed_pipe_lsu", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor", "lts__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__throughput.avg.pct_of_peak_sustained"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for h, u in zip(hdr, units):
            if any(k in h for k in KEYS + extra):
                print("  %-90s %16s %s" % (h, d[h], u))


if __name__ == "__main__":
    main()
