#!/usr/bin/env python3
"""Summarise an .ncu-rep (read offline with `ncu -i`) into the handful of counters the roofline needs.
Usage: tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        print("== kernel:", name, "grid", r[h.index("Grid Size")], "block", r[h.index("Block Size")])
        for k in KEYS:
            if k in h:
                print("  %-86s %-14s %s" % (k, u[h.index(k)], r[h.index(k)]))
        rd = float(r[h.index("dram__bytes_read.sum")]); wr = float(r[h.index("dram__bytes_write.sum")])
        print("  traffic (read+write) per launch: %.3f %s" % (rd + wr, u[h.index("dram__bytes_read.sum")]))


if __name__ == "__main__":
    main()
