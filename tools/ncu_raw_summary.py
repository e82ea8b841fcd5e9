#!/usr/bin/env python3
"""Summarise the raw page of an ncu --set full capture (CSV made on the GPU box with `ncu -i x.ncu-rep --page raw --csv`)
into the counters the roofline discussion needs, one block per captured launch.
Usage: tools/ncu_raw_summary.py gpurun_out/x_full_raw.csv > profiles/x_kernels.summary.txt"""
import csv, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h, u = rows[0], rows[1]
    only_first = len(sys.argv) > 2 and sys.argv[2] == "--first"
    seen = set()
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0]
        grid = r[h.index("Grid Size")]
        if only_first and (name, grid) in seen:
            continue
        seen.add((name, grid))
        print("== kernel:", name, "grid", grid, "block", r[h.index("Block Size")])
        for k in KEYS:
            if k in h:
                print("  %-86s %-14s %s" % (k, u[h.index(k)], r[h.index(k)]))
        rd = float(r[h.index("dram__bytes_read.sum")].replace(",", "")); wr = float(r[h.index("dram__bytes_write.sum")].replace(",", ""))
        print("  traffic (read+write) per launch: %.3f %s" % (rd + wr, u[h.index("dram__bytes_read.sum")]))


if __name__ == "__main__":
    main()
