#!/usr/bin/env python
"""Summarise an Nsight Compute report (read here, no GPU needed) into the few numbers the roofline argument uses.

    python profiles/summarize.py gpurun_out/prof.ncu-rep [> profiles/<name>.summary.txt]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:85s} {r[i]:>18s} {units[i]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
