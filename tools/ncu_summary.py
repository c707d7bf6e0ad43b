#!/usr/bin/env python
"""ncu report -> the text summary committed under profiles/ (same metric list as the round-1/2 summaries).
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt"""
import csv
import subprocess
import sys

KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum dram__bytes.sum.per_second
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed dram__cycles_active.avg.pct_of_peak_sustained_elapsed
lts__t_sectors_srcunit_tex_op_read.sum lts__t_sectors_srcunit_tex_op_write.sum lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum launch__registers_per_thread launch__grid_size launch__block_size launch__waves_per_multiprocessor
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio smsp__average_warp_latency_issue_stalled_membar.ratio
smsp__average_warp_latency_issue_stalled_barrier.ratio smsp__average_warp_latency_issue_stalled_lg_throttle.ratio
smsp__average_warp_latency_issue_stalled_wait.ratio smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio
smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio smsp__average_warp_latency_issue_stalled_not_selected.ratio
smsp__average_warp_latency_issue_stalled_sleeping.ratio smsp__average_warp_latency_issue_stalled_no_instruction.ratio
smsp__average_warp_latency_issue_stalled_branch_resolving.ratio smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio""".split()

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units = rows[0], rows[1]
for r in rows[2:]:
    m = dict(zip(head, r))
    print("kernel:", m.get("Kernel Name", "?").replace("nlbm::", ""))
    for k in KEYS:
        if k in m:
            print(f"  {k:<86s} {m[k]:>18s} {units[head.index(k)]}")
