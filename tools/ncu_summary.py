#!/usr/bin/env python
"""Summarise an ncu --page raw --csv export: one block of key metrics per captured launch."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('----')
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:82s} {r[i][:90]} {units[i]}")
