#!/usr/bin/env python
"""Print the key metrics of .ncu-rep files (first kernel of each): usage tools/ncu_keys.py a.ncu-rep [b.ncu-rep ...]"""
import csv, subprocess, sys
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('===', f)
        for i, h in enumerate(hdr):
            if h in WANT or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
                try:
                    if h.startswith('smsp__average_warps_issue_stalled') and float(vals[i]) < 0.3:
                        continue
                except ValueError:
                    pass
                print(f'{h:95s} {vals[i][:50]:>24s} {units[i]}')
