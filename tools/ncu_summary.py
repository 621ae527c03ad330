#!/usr/bin/env python
"""Summarise .ncu-rep captures (`ncu --set full`) into one JSON for profiles/: usage ncu_summary.py out.json rep1 [rep2 ...]"""
import csv, json, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
out = []
for f in sys.argv[2:]:
    txt = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {'report': f.split('/')[-1], 'Kernel Name': vals[hdr.index('Kernel Name')]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = [vals[i], units[i]]
        out.append(d)
json.dump(out, open(sys.argv[1], 'w'), indent=1)
print('wrote', sys.argv[1], len(out), 'kernels')
