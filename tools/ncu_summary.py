#!/usr/bin/env python
"""Key metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` dump.  usage: ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
WANT = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get('Kernel Name'), d.get('Grid Size'), d.get('Block Size'))
    for k in WANT:
        if k in d: print('   %-70s %s %s' % (k, d[k], rows[1][hdr.index(k)]))
    st = [(float(d[k]), k) for k in hdr if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio')]
    print('   stalls/issue: ' + ', '.join('%s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, k in sorted(st, reverse=True)[:8]))
