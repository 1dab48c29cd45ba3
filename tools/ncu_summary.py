#!/usr/bin/env python3
"""Key per-kernel numbers of an .ncu-rep (via `ncu -i rep --page raw --csv`): tools/ncu_summary.py <rep> [...]"""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'launch__grid_size']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h[i]: (r[i], units[i]) for i in range(len(h))}
        print('=====', rep.split('/')[-1], d['Kernel Name'][0][:70])
        for k in KEYS:
            if k in d: print('   %-60s %s %s' % (k, d[k][0], d[k][1]))
        ks = [k for k in d if 'issue_stalled' in k and k.endswith('.ratio') and 'not_issued' not in k]
        top = sorted(ks, key=lambda k: -float(d[k][0].replace(',', '') or 0))[:7]
        print('   stalls (warps per issue-active cycle):', ', '.join('%s %.2f' % (k.split('stalled_')[1].split('_per')[0], float(d[k][0].replace(',', ''))) for k in top))
