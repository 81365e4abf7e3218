#!/usr/bin/env python3
"""Print the handful of ncu metrics we steer by from a .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__occupancy_limit_registers', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fp64.sum']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', path, r[head.index('Kernel Name')][:60])
        for i, n in enumerate(head):
            if n in WANT or ('issue_stalled' in n and n.endswith('per_issue_active.ratio') and float(r[i] or 0) > 0.3):
                print('   %-86s %14s %s' % (n, r[i], units[i]))
