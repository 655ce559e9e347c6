"""Print the handful of ncu metrics that decide what bounds a kernel (reads a .ncu-rep with `ncu -i --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
pick = sys.argv[2] if len(sys.argv) > 2 else ''
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if pick and pick not in d['Kernel Name']:
        continue
    print('==', d['Kernel Name'][:70], 'grid', d.get('launch__grid_size'))
    for k in KEYS:
        if k in d:
            print('   %-72s %s' % (k, d[k]))
    stalls = sorted(((float(d[k] or 0), k) for k in hdr if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio')), reverse=True)
    print('   stalls/issue: ' + ', '.join('%s=%.2f' % (k.split('stalled_')[1].split('_per_issue')[0], v) for v, k in stalls[:7]))
