"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / bench.py quote."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'sm__cycles_elapsed.avg']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                print('  %-62s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = [(float(r[i]), h) for i, h in enumerate(hdr)
              if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]]
        print('  stalls (warps per issue-active):',
              ', '.join('%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
                        for v, h in sorted(st, reverse=True)[:6]))


if __name__ == '__main__':
    main(sys.argv[1])
