#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (needs `ncu` on PATH): python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']


def main(path):
    text = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print('== {} grid {} block {}'.format(d.get('Kernel Name'), d.get('Grid Size'), d.get('Block Size')))
        for k in KEYS:
            if k in d:
                print('  {:<70s} {} {}'.format(k, d[k], u[k]))
        stalls = []
        for h in hdr:
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(d[h]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        print('  stalls per issue: ' + ', '.join('{} {:.2f}'.format(n, v) for v, n in sorted(stalls, reverse=True) if v > 0.05))


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
