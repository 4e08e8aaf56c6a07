#!/usr/bin/env python
"""Key metrics of every kernel in an `ncu --set full` report, from `ncu -i rep --page raw --csv`."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active']


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("### `{}`\n".format(r[idx['Kernel Name']]))
        print("| metric | value |\n|---|---|")
        for w in WANT:
            if w in idx:
                print("| {} | {} {} |".format(w, r[idx[w]], units[idx[w]]))
        stalls = [(h[34:-23], float(r[idx[h]].replace(',', '') or 0)) for h in hdr
                  if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
        top = sorted(stalls, key=lambda x: -x[1])[:6]
        print("| top stalls (warps per issue) | {} |\n".format(', '.join('{} {:.2f}'.format(k, v) for k, v in top)))


if __name__ == '__main__':
    main(sys.argv[1])
