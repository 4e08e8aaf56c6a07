#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean time and share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    H = rows[hdr]
    ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    d = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(r[ui], v)
        d[r[ki]].append(v)
    tot = sum(sum(v) for v in d.values())
    print("| kernel | launches | mean us | share |\n|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("| `{}` | {} | {:.1f} | {:.1f}% |".format(k[:90], len(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == '__main__':
    main(sys.argv[1])
