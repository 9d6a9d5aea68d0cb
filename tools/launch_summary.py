#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).  usage: launch_summary.py list.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iN, iM, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
by = collections.OrderedDict()
for r in rows[1:]:
    if r[iM] != 'gpu__time_duration.sum':
        continue
    us = float(r[iV].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iU].replace('second', 's').replace('nsecond', 'ns'), 1e-3)
    name = re.sub(r'^void ', '', r[iN]).replace('prosim::', '')
    name = name.split('(')[0][:60]
    by.setdefault(name, [0, 0.0])
    by[name][0] += 1
    by[name][1] += us
tot = sum(v[1] for v in by.values())
print(f'one forward: {sum(v[0] for v in by.values())} launches, sum {tot / 1e3:.1f} ms')
for name, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f'{name:60s} n={n:4d} total_ms={us / 1e3:7.2f} avg_us={us / n:8.1f} share={100 * us / tot:5.1f}%')
