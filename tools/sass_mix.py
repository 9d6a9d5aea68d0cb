#!/usr/bin/env python
"""Instruction mix and hottest SASS lines of one kernel from `ncu -i rep --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == 'Address')
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if r and r[0].startswith('0x')]
first = data[0][0]
# only the first kernel instance in the file
out = []
for i, r in enumerate(data):
    if i > 0 and r[0] == first:
        break
    out.append(r)
data = out
ops, samp, tot, stot = collections.Counter(), collections.Counter(), 0, 0
for r in data:
    parts = r[idx['Source']].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    n, s = int(r[idx['Instructions Executed']]), int(r[idx['# Samples']])
    ops[op] += n; samp[op] += s; tot += n; stot += s
print('sass lines', len(data), 'warp instructions', tot, 'samples', stot)
for op, n in ops.most_common(22):
    print(f'{op:12s} {n:10d} {100 * n / tot:5.1f}%   samples {100 * samp[op] / max(stot, 1):5.1f}%')
print('--- hottest lines (samples, executed, sass)')
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{r[idx['# Samples']]:>6s} {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']][:90]}")
