#!/usr/bin/env python
"""Hot spots from an `ncu --page source --csv` dump: python profiles/ncu_source_hot.py src.csv [topN]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
isrc, iss, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
num = lambda s: int(s) if s.strip().isdigit() else 0
tot = sum(num(r[iss]) for r in data) or 1
totex = sum(num(r[iex]) for r in data)
print('total stall samples', tot, 'warp instructions', totex, 'SASS lines', len(data))
op, opx = collections.Counter(), collections.Counter()
for r in data:
    t = r[isrc].split()
    o = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '?')).split('.')[0]
    op[o] += num(r[iss]); opx[o] += num(r[iex])
print('by opcode (stall samples):')
for k, v in op.most_common(16):
    print(f'  {k:14s} {v:7d} {100*v/tot:5.1f}%   executed {opx[k]:>10d} ({100*opx[k]/max(totex,1):4.1f}%)')
print('top instructions:')
for i, r in sorted(enumerate(data), key=lambda t: -num(t[1][iss]))[:top]:
    print(f'  #{i:5d} {num(r[iss]):6d} {100*num(r[iss])/tot:5.1f}% ex={r[iex]:>8s}  {r[isrc].strip()[:100]}')
