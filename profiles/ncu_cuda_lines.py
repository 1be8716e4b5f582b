#!/usr/bin/env python
"""Per-CUDA-line totals from `ncu --page source --csv --print-source cuda,sass`:
   python profiles/ncu_cuda_lines.py src_cuda.csv [topN]   (lists lines by executed warp instructions and stall samples)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
num = lambda s: int(s) if s.strip().lstrip('-').isdigit() else 0
out, fname, seen_kernel = [], '', 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif len(r) >= 2 and r[0] == 'Function Name':
        seen_kernel += 1
    elif len(r) >= 8 and r[0].strip().isdigit():
        out.append((fname, int(r[0]), r[1].strip(), num(r[4]), num(r[7])))
# the dump repeats per profiled launch; fold identical (file,line)
agg = {}
for f, ln, src, st, ex in out:
    k = (f, ln)
    a = agg.setdefault(k, [src, 0, 0])
    a[1] += st; a[2] += ex
tot_st = sum(a[1] for a in agg.values()) or 1
tot_ex = sum(a[2] for a in agg.values()) or 1
print(f'lines {len(agg)}  stall samples {tot_st}  warp instructions {tot_ex}')
print('--- by executed instructions')
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f'{100*a[2]/tot_ex:5.1f}% ex  {100*a[1]/tot_st:5.1f}% st  {f}:{ln:<4d} {a[0][:95]}')
print('--- by stall samples')
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{100*a[1]/tot_st:5.1f}% st  {100*a[2]/tot_ex:5.1f}% ex  {f}:{ln:<4d} {a[0][:95]}')
