#!/usr/bin/env python
"""Per-launch device times (ncu launch list) of the LAST captured-graph replay in a bench run: kernel, grid, us.
usage: step_launches.py launches.csv [n_launches_per_step]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
ki, vi, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
data = [r for r in rows[hi + 1:] if len(r) > vi]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 231
# one test() pass starts with the image batch's nchw_to_nhwc kernel: take the last complete pass
idx = [i for i, r in enumerate(data) if 'nchw_to_nhwc_kernel' in r[ki] and i + n <= len(data)]
start = idx[-1] if idx else 0
seq = data[start:start + n]
tot = 0.0
for j, r in enumerate(seq):
    name = r[ki]
    short = name.split('(')[0].split('::')[-1][:28]
    if 'conv_umma_kernel' in name:
        short = 'umma<cell>' if '<(bool)1>' in name or '<1>' in name else 'umma<conv>'
    us = float(r[vi].replace(',', '')) / 1000.0
    tot += us
    print(f"{j:4d} {short:28s} grid {r[gi]:>14s} {us:8.1f} us")
print(f"total {tot:.1f} us over {len(seq)} launches")
