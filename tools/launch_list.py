#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into profiles/*_launches_*.txt.
   python tools/launch_list.py gpurun_out/launches.csv "header text" > profiles/rNN_launches_x.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
print(f'# {sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"}')
print('# gpu__time_duration.sum per launch, --clock-control none; cold-cache and serialised: compare SHARES, not absolutes')
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split('(')[0].replace('void ', '').replace('hpem::', '')
    key = (name, r[8], r[7])
    tot[key][0] += 1
    tot[key][1] += float(r[14])
total = sum(v[1] for v in tot.values())
print(f'# {len(rows)} launches, {total / 1e6:.3f} ms of kernel time in total\n# kernel, grid, block: launches, total ms, share, mean us')
for (name, grid, block), (cnt, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{name:60s} {grid:>16s} {block:>12s}  {cnt:5d}  {ns / 1e6:9.3f} ms  {100 * ns / total:5.1f} %  {ns / cnt / 1e3:9.2f} us')
print('\n# first 40 launches: launch#, duration_ns, grid, kernel')
for r in rows[:40]:
    print(f'{int(r[0]):4d} {int(float(r[14])):12d} {r[8]:>16s} {r[4].split("(")[0]}')
