#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into the few numbers that matter here.  Usage:
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 22
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_lsu.sum', 'launch__grid_size', 'launch__block_size']
for i, h in enumerate(hdr):
    if h in want or re.match(r'smsp__average_warps_issue_stalled_.*_per_issue_active.ratio', h):
        vals = [r[i] for r in rows[2:]]
        if h.startswith('smsp__average_warps') and all(float(v or 0) < 0.3 for v in vals):
            continue
        print(f'{h:90s} {units[i]:14s} {vals}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-kernel-base', 'function'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
cols = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) >= len(hdr) and r[0] != 'Address':
        data.append(r)
ie, ismp, isrc = cols['Instructions Executed'], cols['# Samples'], cols['Source']
tot_s = sum(int(r[ismp]) for r in data)
tot_i = sum(int(r[ie]) for r in data)
print(f'\nSASS instructions {len(data)}, warp-instructions executed {tot_i}, samples {tot_s}')
ops = collections.Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op] += int(r[ie])
print('opcode mix:', ', '.join(f'{k} {100 * v / tot_i:.1f}%' for k, v in ops.most_common(14)))
st = collections.Counter()
for r in data:
    for h, i in cols.items():
        if h.startswith('stall_') and 'Not Issued' not in h:
            st[h] += int(r[i])
tt = sum(st.values())
print('stall mix:', ', '.join(f'{k[6:]} {100 * v / tt:.1f}%' for k, v in st.most_common(9)))
print(f'\ntop {top_n} instructions by stall samples:')
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:top_n]):
    r = data[i]
    stl = sorted(((h[6:], int(r[c])) for h, c in cols.items() if h.startswith('stall_') and 'Not Issued' not in h and int(r[c]) > 0),
                 key=lambda kv: -kv[1])[:2]
    print(f'{i:5d} {100 * int(r[ismp]) / tot_s:5.1f}%  {r[isrc].strip()[:64]:64s} {stl}')
