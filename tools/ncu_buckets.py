#!/usr/bin/env python
"""Bucket the SASS of an .ncu-rep by address range: samples, executed instructions, opcode mix and main stall per bucket.
   python tools/ncu_buckets.py rep.ncu-rep [bucket_size]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-kernel-base', 'function'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
cols = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] not in ('Address', 'Kernel Name')]
tot = sum(int(r[cols['# Samples']]) for r in data)
toti = sum(int(r[cols['Instructions Executed']]) for r in data)
print(f'{len(data)} SASS instructions, {toti} executed, {tot} samples')
for b in range(0, len(data), bs):
    chunk = data[b:b + bs]
    smp = sum(int(r[cols['# Samples']]) for r in chunk)
    ex = sum(int(r[cols['Instructions Executed']]) for r in chunk)
    ops = collections.Counter()
    st = collections.Counter()
    for r in chunk:
        t = r[cols['Source']].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] += int(r[cols['Instructions Executed']])
        for h, i in cols.items():
            if h.startswith('stall_') and 'Not Issued' not in h:
                st[h[6:]] += int(r[i])
    print(f'{b:5d}-{b + len(chunk) - 1:5d}  samples {100 * smp / tot:5.1f}%  exec {100 * ex / toti:5.1f}%  '
          f'{", ".join(f"{k} {v * 100 // max(ex, 1)}%" for k, v in ops.most_common(4)):48s} | {", ".join(f"{k} {v * 100 // max(smp, 1)}%" for k, v in st.most_common(3))}')
