#!/usr/bin/env python
"""Per-source-line cost of a kernel from an .ncu-rep captured with --import-source on (needs -lineinfo):
warp-instructions executed and stall samples per CUDA source line, grouped by file.  Usage:
   python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--top 40] [--per N]     (N = divide counts by N, e.g. the warp count)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
per = float(sys.argv[sys.argv.index('--per') + 1]) if '--per' in sys.argv else 1.0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--print-kernel-base',
                      'function'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr = None, None
lines = []          # (file, line, source, instr, samples)
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and r[0] not in ('', 'Function Name', 'Kernel Name') and r[0].isdigit():
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[hdr['Instructions Executed']]), int(r[hdr['# Samples']])))
        except (ValueError, IndexError):
            pass
tot_i = sum(x[3] for x in lines)
tot_s = sum(x[4] for x in lines)
print(f'{len(lines)} source lines, {tot_i} warp-instructions ({tot_i / per:.1f} per unit), {tot_s} samples')
by_file = {}
for f, ln, src, ins, smp in lines:
    a = by_file.setdefault(f, [0, 0])
    a[0] += ins
    a[1] += smp
for f, (ins, smp) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f'  {f:28s} instr {100 * ins / tot_i:5.1f}%  samples {100 * smp / max(tot_s, 1):5.1f}%')
print(f'top {top_n} lines by instructions executed:')
for f, ln, src, ins, smp in sorted(lines, key=lambda x: -x[3])[:top_n]:
    print(f'  {f:22s}:{ln:5d} instr {100 * ins / tot_i:5.1f}% ({ins / per:8.1f})  samples {100 * smp / max(tot_s, 1):5.1f}%  {src[:90]}')
