#!/usr/bin/env python
"""Per-kernel SASS evidence: counts of the instructions that prove the TMA / bulk-copy store paths, the histogram
reductions and the absence of tensor-core / library code.  Runs without a GPU (cuobjdump on the built library).
   python tools/sass_evidence.py > profiles/r02_sass_tma.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200 import _lib  # noqa: E402

lib = str(_lib.build_library())
elf = subprocess.run(['cuobjdump', '-lelf', lib], capture_output=True, text=True).stdout.split('\n')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
print('# cuobjdump -sass hallthrusterpem_b200/lib/libhpem.so (sm_100a only): per-kernel counts of the instructions that prove the')
print('# TMA / bulk-copy store paths, the histogram reductions, the quadrature-table loads (LDG.E.128.CONSTANT in the reduce-only')
print('# kernel) and the absence of tensor-core / library code.  Regenerate: python tools/sass_evidence.py')
print('# arch list:', ' '.join(x.strip() for x in elf if x.strip()))
pat = {'fp64': r'\b(DFMA|DMUL|DADD|DSETP|DMNMX)\b', 'UTMASTG': r'\bUTMASTG', 'UBLKCP': r'\bUBLKCP', 'FENCE.VIEW.ASYNC': r'FENCE\.VIEW\.ASYNC',
       'RED': r'\bREDG?\.', 'LDG.128': r'\bLDG\.E\.128', 'LDGSTS': r'\bLDGSTS', 'tensor-core': r'\b(HMMA|IMMA|DMMA|UTCMMA|UTCHMMA|QMMA|OMMA)\b'}
blocks = re.split(r'\n\s*Function : ', sass)[1:]
rows = []
for i, blk in enumerate(blocks):
    body = [ln for ln in blk.split('\n') if re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s', ln)]
    cnt = collections.OrderedDict((k, sum(1 for ln in body if re.search(p, ln))) for k, p in pat.items())
    rows.append((names[i] if i < len(names) else blk.split('\n')[0], len(body), cnt))
for name, n, cnt in sorted(rows):
    print(f'{name[:150]:150s} instr {n:6d}  ' + '  '.join(f'{k} {v:4d}' for k, v in cnt.items()))
tot = collections.Counter()
for _, _, cnt in rows:
    tot.update(cnt)
print('# totals: ' + '  '.join(f'{k} {v}' for k, v in tot.items()))
