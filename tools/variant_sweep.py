#!/usr/bin/env python
"""Developer probe: time the K1u kernel of several compile-time variants (build/variants/libhpem_*.so)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from hallthrusterpem_b200.engine import PreparedCall
from hallthrusterpem_b200.synthetic import spt100_batch
def run(n, A, want_j=True, reps=12):
    b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items()}
    call = PreparedCall(b, want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=A, want_j_ion=want_j)
    for _ in range(3): call.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); call.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
print('%%-12s' %% sys.argv[1], ' '.join('%%s=%%.3f' %% (lbl, run(*cfg)) for lbl, cfg in [
    ('1Mx200', (1000000, 200)), ('1Mx96', (1000000, 96)), ('1Mx64', (1000000, 64)), ('4Mx256', (4000000, 256)), ('1Mx512', (1000000, 512))]), flush=True)
''' % str(ROOT)
for lib in sorted((ROOT / 'build' / 'variants').glob('libhpem_*.so')):
    env = dict(os.environ, HPEM_LIBRARY=str(lib))
    subprocess.run([sys.executable, '-c', CODE, lib.stem.replace('libhpem_', '')], env=env)
