#!/usr/bin/env python
"""Developer probe: K1u with the branch-free per-sample back end (default) vs libdevice (no_fastmath), 1e6 samples."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.engine import PreparedCall  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402

n = 1_000_000
b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items()}
for A in [int(a) for a in sys.argv[1:]] or [33, 64, 91, 100, 128, 200, 256, 512]:
    row = []
    for kw in ({}, {'no_fastmath': True}, {'want_j_ion': False}, {'want_j_ion': False, 'no_fastmath': True}):
        call = PreparedCall(b, want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=A, **kw)
        for _ in range(3):
            call.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(15):
            e0.record(); call.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        row.append(float(np.median(ts)))
        del call
    gb = (8 + 144 / A) * n * A / 1e9
    print(f'A={A:4d}  fast {row[0]:.4f} ms ({gb / row[0]:.2f} TB/s, {gb / row[0] / 6.4463:.3f})  libdevice {row[1]:.4f} ms ({gb / row[1]:.2f} TB/s)'
          f'  no-store: fast {row[2]:.4f} ms  libdevice {row[3]:.4f} ms', flush=True)
