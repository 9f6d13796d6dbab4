#!/usr/bin/env python
"""Developer probe: K1 kernel time vs angle count (1e6 samples, device-resident), store and no-store."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.engine import PreparedCall  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402

n = 1_000_000
b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items()}
for A in [int(a) for a in sys.argv[1:]] or [64, 96, 128, 176, 184, 192, 200, 208, 216, 224, 256, 320, 384, 512]:
    row = []
    for kw in ({}, {'want_j_ion': False}, {'lanes4': True}, {'no_quad': True}):
        call = PreparedCall(b, want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=A, **kw)
        for _ in range(3):
            call.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(10):
            e0.record(); call.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        row.append(float(np.median(ts)))
        del call
    gb = (8 + 144 / A) * n * A / 1e9
    print(f'A={A:4d}  default {row[0]:.3f} ms ({gb / row[0]:.2f} TB/s)  no-store {row[1]:.3f} ms  lanes4 {row[2]:.3f} ms ({gb / row[2]:.2f} TB/s)  no-quad {row[3]:.3f} ms', flush=True)
