#!/usr/bin/env python
"""Developer probe: multi-radius kernel time vs number of radii (4e5 / R samples, 91 and 200 angles)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.engine import PreparedCall  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402

for A in (91, 200):
    for R in (2, 3, 4, 5, 6, 7, 8, 16, 25):
        n = max(20000, 4_000_000 // (A * R) * 10)
        b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1, c3_test_range=True).items()}
        call = PreparedCall(b, want_cathode=False, want_plume=True, sweep_radius=np.linspace(1.0, 1.3, R), n_angles=A)
        for _ in range(3):
            call.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(7):
            e0.record(); call.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f'A={A:4d} R={R:3d} n={n:7d}: {ms:7.3f} ms  {n * A * R * 8 / ms / 1e9:6.2f} TB/s', flush=True)
        del call, b
