#!/usr/bin/env python
"""Developer probe: reduce-only pass (K2) throughput.  python tools/k2_time.py [A ...]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402

PEAK = 1.8544e13
for A in [int(a) for a in sys.argv[1:]] or [91, 256, 512]:
    n = int(float(os.environ.get('K2_N', 8_000_000 if A <= 256 else 4_000_000)))
    modes = (('sampled+hist8', HistogramSpec(angle_stride=8), True), ('sampled nohist', HistogramSpec(angle_stride=0), True),
             ('arrays+hist8', HistogramSpec(angle_stride=8), False))
    for label, hist, sampled in modes[:int(os.environ.get('K2_MODES', 3))]:
        mc = MonteCarloMoments(n_angles=A, device=0, hist=hist)
        if sampled:
            run = lambda: mc.accumulate_sampled(n, 7, 0)   # noqa: E731
        else:
            b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(min(n, 2_000_000), 1).items()}
            nb = len(b['P_b'])
            run = lambda: mc.accumulate(b)                 # noqa: E731
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        nn = n if sampled else nb
        rate = nn * A / ms * 1e3
        print(f'K2 A={A:4d} {label:15s} n={nn:9d}  {ms:8.3f} ms  {rate / 1e9:8.1f} Geval/s  frac(10 instr/eval) {10 * rate / PEAK:.3f}  '
              f'{ms * 1e3 / (nn / (148 * 768)):7.2f} us/round', flush=True)
