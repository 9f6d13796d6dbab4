#!/usr/bin/env python
"""Developer probe: max relative difference between the recurrence kernels (K1u / quad mode) and the reference-order
direct kernel (K1d) over 1e6 samples, overall and for profile exponents below 40."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.models import current_density  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402

n = 1_000_000
host = spt100_batch(n, 2024)
b = {k: torch.as_tensor(v, device='cuda:0') for k, v in host.items()}
for A in (91, 200, 255, 512):
    fast = current_density(b, 1.0, n_angles=A, extras=True)
    direct = current_density(b, 1.0, n_angles=A, extras=True, direct=True)
    rel = ((fast['j_ion'] - direct['j_ion']).abs() / direct['j_ion'].abs())
    a1 = torch.clamp(b['c2'] * b['P_b'] * 133.322 + b['c3'], max=np.pi / 2)
    theta = torch.linspace(0, np.pi / 2, A, dtype=torch.float64, device='cuda:0')
    arg = (theta[None, :] / a1[:, None]) ** 2
    small = arg < 40
    cd = ((fast['cos_div'] - direct['cos_div']).abs() / direct['cos_div'].abs()).max().item()
    print(f'A={A:4d}: max rel j_ion {rel.max().item():.3e}  (exponent < 40: {rel[small].max().item():.3e})  '
          f'99.99th pct {torch.quantile(rel.flatten()[::97], 0.9999).item():.3e}  cos_div {cd:.3e}', flush=True)
