#!/usr/bin/env python
"""Smoke / sanitizer target: one small call of every kernel family (odd/even angle counts, several radii, reduce-only,
likelihood, compression) with results checked for finiteness only."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.compression import SVD  # noqa: E402
from hallthrusterpem_b200.likelihood import JionMeasurements, jion_log_likelihood  # noqa: E402
from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments  # noqa: E402
from hallthrusterpem_b200.models import current_density, plume_cathode  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402

n = 1003
b = spt100_batch(n, 3)
d = {k: torch.as_tensor(v, device='cuda:0') for k, v in b.items()}
for A in (91, 93, 66, 200, 51, 130, 255):
    for kw in ({}, {'lanes4': True}, {'no_quad': True}, {'direct': True}):
        o = plume_cathode(d, 1.0, n_angles=A, extras=True, **kw)
        assert torch.isfinite(o['j_ion']).all()
    o = plume_cathode(b, 1.0, n_angles=A)                       # host pipeline
    assert np.isfinite(o['j_ion']).all()
for R in (3, 9, 25):
    o = current_density(d, np.linspace(1.0, 1.3, R), n_angles=91, extras=True)
    assert torch.isfinite(o['j_ion']).all()
mc = MonteCarloMoments(n_angles=91, device=0, hist=HistogramSpec(angle_stride=4, sub_bits=2))
mc.accumulate(d)
mc.accumulate_sampled(5000, 1, 0)
assert mc.result().n_samples == n + 5000
meas = JionMeasurements(np.linspace(-1.5, 1.5, 17), np.full(17, 1.0), np.full(17, 0.1), n_angles=91, device=0)
assert torch.isfinite(jion_log_likelihood(d, meas, torr=133.322)).all()
c = SVD.from_samples({k: v[:300] for k, v in b.items() if k != 'T'}, n_angles=91, device=0, rank=5)
z = c.compress_inputs({k: v for k, v in d.items() if k != 'T'})
assert torch.isfinite(c.reconstruct_field(z)).all() and torch.isfinite(c.compress_field(c.reconstruct_field(z))).all()
torch.cuda.synchronize()
print('sanitize target ok')
