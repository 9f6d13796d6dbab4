#!/usr/bin/env python
"""ncu target: a few launches of ONE kernel family.  python tools/prof_target.py {k1u|k1q|k2|k2s|k3|k4|k5|k1w}
(k1u: 1e6 x 200, (n, A) tensor stores; k1q: 1e6 x 91, quad-row tensor stores; k1w: 1e5 x 91 x 25 radii)"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.engine import PreparedCall  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'k1u'
TORR = 133.322


def dev(b):
    return {k: torch.as_tensor(v, device='cuda:0') for k, v in b.items()}


if what in ('k1u', 'k1q'):
    A = 200 if what == 'k1u' else 91
    call = PreparedCall(dev(spt100_batch(1_000_000, 1)), want_cathode=True, want_plume=True, sweep_radius=1.0, n_angles=A)
    for _ in range(3):
        call.run()
elif what in ('k2', 'k2s'):
    from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
    n = 8_000_000 if what == 'k2s' else 2_000_000
    mc = MonteCarloMoments(n_angles=int(os.environ.get('K2_A', 256)), device=0, hist=HistogramSpec(angle_stride=8))
    b = dev(spt100_batch(min(n, 2_000_000), 1))
    for _ in range(3):
        mc.accumulate_sampled(n, 7, 0) if what == 'k2s' else mc.accumulate(b)
elif what == 'k3':
    from hallthrusterpem_b200.likelihood import JionMeasurements, jion_log_likelihood
    rng = np.random.default_rng(0)
    m = 64
    meas = JionMeasurements(rng.uniform(-1.5, 1.5, m), 10 ** rng.uniform(-2, 1, m), np.full(m, 0.1), n_angles=91, device=0)
    b = dev({k: v for k, v in spt100_batch(1_000_000, 1).items() if k not in ('T', 'V_a', 'T_e', 'V_vac', 'Pstar', 'P_T')})
    for _ in range(3):
        jion_log_likelihood(b, meas, torr=TORR)
elif what in ('k4', 'k5'):
    from hallthrusterpem_b200.compression import SVD
    b = {k: v for k, v in spt100_batch(1_000_000, 1).items() if k != 'T'}
    comp = SVD.from_samples({k: v[:500] for k, v in b.items()}, n_angles=200, torr=TORR, device=0, reconstruction_tol=0.01)
    d = dev(b)
    z = comp.compress_inputs(d, torr=TORR)
    for _ in range(3):
        z = comp.compress_inputs(d, torr=TORR) if what == 'k4' else comp.reconstruct_field(z) * 0 + z
elif what == 'k1w':
    call = PreparedCall(dev(spt100_batch(100_000, 1, c3_test_range=True)), want_cathode=False, want_plume=True,
                        sweep_radius=np.linspace(1.0, 1.2, 25), n_angles=91)
    for _ in range(3):
        call.run()
torch.cuda.synchronize()
print('done', what)
