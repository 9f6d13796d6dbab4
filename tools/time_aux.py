#!/usr/bin/env python
"""Developer timing probe for the non-materialising kernels: K2 (moments), K3 (log-likelihood), K4 (latent), K5."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def k2(n=4_000_000, A=256):
    b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items()}
    for label, kw, sampled in (
            ('arrays hist8 cath thrust', dict(hist=HistogramSpec(angle_stride=8)), False),
            ('sampled hist8 cath thrust', dict(hist=HistogramSpec(angle_stride=8)), True),
            ('arrays nohist cath thrust', dict(hist=HistogramSpec(angle_stride=0)), False),
            ('sampled nohist cath thrust', dict(hist=HistogramSpec(angle_stride=0)), True),
            ('arrays nohist plume only', dict(hist=HistogramSpec(angle_stride=0), want_cathode=False, want_thrust=False), False),
            ('arrays hist1 cath thrust sub2', dict(hist=HistogramSpec(angle_stride=1, sub_bits=1)), False),
    ):
        try:
            mc = MonteCarloMoments(n_angles=A, device=0, **kw)
            fn = (lambda: mc.accumulate_sampled(n, 7, 0)) if sampled else (lambda: mc.accumulate(b))
            ms = timeit(fn)
            print(f'K2 n={n} A={A} {label:32s} {ms:8.3f} ms  {n * A / ms / 1e6:8.1f} Geval/s', flush=True)
        except Exception as exc:  # noqa: BLE001
            print(f'K2 {label}: {exc}', flush=True)


def k3(n=1_000_000, A=200, m=64):
    from hallthrusterpem_b200.likelihood import JionMeasurements, jion_log_likelihood
    b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items()}
    rng = np.random.default_rng(0)
    theta = rng.uniform(-1.5, 1.5, m)
    meas = JionMeasurements(theta, 10 ** rng.uniform(-2, 1, m), np.full(m, 0.1), n_angles=A, device=0)
    ms = timeit(lambda: jion_log_likelihood(b, meas, torr=133.322))
    print(f'K3 n={n} A={A} m={m}: {ms:8.3f} ms  {n * A / ms / 1e6:8.1f} Geval/s', flush=True)


def k45(n=1_000_000, A=200):
    from hallthrusterpem_b200.compression import SVD
    fit = {k: v for k, v in spt100_batch(500, 77).items() if k != 'T'}
    b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items() if k != 'T'}
    for kw in (dict(reconstruction_tol=0.01), dict(rank=8), dict(rank=16)):
        c = SVD.from_samples(fit, n_angles=A, torr=133.322, device=0, **kw)
        ms = timeit(lambda: c.compress_inputs(b, torr=133.322))
        z = c.compress_inputs(b, torr=133.322)
        ms5 = timeit(lambda: c.reconstruct_field(z))
        j = c.reconstruct_field(z)
        ms4f = timeit(lambda: c.compress_field(j))
        print(f'K4 n={n} A={A} rank={c.rank}: fused {ms:8.3f} ms ({n * A / ms / 1e6:7.1f} Geval/s)  K5 reconstruct {ms5:8.3f} ms '
              f'({n * A * 8 / ms5 / 1e6:7.1f} GB/s)  K4f field {ms4f:8.3f} ms ({n * A * 8 / ms4f / 1e6:7.1f} GB/s)', flush=True)


if __name__ == '__main__':
    what = sys.argv[1:] or ['k2', 'k3', 'k45']
    if 'k2' in what:
        k2()
        k2(A=91)
    if 'k3' in what:
        k3()
    if 'k45' in what:
        k45()
        k45(A=91)
