#!/usr/bin/env python
"""Developer timing probe (not the contract bench): kernel time of the fused path on device-resident buffers."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.models import plume_cathode  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402


def time_dev(n, A, direct=False, want_j=True, reps=10, cathode=True, no_tma=False):
    from hallthrusterpem_b200.engine import PreparedCall
    b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items()}
    call = PreparedCall(b, want_cathode=cathode, want_plume=True, sweep_radius=1.0, n_angles=A, direct=direct,
                        want_j_ion=want_j, no_tma=no_tma)
    for _ in range(3):
        call.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        call.run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    byts = (8 + 144 / A) * n * A if want_j else 144 * n
    print(f'n={n:>9} A={A:>4} direct={direct!s:5} no_tma={no_tma!s:5} store_j={want_j!s:5}  {ms:8.3f} ms  {n * A / ms / 1e6:10.2f} Geval/s  '
          f'{byts / ms / 1e6:8.1f} GB/s (algorithmic)  min {min(ts):.3f} ms', flush=True)


def time_host(n, A, reps=3):
    b = spt100_batch(n, 1)
    pinned = {k: torch.as_tensor(v).pin_memory().numpy() for k, v in b.items()}
    for label, inp in (('pageable', b), ('pinned', pinned)):
        out = plume_cathode(inp, 1.0, n_angles=A)
        ts = []
        for _ in range(reps):
            del out
            t0 = time.perf_counter()
            out = plume_cathode(inp, 1.0, n_angles=A)
            ts.append(time.perf_counter() - t0)
        t = min(ts)
        print(f'host[{label}] n={n} A={A}: {t * 1e3:.2f} ms  {n * A / t / 1e9:.2f} Geval/s  '
              f'D2H {(A + 3) * 8 * n / t / 1e9:.1f} GB/s', flush=True)


if __name__ == '__main__' and '--multi' not in sys.argv:
    for n, A in ((1_000_000, 200), (1_000_000, 91), (4_000_000, 256), (1_000_000, 512)):
        time_dev(n, A)
        time_dev(n, A, no_tma=True)
        time_dev(n, A, want_j=False)
    time_dev(1_000_000, 200, direct=True)
    time_host(1_000_000, 200)


def time_multi_radius(n=100_000, A=91, R=25, reps=5):
    from hallthrusterpem_b200.engine import PreparedCall
    b = {k: torch.as_tensor(v, device='cuda:0') for k, v in spt100_batch(n, 1).items()}
    radii = np.linspace(1.0, 1.2, R)
    for label, kw in (('K1r', {}), ('K1r-stg', {'no_tma': True}), ('K1d', {'direct': True})):
        call = PreparedCall(b, want_cathode=False, want_plume=True, sweep_radius=radii, n_angles=A, **kw)
        for _ in range(2):
            call.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            e0.record(); call.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f'multi-radius {label:8s} n={n} A={A} R={R}: {ms:.3f} ms  {n * A * R / ms / 1e6:.1f} G(sample x angle x radius)/s  '
              f'{n * A * R * 8 / ms / 1e6:.0f} GB/s', flush=True)


if __name__ == '__main__' and '--multi' in sys.argv:
    time_multi_radius()
    time_multi_radius(A=200, R=8)
