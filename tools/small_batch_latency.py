#!/usr/bin/env python
"""Developer probe: end-to-end latency of the drop-in call (NumPy in -> NumPy out) for the small batches amisc typically
passes, next to the NumPy oracle on one core."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.models import cathode_coupling, current_density  # noqa: E402
from hallthrusterpem_b200.synthetic import spt100_batch  # noqa: E402
from oracle.ref_restated import cathode_coupling_oracle, current_density_oracle  # noqa: E402


def t_of(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


for n in (1, 100, 1000, 10_000, 100_000, 1_000_000):
    b = spt100_batch(n, 5)
    reps = 200 if n <= 10_000 else 20
    t_p = t_of(lambda: current_density(b), reps)
    t_c = t_of(lambda: cathode_coupling(b), reps)
    with np.errstate(all='ignore'):
        r_p = t_of(lambda: current_density_oracle(b, 1.0, 91), max(3, reps // 20))
        r_c = t_of(lambda: cathode_coupling_oracle(b), max(3, reps // 20))
    print(f'n={n:8d} A=91  current_density {t_p * 1e6:10.1f} us (oracle {r_p * 1e6:12.1f} us, x{r_p / t_p:7.1f})   '
          f'cathode_coupling {t_c * 1e6:9.1f} us (oracle {r_c * 1e6:10.1f} us, x{r_c / t_c:6.1f})', flush=True)
