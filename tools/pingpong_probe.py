"""Developer probe: e2e latency when the previous result is still referenced during the next call (out = f(x) loops)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hallthrusterpem_b200.models import plume_cathode
from hallthrusterpem_b200.synthetic import spt100_batch
for n, A in ((100_000, 91), (1_000_000, 91), (1_000_000, 200)):
    b = spt100_batch(n, 1)
    out = plume_cathode(b, 1.0, n_angles=A)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        out = plume_cathode(b, 1.0, n_angles=A)          # previous `out` alive until this returns
        ts.append(time.perf_counter() - t0)
    ts2 = []
    for _ in range(8):
        del out
        t0 = time.perf_counter()
        out = plume_cathode(b, 1.0, n_angles=A)
        ts2.append(time.perf_counter() - t0)
    print(f'n={n} A={A}: rebinding loop {np.median(ts)*1e3:.2f} ms (first {ts[0]*1e3:.2f})  del-first loop {np.median(ts2)*1e3:.2f} ms  '
          f'-> {n*A/np.median(ts)/1e9:.2f} Geval/s', flush=True)
