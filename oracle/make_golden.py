#!/usr/bin/env python
"""Mint tests/golden/*.npz (TEST INFRASTRUCTURE).  Run in the build container, where /root/reference exists:

    python -m oracle.make_golden

For every case the inputs are seeded (np.random.default_rng), the outputs come from

* the UNMODIFIED reference (`oracle.ref_import`) when n_angles == 91, after asserting that the restatement
  (`oracle.ref_restated`) reproduces it bit for bit, and
* the restatement for n_angles != 91 (the reference hard-codes 91 at plume.py:53).

The reference's own tests hold no golden vectors, so these files ARE the pinned vectors;
NumPy/SciPy versions and the CPU they were produced on are recorded in each file's `meta`.
"""
from __future__ import annotations

import json
import platform
import sys
from pathlib import Path

import numpy as np
import scipy

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from hallthrusterpem_b200.synthetic import BASE_SEED, h9_sweep_batch, spt100_batch  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.ref_restated import cathode_coupling_oracle, current_density_oracle  # noqa: E402

GOLDEN = ROOT / 'tests' / 'golden'


def edge_batch() -> dict:
    """Hand-built corner cases (one per row); see tests/test_parity_gpu.py::test_edge_cases for the intent."""
    rows = [
        # P_b     V_a   T_e  V_vac Pstar  P_T    c0    c1     c2     c3      c4    c5    sigma   I_B0  T
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.1,  0.7,   -8.0,  0.2,    1e20, 1e16, 55e-20, 3.0,  0.08),  # SURVEY app. B sanity point
        (1e-4,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   -15.0, 0.1,    1e20, 1e16, 55e-20, 3.0,  0.08),  # alpha1 < 0 -> invalid
        (1e-4,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   15.0,  1.5,    1e20, 1e16, 55e-20, 3.0,  0.08),  # alpha1 clipped to pi/2
        (0.0,     300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   5.0,   0.4,    1e20, 1e16, 55e-20, 3.0,  0.08),  # P_b = 0
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   0.0,   0.0,    1e20, 1e16, 55e-20, 3.0,  0.08),  # alpha1 == 0 -> NaN row, invalid
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.003, 0.0,   0.2,    1e20, 1e16, 55e-20, 3.0,  0.08),  # alpha2 = 66.7: erfi overflow -> NaN
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 1.5,  0.5,   0.0,   0.3,    1e20, 1e16, 55e-20, 3.0,  0.08),  # c0 > 1: negative main beam -> late invalid
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   0.0,   0.3,    1e20, 1e16, 55e-20, 0.0,  0.08),  # I_B0 = 0 -> j_ion == 0 -> invalid, 0/0
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.0,  0.5,   0.0,   0.3,    1e20, 1e16, 55e-20, 3.0,  0.08),  # c0 = 0: no scattered beam
        (np.nan,  300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   0.0,   0.3,    1e20, 1e16, 55e-20, 3.0,  0.08),  # NaN input
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   0.0,   0.004,  1e20, 1e16, 55e-20, 3.0,  0.08),  # needle beam: profile underflows
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   0.0,   0.3,    1e23, 1e16, 55e-20, 3.0,  0.08),  # opaque CEX: decay ~ 1e-32
        (1e-8,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   0.0,   0.3,    1e18, 1e14, 51e-20, 3.0,  0.08),  # thinnest CEX: 1-decay ~ 5e-5
        (1e-4,    200,  5,   60,   10e-6, 10e-6, 0.3,  0.5,   0.0,   0.3,    1e20, 1e16, 55e-20, 3.0,  0.08),  # cathode: large PB/PT
        (1e-4,    20,   5,   60,   100e-6, 10e-6, 0.3, 0.5,   0.0,   0.3,    1e20, 1e16, 55e-20, 3.0,  0.08),  # cathode: V_cc > V_a clamp
        (1e-4,    300,  5,   0.0,  10e-6, 100e-6, 0.3, 0.5,   0.0,   0.3,    1e20, 1e16, 55e-20, 3.0,  0.08),  # cathode: V_cc < 0 clamp
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.0295, 0.0,  1.5707963, 1e20, 1e16, 55e-20, 3.0, 0.08),  # alpha2 = 53.25: just inside erfi range
        (1e-5,    300,  3,   30,   20e-6, 50e-6, 0.3,  0.5,   0.0,   -0.3,   1e20, 1e16, 55e-20, 3.0,  0.08),  # alpha1 = -0.3: invalid but div_angle finite
    ]
    from hallthrusterpem_b200.synthetic import ALL_KEYS
    arr = np.array(rows, dtype=np.float64)
    return {k: np.ascontiguousarray(arr[:, i]) for i, k in enumerate(ALL_KEYS)}


CASES = {
    # name: (batch builder, n_angles, sweep_radius)
    'cfg1_spt100_n256_a100': (lambda: spt100_batch(256, BASE_SEED + 1), 100, 1.0),
    'ref91_spt100_n256': (lambda: spt100_batch(256, BASE_SEED + 11), 91, 1.0),
    'ref91_testrange_n64_r25': (lambda: spt100_batch(64, BASE_SEED + 12, c3_test_range=True), 91,
                                np.random.default_rng(BASE_SEED + 13).uniform(1.0, 1.2, 25)),
    'cfg2_spt100_n128_a200': (lambda: spt100_batch(128, BASE_SEED + 2), 200, 1.0),
    'cfg3_h9_n128_a256': (lambda: h9_sweep_batch(128, BASE_SEED + 3), 256, 1.0),
    'cfg5_spt100_n64_a512': (lambda: spt100_batch(64, BASE_SEED + 5), 512, 1.0),
    'edge_a91': (edge_batch, 91, 1.0),
    'edge_a100': (edge_batch, 100, 1.0),
    'edge_a91_r3': (edge_batch, 91, np.array([0.5, 1.0, 2.0])),
}


def main():
    if not ref_import.available():
        raise SystemExit('the reference is not available here; golden vectors can only be minted in the build container')
    ref_plume, ref_cathode, torr = ref_import.load()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    meta_common = {'numpy': np.__version__, 'scipy': scipy.__version__, 'machine': platform.processor() or platform.machine(),
                   'torr_2_pa': torr, 'generator': 'oracle/make_golden.py'}
    for name, (builder, n_angles, radius) in CASES.items():
        batch = builder()
        with np.errstate(all='ignore'):
            restated = current_density_oracle(batch, radius, n_angles, torr, with_coords=False, return_internals=True)
            v_restated = cathode_coupling_oracle(batch, torr)['V_cc']
            if n_angles == 91:
                ref = ref_plume(dict(batch), radius)
                for key in ('j_ion', 'div_angle', 'T_c'):
                    if not np.array_equal(ref[key], restated[key], equal_nan=True):
                        raise SystemExit(f'{name}: restatement differs from the reference in {key}')
                source = 'reference (unmodified) == restatement (bit-identical)'
            else:
                source = 'restatement (n_angles != 91)'
            if not np.array_equal(ref_cathode(dict(batch))['V_cc'], v_restated, equal_nan=True):
                raise SystemExit(f'{name}: cathode restatement differs from the reference')
        meta = dict(meta_common, case=name, n_angles=n_angles, source=source)
        np.savez(GOLDEN / f'{name}.npz', meta=json.dumps(meta), sweep_radius=np.atleast_1d(radius),
                 **{f'in_{k}': v for k, v in batch.items()},
                 V_cc=v_restated, j_ion=restated['j_ion'], div_angle=restated['div_angle'], T_c=restated['T_c'],
                 cos_div=restated['_cos_div'], invalid=restated['_invalid'])
        print(f'{name}: j_ion {restated["j_ion"].shape}, invalid {int(restated["_invalid"].sum())}, {source}')


if __name__ == '__main__':
    main()
