"""Seeded synthetic input batches for the plume+cathode hot path (host side, NumPy).

The distributions are the priors/domains of the reference's PEM v0 system
(/root/reference/scripts/pem_v0/pem_v0_SPT-100.yml:9-56,219-271) and the ranges its own unit tests draw
from (/root/reference/tests/test_plume.py:19-29, tests/test_cathode.py:19-21).  HallThruster.jl stays
external, so its outputs `I_B0` (pem_to_julia.json:28) and `T` (:30) are synthetic arrays here.

The reference contains no H9 configuration; the "H9" generator is synthetic and flagged as such
(SURVEY.md section 8d): same plume priors, linear background-pressure sweep 0..1e-4 Torr.
"""
from __future__ import annotations

import numpy as np

CATHODE_KEYS = ('P_b', 'V_a', 'T_e', 'V_vac', 'Pstar', 'P_T')
PLUME_KEYS = ('P_b', 'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0')   # + optional 'T'
ALL_KEYS = ('P_b', 'V_a', 'T_e', 'V_vac', 'Pstar', 'P_T',
            'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0', 'T')
BASE_SEED = 20240307


def spt100_batch(n: int, seed: int = BASE_SEED, with_thrust: bool = True, c3_test_range: bool = False) -> dict:
    """SPT-100 Monte-Carlo batch: all 15 named inputs as float64 arrays of shape (n,)."""
    g = np.random.default_rng(seed)
    u = lambda lo, hi: g.uniform(lo, hi, n)  # noqa: E731
    d = {
        'P_b': 10.0 ** u(-8.0, -4.0),            # yml:15 domain (1e-8, 1e-4) Torr, log-uniform as in test_plume.py:20
        'V_a': u(200.0, 400.0),                  # yml:24
        'T_e': u(1.0, 5.0),                      # yml:31
        'V_vac': u(0.0, 60.0),                   # yml:38
        'Pstar': u(10e-6, 100e-6),               # yml:45
        'P_T': u(10e-6, 100e-6),                 # yml:53
        'c0': u(0.1, 0.9),                       # test_plume.py:21 (yml prior is U(0,1))
        'c1': u(0.1, 0.9),                       # yml:232
        'c2': u(-15.0, 15.0),                    # yml:239
        'c3': u(0.1, 1.1) if c3_test_range else u(0.2, 1.570796),   # test_plume.py:24 / yml:246
        'c4': 10.0 ** u(18.0, 22.0),             # yml:253
        'c5': 10.0 ** u(14.0, 18.0),             # yml:261
        'sigma_cex': u(51e-20, 58e-20),          # yml:269
        'I_B0': u(2.0, 8.0),                     # test_plume.py:28
    }
    if with_thrust:
        d['T'] = u(0.02, 0.12)                   # yml:183-184 (domain (0, 0.2) N, nominal 0.08)
    return d


def h9_sweep_batch(n: int, seed: int = BASE_SEED + 3, with_thrust: bool = True) -> dict:
    """Synthetic "H9" batch: linear pressure sweep 0..1e-4 Torr (includes P_b = 0), higher current/thrust."""
    d = spt100_batch(n, seed, with_thrust)
    g = np.random.default_rng(seed + 1000)
    d['P_b'] = np.linspace(0.0, 1e-4, n)
    d['V_a'] = g.uniform(300.0, 600.0, n)
    d['I_B0'] = g.uniform(10.0, 20.0, n)
    if with_thrust:
        d['T'] = g.uniform(0.2, 0.4, n)
    return d


def shard_bounds(n_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous sample range [lo, hi) owned by `rank`, balanced to within one sample -- or, for
    shards of 1024 samples and more, to within 64: interior boundaries are then multiples of 64 samples, so a sample sits
    at the same position modulo 4 in its shard as in the unsharded batch, which keeps the materialised outputs
    bit-identical to a single-GPU run (the quad-row kernel's rounding depends on that position)."""
    align = 64 if n_total // max(world_size, 1) >= 1024 else 1

    def edge(k: int) -> int:
        if k <= 0:
            return 0
        if k >= world_size:
            return n_total
        base, rem = divmod(n_total, world_size)
        return (k * base + min(k, rem)) // align * align
    return edge(rank), edge(rank + 1)
