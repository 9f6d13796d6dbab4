"""NumPy statement of the on-device sampler's stream (TEST INFRASTRUCTURE ONLY; never imported by the product package).

Philox4x32-10 (Salmon et al., SC'11), counter = (sample index lo, hi, input triple, 0), key = seed; three 42-bit uniforms
per call (word j plus ten bits of word 3); Uniform / LogUniform / const transforms as in csrc/hpem_sampler.cuh.  The Random123 known-answer vectors pin
the generator in tests/test_host_cpu.py; tests/test_sampler_gpu.py compares the device draws with this statement."""
from __future__ import annotations

import numpy as np

INPUT_NAMES = ('P_b', 'V_a', 'T_e', 'V_vac', 'Pstar', 'P_T',
               'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0', 'T')   # enum hpem_input order (include/hpem.h)


# ------------------------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 on uint32 arrays (uint64 intermediates)."""
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    w0, w1 = 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0, p1 = m0 * c0, m1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + w0) & 0xFFFFFFFF, (k1 + w1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_uniforms(seed: int, first_index: int, n: int) -> np.ndarray:
    """(n, 15) uniforms in [0, 1): column k is the uniform behind input k.  Call t of a sample serves inputs 3t..3t+2;
    uniform j of a call is k_j 2^-42 with k_j = (bits 10j..10j+9 of word 3) << 32 | word j."""
    idx = np.uint64(first_index) + np.arange(n, dtype=np.uint64)
    lo, hi = idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32)
    out = np.empty((n, 15))
    for t in range(5):
        o = philox4x32_10(lo, hi, np.full(n, t, np.uint64), np.zeros(n, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for j in range(3):
            k = (((o[3] >> np.uint64(10 * j)) & np.uint64(0x3FF)) << np.uint64(32)) | o[j]
            out[:, 3 * t + j] = k.astype(np.float64) * 2.0 ** -42
    return out


def apply_priors_numpy(u: np.ndarray, priors: dict) -> dict:
    """Uniform / LogUniform / const transforms of `philox_uniforms` columns (Normal needs the second stream; not restated)."""
    out = {}
    for k, name in enumerate(INPUT_NAMES):
        if name not in priors:
            continue
        kind, a, b = priors[name]
        if kind == 'uniform':
            out[name] = u[:, k] * (b - a) + a
        elif kind == 'loguniform':
            out[name] = np.exp(u[:, k] * (np.log(b) - np.log(a)) + np.log(a))
        elif kind == 'const':
            out[name] = np.full(u.shape[0], float(a))
        else:
            raise NotImplementedError(kind)
    return out
