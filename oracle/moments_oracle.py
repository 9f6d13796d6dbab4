"""NumPy reference for the packed moments vector of the reduce-only pass (TEST INFRASTRUCTURE).

Given outputs of the plume/cathode functions for a batch (from the oracle, or from the materialising CUDA path) this
builds exactly what `hpem_moments_accumulate` is specified to produce (include/hpem.h), so tests can compare."""
from __future__ import annotations

import numpy as np


def hist_bins(j: np.ndarray, sub_bits: int, min_exp2: int, max_exp2: int) -> np.ndarray:
    """Log-linear bin index from the leading bits of the float64 pattern (see hpem.h)."""
    n_bins = ((max_exp2 - min_exp2) << sub_bits) + 2
    hi = (np.ascontiguousarray(j, dtype=np.float64).view(np.int64) >> 32).astype(np.int64)
    hi = np.where(hi >= 2 ** 31, hi - 2 ** 32, hi)
    key = hi >> (20 - sub_bits)
    b = key - ((min_exp2 + 1023) << sub_bits) + 1
    b = np.clip(b, 0, n_bins - 1)
    return np.where(hi < 0, 0, b).astype(np.int64)


def packed_moments(layout, j_ion, v_cc, div_angle, t_c, invalid):
    """layout: hallthrusterpem_b200.mc.Layout.  j_ion (n, A) as returned by current_density (1e-20 rows for invalid)."""
    n, A = j_ion.shape
    sums = np.zeros(layout.n_sums)
    minmax = np.full(6, -np.inf)
    row_ok = np.all(np.isfinite(j_ion), axis=1)
    sums[0] = n
    sums[1] = int(np.sum(invalid))
    sums[2] = int(np.sum(~row_ok))
    for k, x in enumerate((v_cc, div_angle, t_c)):
        if x is None:
            continue
        ok = ~np.isnan(x)
        sums[3 + 3 * k] = ok.sum()
        sums[4 + 3 * k] = x[ok].sum()
        sums[5 + 3 * k] = ((x[ok] - x[ok].mean()) ** 2).sum() if ok.any() else 0.0     # M2: centred (include/hpem.h)
        if ok.any():
            minmax[2 * k] = -x[ok].min()
            minmax[2 * k + 1] = x[ok].max()
    jj = j_ion[row_ok]
    sums[layout.off_angle_sum:layout.off_angle_sum + A] = jj.sum(axis=0)
    if jj.shape[0]:
        sums[layout.off_angle_sumsq:layout.off_angle_sumsq + A] = ((jj - jj.mean(axis=0)) ** 2).sum(axis=0)
    sp = layout.spec
    if sp.angle_stride > 0:
        h = np.zeros((layout.n_hist_angles, layout.n_bins))
        for a, i in enumerate(layout.hist_angle_index):
            b = hist_bins(jj[:, i], sp.sub_bits, sp.min_exp2, sp.max_exp2)
            h[a] = np.bincount(b, minlength=layout.n_bins)
        sums[layout.off_hist:] = h.reshape(-1)
    return sums, minmax
