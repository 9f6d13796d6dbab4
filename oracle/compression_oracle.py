"""NumPy restatement of amisc's SVD field compression as the reference applies it to j_ion (TEST INFRASTRUCTURE).

The reference declares `j_ion` with `norm: log10` and `compression: {method: svd, reconstruction_tol: 0.01}`
(/root/reference/scripts/pem_v0/pem_v0_SPT-100.yml:272-280) and builds the map in scripts/gen_data.py:279-290 via
`var.normalize(...)` + `var.compression.compute_map(...)`.  The implementation lives in `amisc` (v0.8.1,
archermarx/amisc@ad5d48af, uv.lock:14-16), which is NOT present under /root/reference and not installed here:
this file restates its published algorithm (`amisc.compression.SVD`), so "parity unpinned" applies to it -- the device
kernels are checked against THIS restatement, and against size-independent properties (orthonormal projection,
compress(reconstruct(z)) == z, reconstruction error <= the tolerance on the compression set).

    normalize        x = log10(j)                                     amisc Variable.normalize, norm 'log10'
    compute_map      U, s, Vt = svd(X);  rank = first r with ||U_r U_r^T X - X||_F / ||X||_F <= reconstruction_tol
    compress         z = U_r^T x
    reconstruct      x_hat = U_r z ;  denormalize  j = 10 ** x_hat
"""
from __future__ import annotations

import numpy as np


def relative_error(pred, targ):
    return float(np.sqrt(np.sum((pred - targ) ** 2) / np.sum(targ ** 2)))


def compute_map_oracle(data_matrix: np.ndarray, rank=None, energy_tol=None, reconstruction_tol=None):
    """data_matrix: (dof, num_samples), already normalised.  Returns (projection (dof, rank), rank, energy, recon_err)."""
    dm = np.asarray(data_matrix, dtype=np.float64)
    dm = dm[:, ~np.any(np.isnan(dm), axis=0)]
    u, s, _ = np.linalg.svd(dm, full_matrices=False)
    energy_frac = np.cumsum(s ** 2 / np.sum(s ** 2))
    if rank:
        pass
    elif reconstruction_tol:
        rank = u.shape[1]
        for r in range(1, u.shape[1] + 1):
            if relative_error(u[:, :r] @ u[:, :r].T @ dm, dm) <= reconstruction_tol:
                rank = r
                break
    else:
        energy_tol = energy_tol or 0.95
        rank = int(np.where(energy_frac >= energy_tol)[0][0]) + 1
    proj = u[:, :rank]
    return proj, rank, float(energy_frac[rank - 1]), relative_error(proj @ proj.T @ dm, dm)


def compress_oracle(projection: np.ndarray, data: np.ndarray) -> np.ndarray:
    """(..., dof) normalised data -> (..., rank):  squeeze(P^T @ data[..., None])."""
    return np.squeeze(projection.T @ np.asarray(data)[..., np.newaxis], axis=-1)


def reconstruct_oracle(projection: np.ndarray, latent: np.ndarray) -> np.ndarray:
    return np.squeeze(projection @ np.asarray(latent)[..., np.newaxis], axis=-1)


def normalize_log10(j):
    with np.errstate(all='ignore'):
        return np.log10(j)


def denormalize_log10(x):
    with np.errstate(all='ignore'):
        return 10.0 ** x
