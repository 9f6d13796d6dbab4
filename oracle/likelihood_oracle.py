"""NumPy/SciPy restatement of the reference's j_ion interpolation + Gaussian log-likelihood (TEST INFRASTRUCTURE).

Follows /root/reference/scripts/pem_v0/monte_carlo.py:265-270 (mirror the sweep, `interp1d`, evaluate at the probe
angles) and scripts/pem_v0/mcmc.py:103 (`np.sum(-0.5 * ((ye - y_curr) / std) ** 2)`).  Those scripts are stale in
the reference tree (they import modules that no longer exist), so this cannot be pinned by RUNNING them; it is pinned
to their text, with the plume profile coming from the pinned plume oracle."""
from __future__ import annotations

import numpy as np
from scipy.interpolate import interp1d

from .ref_restated import angle_grid, current_density_oracle


def jion_interp_oracle(j_ion: np.ndarray, theta: np.ndarray, n_angles: int) -> np.ndarray:
    alpha_g = angle_grid(n_angles)                                                        # monte_carlo.py:265
    alpha_g2 = np.concatenate((-np.flip(alpha_g)[:-1], alpha_g))                          # :267
    jion_g2 = np.concatenate((np.flip(j_ion, axis=-1)[..., :-1], j_ion), axis=-1)         # :268
    f = interp1d(alpha_g2, jion_g2, axis=-1)                                              # :269
    return f(theta)                                                                       # :270


def jion_log_likelihood_oracle(inputs: dict, theta, y, sigma, n_angles: int, torr_2_pa: float):
    with np.errstate(all='ignore'):
        j = current_density_oracle(inputs, 1.0, n_angles, torr_2_pa, with_coords=False)['j_ion']
    pred = jion_interp_oracle(j, np.asarray(theta, dtype=np.float64), n_angles)
    ll = np.sum(-0.5 * ((np.asarray(y) - pred) / np.asarray(sigma)) ** 2, axis=-1)       # mcmc.py:103
    return ll, pred


def marginal_log_likelihood_oracle(loglike: np.ndarray) -> np.ndarray:
    """scripts/pem_v0/mcmc.py:101-102: log-sum-exp over the M axis."""
    with np.errstate(all='ignore'):
        mx = np.max(loglike, axis=-1, keepdims=True)                                           # :101
        return np.squeeze(mx, axis=-1) + np.log(np.sum(np.exp(loglike - mx), axis=-1))        # :102
