"""Host-side constants of the plume path: the angle grid and the fused Simpson weights.

The reference integrates the flipped, cos/sin-weighted beam profile with
`scipy.integrate.simpson(y, x=alpha_rad, axis=-2)` (/root/reference/src/hallmd/models/plume.py:117-123).
Because `x` is passed SciPy uses the composite Simpson rule for irregular spacing and, for an even number of
points, Cartwright's correction on the last interval.  Both are linear in `y`, so

    num = sum_k W[k] * cos(a[k]) * sin(a[k]) * f[A-1-k],     den = sum_k W[k] * cos(a[k]) * f[A-1-k]

with f = j_beam + j_scat in natural (un-flipped) order.  Re-indexing i = A-1-k folds the `np.flip` into the
weights:  wd[i] = W[A-1-i]*cos(a[A-1-i]),  wn[i] = wd[i]*sin(a[A-1-i]).  These 2*A numbers are computed once per
grid on the host (here) and handed to the library (`hpem_grid_create`); the device does two dot products.
"""
from __future__ import annotations

import numpy as np


def angle_grid(n_angles: int = 91) -> np.ndarray:
    """The reference's sweep: `np.linspace(0, np.pi / 2, 91)` (plume.py:53), angle count as a parameter."""
    if n_angles < 2:
        raise ValueError('n_angles must be >= 2')
    return np.linspace(0, np.pi / 2, n_angles)


def simpson_weights(x: np.ndarray) -> np.ndarray:
    """Weights W with `scipy.integrate.simpson(y, x=x) == W @ y` (composite Simpson for irregular spacing;
    even point counts get Cartwright's last-interval correction, as SciPy >= 1.11 does)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    w = np.zeros(n)
    if n == 2:  # SciPy falls back to the trapezoid rule
        w[:] = 0.5 * (x[1] - x[0])
        return w
    m = n if n % 2 == 1 else n - 1          # points covered by whole Simpson panels
    h = np.diff(x)
    h0, h1 = h[0:m - 1:2], h[1:m - 1:2]     # left / right interval of each panel
    hsum, hprod, hdiv = h0 + h1, h0 * h1, h0 / h1
    c = hsum / 6.0
    w_left = c * (2.0 - 1.0 / hdiv)
    w_mid = c * (hsum * hsum / hprod)
    w_right = c * (2.0 - hdiv)
    np.add.at(w, np.arange(0, m - 1, 2), w_left)
    np.add.at(w, np.arange(1, m, 2), w_mid)
    np.add.at(w, np.arange(2, m + 1, 2), w_right)
    if n % 2 == 0:                           # Cartwright correction for the last interval
        g0, g1 = h[-2], h[-1]
        w[-1] += (2.0 * g1 * g1 + 3.0 * g0 * g1) / (6.0 * (g0 + g1))
        w[-2] += (g1 * g1 + 3.0 * g1 * g0) / (6.0 * g0)
        w[-3] -= g1 ** 3 / (6.0 * g0 * (g0 + g1))
    return w


def fused_weights(alpha: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(wd, wn) in un-flipped angle order; see module docstring.  Uses the reference's own expressions
    cos(alpha[k]), sin(alpha[k]) of the *flipped partner* (cos(alpha[A-1-i]) is not bit-equal to sin(alpha[i]))."""
    w = simpson_weights(alpha)
    wd_flipped = w * np.cos(alpha)           # multiplies f[A-1-k]
    wn_flipped = wd_flipped * np.sin(alpha)
    return np.ascontiguousarray(wd_flipped[::-1]), np.ascontiguousarray(wn_flipped[::-1])
