"""Ion-current-density likelihood: the step right after the plume model in the reference's calibration scripts.

`scripts/pem_v0/monte_carlo.py:265-270` mirrors the 0..90 deg sweep to (-90, 90) deg, interpolates it linearly
(`scipy.interpolate.interp1d`) to the Faraday-probe angles, and `scripts/pem_v0/mcmc.py:103` sums
`-0.5 * ((y - y_hat) / sigma)**2` over the probe points.  Here both happen inside one CUDA kernel (K3) that never
materialises `j_ion`: per sample only the log-likelihood (and optionally the interpolated predictions) leave the GPU.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .engine import _Batch, _is_torch_tensor as _is_torch, get_grid, torr_2_pa


class JionMeasurements:
    """Probe angles (rad, |theta| <= pi/2), measured j_ion and standard deviations, resident on one device."""

    def __init__(self, theta, y, sigma, n_angles: int = 91, sweep_radius: float = 1.0, device: int | None = None):
        import torch
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.grid = get_grid(self.device, n_angles, np.atleast_1d(np.float64(sweep_radius)))
        th, yy, sg = (np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1)) for v in (theta, y, sigma))
        if not (th.shape == yy.shape == sg.shape):
            raise ValueError('theta, y and sigma must have the same length')
        if np.any(np.abs(th) > np.pi / 2):
            raise ValueError('A value in x_new is outside the interpolation range.')     # what interp1d raises
        self.m = int(th.shape[0])
        dptr = ctypes.POINTER(ctypes.c_double)
        h = ctypes.c_void_p()
        _lib.check(self.lib.hpem_measurements_create(self.grid.handle, self.m, th.ctypes.data_as(dptr),
                                                     yy.ctypes.data_as(dptr), sg.ctypes.data_as(dptr), ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, '_h', None):
            self.lib.hpem_measurements_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def jion_log_likelihood(inputs: dict, meas: JionMeasurements, *, torr: float | None = None, return_pred: bool = False):
    """Gaussian log-likelihood of the probe data under the plume model for every sample of `inputs`
    (torch CUDA float64 tensors -> torch results on the device; NumPy arrays / scalars -> NumPy results).  Returns the
    log-likelihood of loop shape (and the interpolated predictions, loop shape + (m,), with `return_pred`)."""
    import torch
    batch = _Batch(inputs, _lib.PLUME_INPUTS)
    if not batch.on_device:        # NumPy / scalar inputs (an MCMC step): through the device, NumPy back
        moved = {k: (torch.as_tensor(np.asarray(inputs[k], dtype=np.float64), device=f'cuda:{meas.device}')
                     if np.ndim(inputs[k]) > 0 else inputs[k]) for k in _lib.PLUME_INPUTS}
        if not any(_is_torch(v) for v in moved.values()):      # all scalars: one sample
            moved['P_b'] = torch.as_tensor(np.atleast_1d(np.float64(inputs['P_b'])), device=f'cuda:{meas.device}')
        res = jion_log_likelihood(moved, meas, torr=torr, return_pred=return_pred)
        return tuple(r.cpu().numpy() for r in res) if return_pred else res.cpu().numpy()
    if batch.device_index != meas.device:
        raise ValueError('jion_log_likelihood expects inputs on the device of the measurement set')
    dev = f'cuda:{meas.device}'
    ll = torch.empty(batch.out_shape, dtype=torch.float64, device=dev)
    pred = torch.empty(batch.out_shape + (meas.m,), dtype=torch.float64, device=dev) if return_pred else None
    stream = torch.cuda.current_stream(meas.device).cuda_stream
    _lib.check(meas.lib.hpem_loglike(meas.grid.handle, meas._h, batch.n, ctypes.byref(batch.struct),
                                     torr_2_pa() if torr is None else float(torr), ctypes.c_void_p(ll.data_ptr()),
                                     ctypes.c_void_p(pred.data_ptr()) if return_pred else None, ctypes.c_void_p(stream)))
    return (ll, pred) if return_pred else ll


def marginal_log_likelihood(loglike):
    """Log-sum-exp of per-draw log-likelihoods over the trailing axis (the M Monte-Carlo draws of the nuisance parameters
    behind one calibration vector): `max + log(sum(exp(ll - max)))`, scripts/pem_v0/mcmc.py:101-102.  torch CUDA tensor
    `(..., M)` -> `(...,)` on the device; NumPy in -> NumPy out (through the device)."""
    import torch
    was_torch = _is_torch(loglike)
    t = loglike if was_torch else torch.as_tensor(np.asarray(loglike, dtype=np.float64))
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError('hallthrusterpem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        t = t.cuda()
    t = t.to(torch.float64).contiguous()
    if t.dim() < 1 or t.shape[-1] < 1:
        raise ValueError('need a trailing axis of at least one draw')
    out = torch.empty(t.shape[:-1], dtype=torch.float64, device=t.device)
    n_groups = int(np.prod(t.shape[:-1], dtype=np.int64)) if t.dim() > 1 else 1
    dev = t.device.index
    _lib.check(_lib.load().hpem_logsumexp(dev, n_groups, int(t.shape[-1]), ctypes.c_void_p(t.data_ptr()),
                                          ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out if was_torch else out.cpu().numpy()
