"""Ion-current-density likelihood: the step right after the plume model in the reference's calibration scripts.

`scripts/pem_v0/monte_carlo.py:265-270` mirrors the 0..90 deg sweep to (-90, 90) deg, interpolates it linearly
(`scipy.interpolate.interp1d`) to the Faraday-probe angles, and `scripts/pem_v0/mcmc.py:103` sums
`-0.5 * ((y - y_hat) / sigma)**2` over the probe points.  Here both happen inside one CUDA kernel (K3) that never
materialises `j_ion`: per sample only the log-likelihood (and optionally the interpolated predictions) leave the GPU.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _lib
from .engine import _Batch, _is_torch_tensor as _is_torch, get_grid, torr_2_pa


class JionMeasurements:
    """Probe angles (rad, |theta| <= pi/2), measured j_ion and standard deviations.  `device` is one CUDA device index or
    -- for host inputs evaluated from one process on several GPUs -- 'all' / a list; the sorted, pre-weighted probe table is
    created on each device the first time it is used there."""

    def __init__(self, theta, y, sigma, n_angles: int = 91, sweep_radius: float = 1.0, device=None):
        import torch
        self.lib = _lib.load()
        if device is None:
            self.devices = [torch.cuda.current_device()]
        else:
            from .engine import resolve_devices
            self.devices = resolve_devices(device)
        self.device = self.devices[0]
        self.n_angles, self.sweep_radius = int(n_angles), float(sweep_radius)
        th, yy, sg = (np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1)) for v in (theta, y, sigma))
        if not (th.shape == yy.shape == sg.shape):
            raise ValueError('theta, y and sigma must have the same length')
        if np.any(np.abs(th) > np.pi / 2):
            raise ValueError('A value in x_new is outside the interpolation range.')     # what interp1d raises
        self.m = int(th.shape[0])
        self._arrays = (th, yy, sg)
        self._handles: dict[int, tuple] = {}
        self._lock = threading.Lock()
        self.on(self.device)

    def on(self, dev: int):
        """(grid handle, measurement handle) on device `dev`."""
        with self._lock:
            got = self._handles.get(dev)
            if got is None:
                grid = get_grid(dev, self.n_angles, np.atleast_1d(np.float64(self.sweep_radius)))
                th, yy, sg = self._arrays
                dptr = ctypes.POINTER(ctypes.c_double)
                h = ctypes.c_void_p()
                _lib.check(self.lib.hpem_measurements_create(grid.handle, self.m, th.ctypes.data_as(dptr),
                                                             yy.ctypes.data_as(dptr), sg.ctypes.data_as(dptr), ctypes.byref(h)))
                got = self._handles[dev] = (grid, h)
            return got

    @property
    def grid(self):
        return self.on(self.device)[0]

    @property
    def _h(self):
        return self.on(self.device)[1]

    def close(self):
        for _, h in getattr(self, '_handles', {}).values():
            self.lib.hpem_measurements_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _loglike_on_device(batch: _Batch, meas: JionMeasurements, dev: int, torr: float | None, return_pred: bool):
    import torch
    grid, h = meas.on(dev)
    ll = torch.empty(batch.out_shape, dtype=torch.float64, device=f'cuda:{dev}')
    pred = torch.empty(batch.out_shape + (meas.m,), dtype=torch.float64, device=f'cuda:{dev}') if return_pred else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(meas.lib.hpem_loglike(grid.handle, h, batch.n, ctypes.byref(batch.struct),
                                     torr_2_pa() if torr is None else float(torr), ctypes.c_void_p(ll.data_ptr()),
                                     ctypes.c_void_p(pred.data_ptr()) if return_pred else None, ctypes.c_void_p(stream)))
    return (ll, pred) if return_pred else ll


def jion_log_likelihood(inputs: dict, meas: JionMeasurements, *, torr: float | None = None, return_pred: bool = False):
    """Gaussian log-likelihood of the probe data under the plume model for every sample of `inputs`
    (torch CUDA float64 tensors -> torch results on the device; NumPy arrays / scalars -> NumPy results).  Returns the
    log-likelihood of loop shape (and the interpolated predictions, loop shape + (m,), with `return_pred`).
    Host inputs with a measurement set created for several devices are sharded over them from this one process."""
    import torch
    batch = _Batch(inputs, _lib.PLUME_INPUTS)
    if batch.on_device:
        if batch.device_index not in meas.devices:
            raise ValueError('jion_log_likelihood expects inputs on a device of the measurement set')
        return _loglike_on_device(batch, meas, batch.device_index, torr, return_pred)

    # NumPy / scalar inputs (an MCMC step): through the device(s), NumPy back
    host = {k: np.asarray(inputs[k], dtype=np.float64) for k in _lib.PLUME_INPUTS}
    n, shape = batch.n, batch.out_shape
    flat = {k: (np.ascontiguousarray(np.broadcast_to(v, shape)).reshape(-1) if v.size > 1 else v.reshape(-1)) for k, v in host.items()}
    ll = np.empty(n)
    pred = np.empty((n, meas.m)) if return_pred else None
    from .engine import _thread_pool, device_to_host, host_to_device
    from .synthetic import shard_bounds
    devs = meas.devices if n >= 64 * len(meas.devices) else meas.devices[:1]

    def one(r):
        lo, hi = shard_bounds(n, len(devs), r)
        if hi <= lo:
            return
        d = devs[r]
        with torch.cuda.device(d):
            part = host_to_device({k: (v[lo:hi] if v.size > 1 else v) for k, v in flat.items()}, d)
            if not any(_is_torch(v) for v in part.values()):      # all scalars: one sample
                part['P_b'] = torch.full((hi - lo,), float(flat['P_b'][0]), dtype=torch.float64, device=f'cuda:{d}')
            res = _loglike_on_device(_Batch(part, _lib.PLUME_INPUTS), meas, d, torr, return_pred)
            if return_pred:
                ll[lo:hi] = device_to_host(res[0])
                pred[lo:hi] = device_to_host(res[1])
            else:
                ll[lo:hi] = device_to_host(res)

    if len(devs) == 1:
        one(0)
    else:
        list(_thread_pool().map(one, range(len(devs))))
    ll = ll.reshape(shape)
    return (ll, pred.reshape(shape + (meas.m,))) if return_pred else ll


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
