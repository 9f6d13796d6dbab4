"""SVD compression of the `j_ion` field quantity -- the data format immediately downstream of the plume model.

In the reference's surrogate pipeline `j_ion` is declared with `norm: log10` and `compression: {method: svd,
reconstruction_tol: 0.01}` (/root/reference/scripts/pem_v0/pem_v0_SPT-100.yml:272-280); `scripts/gen_data.py:279-290`
builds the map from ~500 compression samples (`--compression-samples`, gen_data.py:73) with
`var.compression.compute_map(var.normalize(...))`, and from then on every model output is stored and trained on as
`rank` latent coefficients per sample instead of the `(A,)` field.

`amisc` (v0.8.1, archermarx/amisc@ad5d48af, uv.lock:14-16) is NOT vendored with the reference, so the class below
mirrors the published interface of `amisc.compression.SVD` (`compute_map`, `compress`, `reconstruct`, `latent_size`,
`estimate_latent_ranges`, attributes `projection_matrix`, `rank`, `energy_tol`, `reconstruction_tol`) and restates its
arithmetic; parity is pinned against the NumPy restatement in `oracle/compression_oracle.py`, not against amisc itself.

What runs where:
* `compute_map`   -- host, once: `np.linalg.svd` of the (dof, ~500) normalised compression matrix, rank selection.
* `compress_inputs` -- K4 `latent_kernel`: plume model + log10 + projection fused, `j_ion` never materialised.
* `compress` / `compress_field` -- K4f: projection of a materialised (normalised / raw) field.
* `reconstruct` / `reconstruct_field` -- K5: latent -> (normalised / raw) field.
All device entry points take and return torch CUDA tensors, or NumPy arrays (copied through the device).
There is no CPU fallback for them.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .engine import _Batch, _is_torch_tensor, get_grid, torr_2_pa

MAX_RANK = 32


def relative_error(pred: np.ndarray, targ: np.ndarray) -> float:
    """amisc.utils.relative_error: sqrt(sum (pred - targ)^2 / sum targ^2)."""
    with np.errstate(divide='ignore', invalid='ignore'):
        return float(np.sqrt(np.sum((pred - targ) ** 2) / np.sum(targ ** 2)))


class SVD:
    """Truncated-SVD compression map of one field quantity (mirror of `amisc.compression.SVD`).

    :ivar projection_matrix: `(dof, rank)` leading left-singular vectors of the data matrix
    :ivar rank: number of latent coefficients
    :ivar energy_tol: fraction of the squared singular values captured by `rank`
    :ivar reconstruction_tol: relative Frobenius reconstruction error of the data matrix at `rank`
    :ivar coords: the grid the field lives on (the `j_ion_coords` of the model output, gen_data.py:281-282)
    :ivar norm: `'log10'` (pem_v0_SPT-100.yml:277) or `None`; only used by the `*_field` / `*_inputs` methods, which
                fuse the normalisation into the kernel.  `compress` / `reconstruct` work on normalised data like amisc's.
    """

    def __init__(self, data_matrix=None, *, fields=('j_ion',), coords=None, rank: int | None = None,
                 energy_tol: float | None = None, reconstruction_tol: float | None = None, norm: str | None = 'log10',
                 device: int | None = None):
        if norm not in ('log10', None):
            raise ValueError(f"norm must be 'log10' or None, got {norm!r}")
        self.fields = list(fields)
        self.coords = coords
        self.rank = rank
        self.energy_tol = energy_tol
        self.reconstruction_tol = reconstruction_tol
        self.norm = norm
        self.device = device
        self.data_matrix = None
        self.projection_matrix = None
        self._handles: dict[tuple[int, int], ctypes.c_void_p] = {}
        if data_matrix is not None:
            self.compute_map(data_matrix, rank=rank, energy_tol=energy_tol, reconstruction_tol=reconstruction_tol)

    # ------------------------------------------------------------------------------------------------------------
    # host: the compression map (runs once on ~500 samples)
    # ------------------------------------------------------------------------------------------------------------
    def compute_map(self, data_matrix, rank: int | None = None, energy_tol: float | None = None,
                    reconstruction_tol: float | None = None) -> None:
        """Compute the projection from a NORMALISED data matrix `(dof, num_samples)` (or a dict `{field: (num_samples,
        dof)}` as gen_data.py:288-290 passes).  Rank priority as in amisc: `rank`, else the smallest rank whose relative
        reconstruction error is <= `reconstruction_tol`, else the smallest rank whose energy fraction is >= `energy_tol`
        (default 0.95).  Samples containing NaN are dropped."""
        if isinstance(data_matrix, dict):
            cols = [np.asarray(data_matrix[f], dtype=np.float64)[..., np.newaxis] for f in self.fields]
            dm = np.concatenate(cols, axis=-1)
            dm = dm.reshape(*dm.shape[:-2], -1).T
        else:
            dm = np.asarray(data_matrix, dtype=np.float64)
        if dm.ndim != 2:
            raise ValueError(f'data matrix must be (dof, num_samples), got shape {dm.shape}')
        dm = dm[:, ~np.any(np.isnan(dm), axis=0)]
        if dm.shape[1] == 0:
            raise ValueError('every compression sample contains NaN')
        u, s, _ = np.linalg.svd(dm, full_matrices=False)
        energy_frac = np.cumsum(s ** 2 / np.sum(s ** 2))
        rank = rank or self.rank
        reconstruction_tol = reconstruction_tol or self.reconstruction_tol
        if rank:
            rank = int(rank)
            reconstruction_tol = relative_error(u[:, :rank] @ (u[:, :rank].T @ dm), dm)
        elif reconstruction_tol:
            target, rank = reconstruction_tol, u.shape[1]
            for r in range(1, u.shape[1] + 1):
                reconstruction_tol = relative_error(u[:, :r] @ (u[:, :r].T @ dm), dm)
                if reconstruction_tol <= target:
                    rank = r
                    break
        else:
            energy_tol = energy_tol or self.energy_tol or 0.95
            rank = int(np.where(energy_frac >= energy_tol)[0][0]) + 1
            reconstruction_tol = relative_error(u[:, :rank] @ (u[:, :rank].T @ dm), dm)
        if rank > MAX_RANK:
            raise ValueError(f'rank {rank} exceeds the {MAX_RANK} latent coefficients the device kernels hold in registers')
        self.data_matrix = dm
        self.projection_matrix = np.ascontiguousarray(u[:, :rank])
        self.rank = rank
        self.energy_tol = float(energy_frac[rank - 1])
        self.reconstruction_tol = float(reconstruction_tol)
        self._release()

    @classmethod
    def from_samples(cls, inputs: dict, *, n_angles: int = 91, sweep_radius: float = 1.0, torr: float | None = None,
                     device: int | None = None, **kwargs) -> 'SVD':
        """gen_data.py:238-290 in one call: evaluate the plume model on the compression samples (on the device),
        normalise, and compute the map."""
        from .models import current_density
        out = current_density(inputs, sweep_radius, n_angles=n_angles, torr_2_pa=torr, device=device)
        j = out['j_ion']
        j = j.detach().cpu().numpy() if _is_torch_tensor(j) else np.asarray(j)
        norm = kwargs.get('norm', 'log10')
        with np.errstate(all='ignore'):
            x = np.log10(j) if norm == 'log10' else j
        comp = cls(coords=out['j_ion_coords'].reshape(-1)[0], device=device, **kwargs)
        comp.compute_map(x.reshape(-1, x.shape[-1]).T)
        return comp

    def latent_size(self) -> int:
        return int(self.rank)

    def estimate_latent_ranges(self) -> list[tuple[float, float]]:
        """(min, max) of every latent coefficient over the compression samples."""
        z = self.projection_matrix.T @ self.data_matrix
        return [(float(lo), float(hi)) for lo, hi in zip(z.min(axis=1), z.max(axis=1))]

    @property
    def dof(self) -> int:
        return int(self.projection_matrix.shape[0])

    # ------------------------------------------------------------------------------------------------------------
    # device side
    # ------------------------------------------------------------------------------------------------------------
    def _release(self):
        if getattr(self, '_handles', None):
            lib = _lib.load()
            for h in self._handles.values():
                lib.hpem_basis_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _basis(self, device: int, fused_norm: bool) -> ctypes.c_void_p:
        if self.projection_matrix is None:
            raise RuntimeError('compute_map() has not been called')
        key = (int(device), int(fused_norm))
        h = self._handles.get(key)
        if h is None:
            lib = _lib.load()
            h = ctypes.c_void_p()
            proj = np.ascontiguousarray(self.projection_matrix, dtype=np.float64)
            _lib.check(lib.hpem_basis_create(int(device), self.dof, int(self.rank),
                                             proj.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                             1 if (fused_norm and self.norm == 'log10') else 0, ctypes.byref(h)))
            self._handles[key] = h
        return h

    def _device_of(self, x) -> int:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('hallthrusterpem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        if _is_torch_tensor(x) and x.is_cuda:
            return x.device.index
        return torch.cuda.current_device() if self.device is None else int(self.device)

    def _project(self, data, fused_norm: bool):
        import torch
        dev = self._device_of(data)
        was_torch = _is_torch_tensor(data)
        t = data if was_torch else torch.as_tensor(np.asarray(data, dtype=np.float64))
        t = t.to(device=f'cuda:{dev}', dtype=torch.float64).contiguous()
        if t.shape[-1] != self.dof:
            raise ValueError(f'last axis of the data must be the dof ({self.dof}), got shape {tuple(t.shape)}')
        loop = tuple(t.shape[:-1])
        n = int(np.prod(loop, dtype=np.int64)) if loop else 1
        z = torch.empty(loop + (self.rank,), dtype=torch.float64, device=t.device)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().hpem_compress_field(self._basis(dev, fused_norm), n, ctypes.c_void_p(t.data_ptr()),
                                                   ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(st)))
        return z if was_torch else z.cpu().numpy()

    def _expand(self, latent, fused_norm: bool):
        import torch
        dev = self._device_of(latent)
        was_torch = _is_torch_tensor(latent)
        t = latent if was_torch else torch.as_tensor(np.asarray(latent, dtype=np.float64))
        t = t.to(device=f'cuda:{dev}', dtype=torch.float64).contiguous()
        if t.shape[-1] != self.rank:
            raise ValueError(f'last axis of the latent array must be the rank ({self.rank}), got shape {tuple(t.shape)}')
        loop = tuple(t.shape[:-1])
        n = int(np.prod(loop, dtype=np.int64)) if loop else 1
        f = torch.empty(loop + (self.dof,), dtype=torch.float64, device=t.device)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().hpem_reconstruct(self._basis(dev, fused_norm), n, ctypes.c_void_p(t.data_ptr()),
                                                ctypes.c_void_p(f.data_ptr()), ctypes.c_void_p(st)))
        return f if was_torch else f.cpu().numpy()

    def compress(self, data):
        """`(..., dof)` NORMALISED data -> `(..., rank)` latent coefficients (amisc `SVD.compress`)."""
        return self._project(data, fused_norm=False)

    def reconstruct(self, compressed):
        """`(..., rank)` latent coefficients -> `(..., dof)` NORMALISED data (amisc `SVD.reconstruct`)."""
        return self._expand(compressed, fused_norm=False)

    def compress_field(self, field):
        """`(..., dof)` raw field (e.g. `j_ion`) -> latent: `Variable.normalize` + `compress` in one kernel."""
        return self._project(field, fused_norm=True)

    def reconstruct_field(self, compressed):
        """latent -> raw field: `reconstruct` + `Variable.denormalize` (10**x) in one kernel."""
        return self._expand(compressed, fused_norm=True)

    def compress_inputs(self, inputs: dict, *, sweep_radius: float = 1.0, torr: float | None = None):
        """Plume model + normalisation + projection fused (K4): the latent coefficients of `current_density(inputs)
        ['j_ion']` without ever materialising the field.  Inputs: the plume input names; torch CUDA float64 tensors
        (output stays on the device) or NumPy arrays / scalars."""
        import torch
        batch = _Batch(inputs, _lib.PLUME_INPUTS)
        if not batch.on_device:
            if not torch.cuda.is_available():
                raise RuntimeError('hallthrusterpem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
            from .engine import device_to_host, host_to_device
            dev = torch.cuda.current_device() if self.device is None else int(self.device)
            shape = batch.out_shape
            moved = host_to_device({k: (np.ascontiguousarray(np.broadcast_to(np.asarray(inputs[k], dtype=np.float64), shape))
                                        if np.ndim(inputs[k]) > 0 else inputs[k]) for k in _lib.PLUME_INPUTS}, dev)
            if not any(hasattr(v, 'data_ptr') for v in moved.values()):          # all scalars: one sample
                moved['P_b'] = torch.full(shape, float(moved['P_b']), dtype=torch.float64, device=f'cuda:{dev}')
            with torch.cuda.device(dev):
                return device_to_host(self.compress_inputs(moved, sweep_radius=sweep_radius, torr=torr))
        dev = batch.device_index
        grid = get_grid(dev, self.dof, np.atleast_1d(np.float64(sweep_radius)))
        z = torch.empty(batch.out_shape + (self.rank,), dtype=torch.float64, device=f'cuda:{dev}')
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().hpem_compress(grid.handle, self._basis(dev, True), batch.n, ctypes.byref(batch.struct),
                                             torr_2_pa() if torr is None else float(torr), ctypes.c_void_p(z.data_ptr()),
                                             ctypes.c_void_p(st)))
        return z
