"""Reduce-only Monte-Carlo over the fused cathode + plume chain, sharded over the GPUs of one box.

Configs 4-5 of BASELINE.json (1e8-1e9 samples x up to 512 angles) cannot materialise `j_ion` (512 GB per GPU), and their
consumers only want statistics over the sample axis (percentiles / moments: tests/test_plume.py:50-52,
scripts/gen_data.py:402-404).  `MonteCarloMoments` streams chunks of device-resident input samples through libhpem's
reduce-only kernel (K2) and keeps two small device buffers:

* `sums`   -- packed float64 vector (counts, sums, sums of squares, per-angle sums, log-linear histograms),
* `minmax` -- (-min, max) of the per-sample scalars.

Samples shard trivially across ranks (contiguous ranges, `synthetic.shard_bounds`); the ONLY collective of the whole
path is `merge()`: one all-reduce(SUM) of `sums` and one all-reduce(MAX) of `minmax` over torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  `MomentsResult` turns the packed vector into means,
variances and percentile estimates.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from .engine import _Batch, get_grid, torr_2_pa

N_SCALARS = 12
SCALAR_NAMES = ('V_cc', 'div_angle', 'T_c')


@dataclass(frozen=True)
class HistogramSpec:
    angle_stride: int = 8      # histogram every 8th angle (power of two; 0 disables histograms)
    sub_bits: int = 3          # 8 log-linear bins per octave (~9 % relative resolution)
    min_exp2: int = -30        # 2^-30 ~ 1e-9 A/m^2
    max_exp2: int = 14         # 2^14  ~ 1.6e4 A/m^2


class Layout:
    """Python view of hpem_moments_layout (+ bin edges)."""

    def __init__(self, n_angles: int, spec: HistogramSpec, c_layout: '_lib.HpemMomentsLayout | None' = None):
        self.n_angles = n_angles
        self.spec = spec
        st = spec.angle_stride
        self.n_hist_angles = (n_angles + st - 1) // st if st > 0 else 0
        self.n_bins = ((spec.max_exp2 - spec.min_exp2) << spec.sub_bits) + 2 if st > 0 else 0
        self.off_angle_sum = N_SCALARS
        self.off_angle_sumsq = self.off_angle_sum + n_angles
        self.off_hist = self.off_angle_sumsq + n_angles
        self.n_sums = self.off_hist + self.n_hist_angles * self.n_bins
        if c_layout is not None:   # the library is authoritative; this class must agree with it
            assert (c_layout.n_sums, c_layout.off_angle_sum, c_layout.off_angle_sumsq, c_layout.off_hist,
                    c_layout.n_hist_angles, c_layout.n_bins) == (self.n_sums, self.off_angle_sum, self.off_angle_sumsq,
                                                                 self.off_hist, self.n_hist_angles, self.n_bins)

    @property
    def hist_angle_index(self) -> np.ndarray:
        return np.arange(self.n_hist_angles) * max(self.spec.angle_stride, 1)

    def bin_edges(self) -> np.ndarray:
        """Upper edges of bins 0 .. n_bins-2 (bin 0 = underflow, last bin = overflow)."""
        sub = 1 << self.spec.sub_bits
        octaves = np.arange(self.spec.min_exp2, self.spec.max_exp2)
        edges = (2.0 ** octaves)[:, None] * (1.0 + np.arange(sub) / sub)[None, :]
        return np.concatenate([edges.reshape(-1), [2.0 ** self.spec.max_exp2]])


class MomentsResult:
    """Decoded statistics of a packed `sums` / `minmax` pair (host NumPy arrays)."""

    def __init__(self, layout: Layout, sums: np.ndarray, minmax: np.ndarray):
        self.layout = layout
        self.sums = np.asarray(sums, dtype=np.float64)
        self.minmax = np.asarray(minmax, dtype=np.float64)
        self.n_samples = int(self.sums[0])
        self.n_invalid = int(self.sums[1])
        self.n_nonfinite_rows = int(self.sums[2])

    def scalar(self, name: str) -> dict:
        k = SCALAR_NAMES.index(name)
        n, s1, s2 = self.sums[3 + 3 * k: 6 + 3 * k]
        mean = s1 / n if n > 0 else np.nan
        var = max(s2 / n - mean * mean, 0.0) if n > 0 else np.nan
        return {'n': int(n), 'mean': mean, 'var': var, 'min': -self.minmax[2 * k], 'max': self.minmax[2 * k + 1]}

    @property
    def j_mean(self) -> np.ndarray:
        n = self.n_samples - self.n_nonfinite_rows
        L = self.layout
        return self.sums[L.off_angle_sum:L.off_angle_sum + L.n_angles] / max(n, 1)

    @property
    def j_var(self) -> np.ndarray:
        n = self.n_samples - self.n_nonfinite_rows
        L = self.layout
        m2 = self.sums[L.off_angle_sumsq:L.off_angle_sumsq + L.n_angles] / max(n, 1)
        return np.maximum(m2 - self.j_mean ** 2, 0.0)

    @property
    def histograms(self) -> np.ndarray:
        L = self.layout
        return self.sums[L.off_hist:L.off_hist + L.n_hist_angles * L.n_bins].reshape(L.n_hist_angles, L.n_bins)

    def j_percentile(self, q) -> np.ndarray:
        """Percentile estimates (per histogrammed angle) of j_ion over the samples, to bin resolution: the upper edge
        of the first bin whose cumulative count reaches q %."""
        L = self.layout
        h = self.histograms
        cum = np.cumsum(h, axis=1)
        total = cum[:, -1:]
        edges = np.concatenate([L.bin_edges(), [np.inf]])
        q = np.atleast_1d(np.asarray(q, dtype=np.float64))
        out = np.empty((q.shape[0], L.n_hist_angles))
        for i, qq in enumerate(q):
            idx = np.argmax(cum >= qq / 100.0 * total, axis=1)
            out[i] = edges[idx]
        return out


class MonteCarloMoments:
    """Accumulates moments/histograms of the cathode+plume chain over device-resident sample chunks on one GPU."""

    def __init__(self, n_angles: int = 91, sweep_radius: float = 1.0, hist: HistogramSpec = HistogramSpec(),
                 device: int | None = None, torr: float | None = None, want_cathode: bool = True, want_thrust: bool = True):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.grid = get_grid(self.device, n_angles, np.atleast_1d(np.float64(sweep_radius)))
        self.torr = torr_2_pa() if torr is None else float(torr)
        self.spec = _lib.HpemMomentsSpec(hist.angle_stride, hist.sub_bits, hist.min_exp2, hist.max_exp2,
                                         int(want_cathode), int(want_thrust))
        c_layout = _lib.HpemMomentsLayout()
        _lib.check(self.lib.hpem_moments_layout_query(self.grid.handle, ctypes.byref(self.spec), ctypes.byref(c_layout)))
        self.layout = Layout(n_angles, hist, c_layout)
        self.want_cathode, self.want_thrust = want_cathode, want_thrust
        dev = f'cuda:{self.device}'
        self.sums = torch.zeros(self.layout.n_sums, dtype=torch.float64, device=dev)
        self.minmax = torch.full((6,), -np.inf, dtype=torch.float64, device=dev)

    def reset(self):
        self.sums.zero_()
        self.minmax.fill_(-np.inf)

    def accumulate(self, inputs: dict) -> int:
        """Add one chunk of samples (dict of torch CUDA float64 tensors / scalars); asynchronous on the current stream."""
        names = tuple(_lib.CATHODE_INPUTS) if self.want_cathode else ()
        names += tuple(k for k in _lib.PLUME_INPUTS if k not in names)
        batch = _Batch(inputs, names, optional=('T',) if self.want_thrust else ())
        if self.want_thrust and 'T' not in batch.present:
            raise KeyError('T')
        if not batch.on_device:
            raise ValueError('MonteCarloMoments.accumulate expects device-resident (torch CUDA) inputs')
        if batch.device_index != self.device:
            raise ValueError(f'inputs live on cuda:{batch.device_index}, the reducer on cuda:{self.device}')
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.hpem_moments_accumulate(self.grid.handle, batch.n, ctypes.byref(batch.struct), self.torr,
                                                    ctypes.byref(self.spec), ctypes.c_void_p(self.sums.data_ptr()),
                                                    ctypes.c_void_p(self.minmax.data_ptr()), ctypes.c_void_p(stream)))
        return batch.n

    def accumulate_sampled(self, n: int, seed: int, first_index: int = 0, priors: dict | None = None) -> int:
        """Add samples [first_index, first_index + n) of the global index space, drawn on the fly by the on-device
        sampler (no input arrays exist); asynchronous on the current stream."""
        from .sampler import SPT100_PRIORS, priors_struct
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.hpem_moments_accumulate_sampled(
            self.grid.handle, int(n), int(seed), int(first_index), priors_struct(priors or SPT100_PRIORS), self.torr,
            ctypes.byref(self.spec), ctypes.c_void_p(self.sums.data_ptr()), ctypes.c_void_p(self.minmax.data_ptr()),
            ctypes.c_void_p(stream)))
        return int(n)

    def merge(self, group=None) -> None:
        """The path's only collective: all-reduce the packed buffers over the ranks (no-op without a process group)."""
        merge_buffers(self.sums, self.minmax, group)

    def result(self) -> MomentsResult:
        return MomentsResult(self.layout, self.sums.cpu().numpy(), self.minmax.cpu().numpy())


def merge_buffers(sums, minmax, group=None) -> None:
    """all-reduce(SUM) of `sums`, all-reduce(MAX) of `minmax` (torch tensors, any backend); in place."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(minmax, op=dist.ReduceOp.MAX, group=group)
