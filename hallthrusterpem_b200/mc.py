"""Reduce-only Monte-Carlo over the fused cathode + plume chain, sharded over the GPUs of one box.

Configs 4-5 of BASELINE.json (1e8-1e9 samples x up to 512 angles) cannot materialise `j_ion` (512 GB per GPU), and their
consumers only want statistics over the sample axis (percentiles / moments: tests/test_plume.py:50-52,
scripts/gen_data.py:402-404).  `MonteCarloMoments` streams chunks of samples through libhpem's reduce-only kernel (K2)
and keeps ONE small device buffer, `packed` = [sums | minmax]:

* `sums`   -- float64 vector: counts, per-scalar (n, sum, M2), per-angle sum and M2, log-linear histograms,
* `minmax` -- (-min, max) of the per-sample scalars.

Second moments are centred (M2 = sum of squared deviations from the mean) and merged with Chan's pairwise update, so a
packed vector is not additive.  Samples shard trivially (contiguous ranges, `synthetic.shard_bounds`); the ONLY collective
of the whole path is in `merge()`: one all-gather of the packed vectors over torch.distributed (NCCL over NVLink on the GPU
box, gloo in the CPU tests), after which every rank merges the copies in rank order with the same kernel and holds the same
bits.  With `devices=[...]` one process drives several GPUs itself (no process group): the index range is split over the
devices and `result()` merges their vectors on the first one.  `MomentsResult` turns the packed vector into means,
variances and percentile estimates.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from .engine import _Batch, get_grid, host_to_device, resolve_devices, torr_2_pa

N_SCALARS = 12
N_MINMAX = 6
SCALAR_NAMES = ('V_cc', 'div_angle', 'T_c')
PILOT_SAMPLES = 4096


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
    def n_packed(self) -> int:
        return self.n_sums + N_MINMAX

    def c_struct(self) -> '_lib.HpemMomentsLayout':
        return _lib.HpemMomentsLayout(self.n_sums, self.off_angle_sum, self.off_angle_sumsq, self.off_hist,
                                      self.n_hist_angles, self.n_bins, N_MINMAX, 0)

    @property
    def hist_angle_index(self) -> np.ndarray:
        return np.arange(self.n_hist_angles) * max(self.spec.angle_stride, 1)

    def bin_edges(self) -> np.ndarray:
        """Upper edges of bins 0 .. n_bins-2 (bin 0 = underflow, last bin = overflow)."""
        sub = 1 << self.spec.sub_bits
        octaves = np.arange(self.spec.min_exp2, self.spec.max_exp2)
        edges = (2.0 ** octaves)[:, None] * (1.0 + np.arange(sub) / sub)[None, :]
        return np.concatenate([edges.reshape(-1), [2.0 ** self.spec.max_exp2]])


def _chan(n, s, m2, nb, sb, m2b):
    """Pairwise update of (n, S, M2) with (nb, Sb, M2b) (Chan et al.); n, nb scalars, the rest scalars or arrays."""
    if not nb > 0:
        return n, s, m2
    if not n > 0:
        return nb, sb, m2b
    delta = sb / nb - s / n
    return n + nb, s + sb, m2 + m2b + delta * delta * (n * nb / (n + nb))


def merge_packed_host(layout: Layout, parts: np.ndarray) -> np.ndarray:
    """NumPy statement of hpem_moments_merge: `parts` (n_parts, n_packed) -> one packed vector, parts in index order.
    Used for host (gloo / CPU-tensor) buffers and by the tests as the checker of the device kernel."""
    parts = np.asarray(parts, dtype=np.float64).reshape(-1, layout.n_packed)
    L, A = layout, layout.n_angles
    out = np.zeros(L.n_packed)
    out[:3] = parts[:, :3].sum(axis=0)
    for g in range(3):
        n, s, m2 = 0.0, 0.0, 0.0
        for p in parts:
            n, s, m2 = _chan(n, s, m2, float(p[3 + 3 * g]), float(p[4 + 3 * g]), float(p[5 + 3 * g]))
        out[3 + 3 * g: 6 + 3 * g] = (n, s, m2)
    n, s, m2 = 0.0, np.zeros(A), np.zeros(A)
    for p in parts:
        n, s, m2 = _chan(n, s, m2, float(p[0] - p[2]), p[L.off_angle_sum:L.off_angle_sum + A].copy(),
                         p[L.off_angle_sumsq:L.off_angle_sumsq + A].copy())
    out[L.off_angle_sum:L.off_angle_sum + A] = s
    out[L.off_angle_sumsq:L.off_angle_sumsq + A] = m2
    out[L.off_hist:L.n_sums] = parts[:, L.off_hist:L.n_sums].sum(axis=0)
    out[L.n_sums:] = parts[:, L.n_sums:].max(axis=0)
    return out


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
        n, s1, m2 = self.sums[3 + 3 * k: 6 + 3 * k]
        mean = s1 / n if n > 0 else np.nan
        var = m2 / n if n > 0 else np.nan
        return {'n': int(n), 'mean': mean, 'var': var, 'min': -self.minmax[2 * k], 'max': self.minmax[2 * k + 1]}

    @property
    def j_mean(self) -> np.ndarray:
        n = self.n_samples - self.n_nonfinite_rows
        L = self.layout
        return self.sums[L.off_angle_sum:L.off_angle_sum + L.n_angles] / max(n, 1)

    @property
    def j_var(self) -> np.ndarray:
        """Population variance of j_ion per angle (M2 / n; M2 is accumulated centred, see include/hpem.h)."""
        n = self.n_samples - self.n_nonfinite_rows
        L = self.layout
        return self.sums[L.off_angle_sumsq:L.off_angle_sumsq + L.n_angles] / max(n, 1)

    @property
    def histograms(self) -> np.ndarray:
        L = self.layout
        return self.sums[L.off_hist:L.off_hist + L.n_hist_angles * L.n_bins].reshape(L.n_hist_angles, L.n_bins)

    def j_percentile(self, q) -> np.ndarray:
        """Percentile estimates (per histogrammed angle) of j_ion over the samples -- what the reference's consumers take
        with np.percentile(j_ion, q, axis=0) (tests/test_plume.py:50-52, scripts/gen_data.py:402-404).  The rank
        q/100 (n-1) (NumPy's 'linear' definition) is located in the cumulative histogram and interpolated linearly inside
        its bin (the sub-bins of an octave are linear in j).  Ranks that fall into the under-/overflow bin return the
        histogram's lowest / highest edge."""
        L = self.layout
        h = self.histograms
        cum = np.cumsum(h, axis=1)
        total = cum[:, -1]
        upper = np.concatenate([L.bin_edges(), [np.inf]])                 # upper edge of bin b
        lower = np.concatenate([[0.0], L.bin_edges()])                    # lower edge of bin b
        q = np.atleast_1d(np.asarray(q, dtype=np.float64))
        out = np.empty((q.shape[0], L.n_hist_angles))
        rows = np.arange(L.n_hist_angles)
        for i, qq in enumerate(q):
            rank = qq / 100.0 * np.maximum(total - 1, 0)                   # 0-based fractional rank
            idx = np.array([np.searchsorted(cum[a], rank[a], side='right') for a in rows])
            idx = np.minimum(idx, L.n_bins - 1)
            below = np.where(idx > 0, cum[rows, np.maximum(idx - 1, 0)], 0.0)
            cnt = np.maximum(h[rows, idx], 1.0)
            frac = np.clip((rank - below + 0.5) / cnt, 0.0, 1.0)
            lo, hi = lower[idx], upper[idx]
            val = lo + frac * (hi - lo)
            val = np.where(idx == 0, lower[1], np.where(idx == L.n_bins - 1, lower[L.n_bins - 1], val))
            out[i] = val
        return out


class MonteCarloMoments:
    """Accumulates moments/histograms of the cathode+plume chain over sample chunks, on one GPU (`device=`) or -- from ONE
    process -- on several (`devices=[0, 1, ...]` or `'all'`): the sampled index range / the host arrays are split over
    the devices in contiguous 64-aligned shards, each device reduces its shard, and `result()` merges the vectors."""

    def __init__(self, n_angles: int = 91, sweep_radius: float = 1.0, hist: HistogramSpec = HistogramSpec(),
                 device: int | None = None, torr: float | None = None, want_cathode: bool = True, want_thrust: bool = True,
                 scalar_shift='auto', devices=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.torr = torr_2_pa() if torr is None else float(torr)
        self.hist = hist
        self.n_angles = int(n_angles)
        self.sweep_radius = float(sweep_radius)
        self.want_cathode, self.want_thrust = want_cathode, want_thrust
        self._auto_shift = isinstance(scalar_shift, str)
        self._shift = (0.0, 0.0, 0.0) if self._auto_shift else tuple(float(x) for x in scalar_shift)
        self.children: list[MonteCarloMoments] = []
        if devices is not None:
            devs = resolve_devices(devices)
            self.device = devs[0]
            self.children = [MonteCarloMoments(n_angles, sweep_radius, hist, d, self.torr, want_cathode, want_thrust,
                                               scalar_shift if not self._auto_shift else 'auto') for d in devs]
            self.layout = self.children[0].layout
            self.packed = torch.empty(self.layout.n_packed, dtype=torch.float64, device=f'cuda:{self.device}')
            self._views()
            self.reset()
            return
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.grid = get_grid(self.device, n_angles, np.atleast_1d(np.float64(sweep_radius)))
        self._make_spec()
        c_layout = _lib.HpemMomentsLayout()
        _lib.check(self.lib.hpem_moments_layout_query(self.grid.handle, ctypes.byref(self.spec), ctypes.byref(c_layout)))
        self.layout = Layout(n_angles, hist, c_layout)
        self.packed = torch.empty(self.layout.n_packed, dtype=torch.float64, device=f'cuda:{self.device}')
        self._views()
        self.reset()

    def _make_spec(self):
        h = self.hist
        self.spec = _lib.HpemMomentsSpec(h.angle_stride, h.sub_bits, h.min_exp2, h.max_exp2, int(self.want_cathode),
                                         int(self.want_thrust), (ctypes.c_double * 3)(*self._shift))

    def _views(self):
        self.sums = self.packed[:self.layout.n_sums]
        self.minmax = self.packed[self.layout.n_sums:]

    def reset(self):
        self.sums.zero_()
        self.minmax.fill_(-np.inf)
        for c in self.children:
            c.reset()

    # -- accumulation shift of the per-sample scalars: the means of a small pilot batch (first call only) --
    def _pilot(self, run) -> None:
        if not self._auto_shift:
            return
        self._auto_shift = False
        saved = self.packed.clone()
        self.reset()
        run()
        v = self.packed[:N_SCALARS].cpu().numpy()
        self.packed.copy_(saved)
        self._shift = tuple(float(v[4 + 3 * k] / v[3 + 3 * k]) if v[3 + 3 * k] > 0 else 0.0 for k in range(3))
        self._make_spec()

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def accumulate(self, inputs: dict) -> int:
        """Add one chunk of samples.  Single device: a dict of torch CUDA float64 tensors / scalars, asynchronous on the
        current stream.  `devices=`: a dict of host arrays (or scalars), split over the devices."""
        if self.children:
            return self._accumulate_split(inputs)
        names = tuple(_lib.CATHODE_INPUTS) if self.want_cathode else ()
        names += tuple(k for k in _lib.PLUME_INPUTS if k not in names)
        batch = _Batch(inputs, names, optional=('T',) if self.want_thrust else ())
        if self.want_thrust and 'T' not in batch.present:
            raise KeyError('T')
        if not batch.on_device:
            raise ValueError('MonteCarloMoments.accumulate expects device-resident (torch CUDA) inputs')
        if batch.device_index != self.device:
            raise ValueError(f'inputs live on cuda:{batch.device_index}, the reducer on cuda:{self.device}')

        def run(n):
            _lib.check(self.lib.hpem_moments_accumulate(self.grid.handle, n, ctypes.byref(batch.struct), self.torr,
                                                        ctypes.byref(self.spec), ctypes.c_void_p(self.sums.data_ptr()),
                                                        ctypes.c_void_p(self.minmax.data_ptr()), self._stream()))
        self._pilot(lambda: run(min(batch.n, PILOT_SAMPLES)))
        run(batch.n)
        return batch.n

    def accumulate_sampled(self, n: int, seed: int, first_index: int = 0, priors: dict | None = None) -> int:
        """Add samples [first_index, first_index + n) of the global index space, drawn on the fly by the on-device
        sampler (no input arrays exist); asynchronous on the current stream(s)."""
        from .sampler import SPT100_PRIORS, priors_struct
        from .synthetic import shard_bounds
        if self.children:
            for r, c in enumerate(self.children):
                lo, hi = shard_bounds(int(n), len(self.children), r)
                if hi > lo:
                    with self.torch.cuda.device(c.device):
                        c.accumulate_sampled(hi - lo, seed, first_index + lo, priors)
            return int(n)
        pr = priors_struct(priors or SPT100_PRIORS)

        def run(count, first):
            _lib.check(self.lib.hpem_moments_accumulate_sampled(
                self.grid.handle, int(count), int(seed), int(first), pr, self.torr, ctypes.byref(self.spec),
                ctypes.c_void_p(self.sums.data_ptr()), ctypes.c_void_p(self.minmax.data_ptr()), self._stream()))
        # the pilot always draws global indices [0, PILOT_SAMPLES): every rank / device / chunk derives the same shifts
        self._pilot(lambda: run(PILOT_SAMPLES, 0))
        run(n, first_index)
        return int(n)

    def _accumulate_split(self, inputs: dict) -> int:
        from .synthetic import shard_bounds
        torch = self.torch
        arrays = {k: np.asarray(v, dtype=np.float64) for k, v in inputs.items() if k in _lib.INPUT_NAMES}
        n = max((a.size for a in arrays.values()), default=1)
        for r, c in enumerate(self.children):
            lo, hi = shard_bounds(n, len(self.children), r)
            if hi <= lo:
                continue
            with torch.cuda.device(c.device):
                c.accumulate(host_to_device({k: (a.reshape(-1)[lo:hi] if a.size > 1 else a) for k, a in arrays.items()}, c.device))
        return n

    def _gather_children(self) -> None:
        """Bring the children's packed vectors to the first device and merge them there (fixed device order)."""
        torch = self.torch
        dev = f'cuda:{self.device}'
        for c in self.children:
            torch.cuda.synchronize(c.device)
        parts = torch.stack([c.packed.to(dev) for c in self.children])
        merge_packed(self.layout, parts, out=self.packed)

    def merge(self, group=None) -> None:
        """The path's only collective: all-gather the packed vectors over the ranks, then merge the copies in rank order
        (no-op without a process group).  Every rank ends up with the same bits."""
        import torch.distributed as dist
        if self.children:
            self._gather_children()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        merge_across_ranks(self.layout, self.packed, group)

    def result(self) -> MomentsResult:
        if self.children:
            self._gather_children()
        host = self.packed.cpu().numpy()
        return MomentsResult(self.layout, host[:self.layout.n_sums].copy(), host[self.layout.n_sums:].copy())


def merge_packed(layout: Layout, parts, out=None):
    """Merge `parts` (n_parts, n_packed) -- a torch tensor -- in index order into `out` (n_packed).  CUDA tensors go through
    libhpem's merge kernel (hpem_moments_merge); host tensors (the gloo tests) through the NumPy statement of the same update."""
    import torch
    parts = parts.reshape(-1, layout.n_packed)
    if out is None:
        out = torch.empty(layout.n_packed, dtype=torch.float64, device=parts.device)
    if parts.is_cuda:
        lib = _lib.load()
        parts = parts.contiguous()
        dev = parts.device.index
        c_lay = layout.c_struct()
        stream = torch.cuda.current_stream(dev).cuda_stream
        tmp = torch.empty(layout.n_packed, dtype=torch.float64, device=parts.device)     # the kernel's output must not alias its input
        _lib.check(lib.hpem_moments_merge(dev, ctypes.byref(c_lay), parts.shape[0], ctypes.c_void_p(parts.data_ptr()),
                                          layout.n_packed, ctypes.c_void_p(tmp.data_ptr()),
                                          ctypes.c_void_p(tmp.data_ptr() + 8 * layout.n_sums), ctypes.c_void_p(stream)))
        out.copy_(tmp)
    else:
        out.copy_(torch.from_numpy(merge_packed_host(layout, parts.numpy())))
    return out


def merge_across_ranks(layout: Layout, packed, group=None) -> None:
    """ONE all-gather of the packed [sums | minmax] vector, then the fixed-order merge on every rank; in place."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    gathered = torch.empty(world * layout.n_packed, dtype=torch.float64, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed.contiguous(), group=group)
    merge_packed(layout, gathered.reshape(world, layout.n_packed), out=packed)
