"""Host side of the plume + cathode path: broadcasting, buffer carriers, grid handles, C-ABI calls.

Mirrors the calling convention of the reference's model functions
(/root/reference/src/hallmd/models/plume.py:21, cathode.py:16): a dict of named inputs (Python scalars, NumPy
arrays of any common loop shape, or -- new -- torch CUDA float64 tensors) in, a dict of outputs out.

* host inputs  -> `hpem_eval_host` (chunked H2D / kernel / D2H pipeline inside the library) -> NumPy outputs
  (backed by pinned memory so the D2H DMA goes straight into the returned arrays);
* torch CUDA inputs -> `hpem_eval` on torch's current stream, zero-copy via `data_ptr()` -> torch CUDA outputs.

PyTorch is only the buffer carrier (device allocations, pinned host memory, streams).  All arithmetic happens
in libhpem's CUDA kernels; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Any

import numpy as np

from . import _lib
from .quadrature import angle_grid, fused_weights

DEFAULT_TORR_2_PA = 133.322
"""pem_core.constants.TORR_2_PA is not vendored with the reference (uv.lock:1655-1657); 133.322 is the value hallmd
used historically.  If `pem_core` is importable its value is used instead (see `torr_2_pa()`)."""

_PIN_THRESHOLD_BYTES = 1 << 20


_torr_cache: float | None = None


def torr_2_pa() -> float:
    """pem_core's constant when `pem_core` is importable (the real dependency is authoritative), else 133.322.
    Resolved once: a failing import costs ~100 us per attempt, more than a small batch takes on the GPU."""
    global _torr_cache
    if _torr_cache is None:
        try:
            from pem_core.constants import TORR_2_PA  # type: ignore
            _torr_cache = float(TORR_2_PA)
        except Exception:
            _torr_cache = DEFAULT_TORR_2_PA
    return _torr_cache


def _torch():
    import torch
    return torch


def resolve_devices(devices) -> list[int]:
    """'all' -> every visible CUDA device; an int -> [int]; an iterable of ints -> list."""
    torch = _torch()
    if isinstance(devices, str):
        if devices != 'all':
            raise ValueError(f"devices must be 'all', an index or a list of indices, got {devices!r}")
        devs = list(range(torch.cuda.device_count()))
    elif isinstance(devices, int):
        devs = [devices]
    else:
        devs = [int(d) for d in devices]
    if not devs:
        raise RuntimeError('hallthrusterpem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return devs


def _is_torch_tensor(x: Any) -> bool:
    return type(x).__module__.startswith('torch') and hasattr(x, 'data_ptr')


# ----------------------------------------------------------------------------------------------
# grid handles (angle grid + fused weights + radii live on the device, cached per (device, A, radii))
# ----------------------------------------------------------------------------------------------
class GridHandle:
    def __init__(self, device: int, n_angles: int, radii: np.ndarray, alpha: np.ndarray | None = None):
        lib = _lib.load()
        self.device = int(device)
        self.alpha = angle_grid(n_angles) if alpha is None else np.ascontiguousarray(alpha, dtype=np.float64)
        self.alpha.setflags(write=False)
        self.n_angles = int(self.alpha.shape[0])
        self.radii = np.ascontiguousarray(radii, dtype=np.float64).reshape(-1)
        self.n_radii = int(self.radii.shape[0])
        wd, wn = fused_weights(self.alpha)
        dptr = ctypes.POINTER(ctypes.c_double)
        handle = ctypes.c_void_p()
        _lib.check(lib.hpem_grid_create(self.device, self.n_angles, self.alpha.ctypes.data_as(dptr),
                                        wd.ctypes.data_as(dptr), wn.ctypes.data_as(dptr), self.n_radii,
                                        self.radii.ctypes.data_as(dptr), ctypes.byref(handle)))
        self._h = handle
        self.uniform = bool(lib.hpem_grid_is_uniform(handle))
        self.host_lock = threading.Lock()

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h

    def close(self):
        if self._h:
            _lib.load().hpem_grid_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_grid_cache: dict[tuple, GridHandle] = {}
_grid_lock = threading.Lock()


def get_grid(device: int, n_angles: int, radii: np.ndarray) -> GridHandle:
    radii = np.ascontiguousarray(np.atleast_1d(radii), dtype=np.float64)
    key = (int(device), int(n_angles), radii.tobytes())
    with _grid_lock:
        g = _grid_cache.get(key)
        if g is None:
            if len(_grid_cache) >= 64:
                # evict the oldest entry from the cache only: callers that still hold the handle (PreparedCall, reducers,
                # measurement sets, another thread inside hpem_eval_host) keep it alive; GridHandle.__del__ destroys it
                # when the last of them lets go
                _grid_cache.pop(next(iter(_grid_cache)))
            g = _grid_cache[key] = GridHandle(device, n_angles, radii)
        return g


# ----------------------------------------------------------------------------------------------
# input marshalling
# ----------------------------------------------------------------------------------------------
_INPUT_INDEX = {name: k for k, name in enumerate(_lib.INPUT_NAMES)}


class _Batch:
    """Broadcast the named inputs to a common loop shape and build the hpem_inputs struct.
    (Kept lean: for the small batches amisc often passes, this marshalling is most of the call's latency.)"""

    def __init__(self, inputs: dict, names: tuple[str, ...], optional: tuple[str, ...] = ()):
        vals = {}
        for name in names:
            vals[name] = inputs[name]                     # KeyError for a missing key, like the reference
        for name in optional:
            v = inputs.get(name, None)
            if v is not None:
                vals[name] = v
        # classify once: (value, is_torch, shape)
        info = {}
        devs = set()
        shapes = set()
        for name, v in vals.items():
            is_t = _is_torch_tensor(v)
            shape = tuple(v.shape) if (is_t or hasattr(v, 'shape')) else ()
            if is_t and v.is_cuda:
                devs.add(v.device.index)
            info[name] = (v, is_t, shape)
            shapes.add(shape)
        self.on_device = bool(devs)
        self.device_index = None
        if self.on_device:
            if len(devs) != 1:
                raise ValueError(f'inputs live on several CUDA devices: {sorted(devs)}')
            self.device_index = devs.pop()
        if len(shapes) == 1:
            self.loop_shape = next(iter(shapes))
        else:
            self.loop_shape = tuple(np.broadcast_shapes(*shapes)) if shapes else ()
        self.out_shape = self.loop_shape if len(self.loop_shape) > 0 else (1,)   # np.atleast_1d (plume.py:59)
        n = 1
        for d in self.out_shape:
            n *= int(d)
        self.n = n
        self.struct = struct = _lib.HpemInputs()
        self._keep = keep = []                            # keep converted buffers alive during the call
        out_shape = self.out_shape
        for name, (v, is_t, shape) in info.items():
            k = _INPUT_INDEX[name]
            size = 1
            for d in shape:
                size *= int(d)
            if size == 1 and not (is_t and v.is_cuda):    # NumPy scalar broadcasting (test_plume.py:67-77)
                struct.ptr[k] = None
                struct.scalar[k] = float(v.reshape(-1)[0]) if shape != () or hasattr(v, 'reshape') else float(v)
                continue
            # (a one-element CUDA tensor stays a pointer, expanded to the loop shape: reading its value here would be a
            #  blocking device->host copy, break CUDA-graph capture and bake a stale value into a cached PreparedCall)
            if self.on_device:
                torch = _torch()
                t = v if is_t else torch.as_tensor(np.asarray(v, dtype=np.float64))
                if t.dtype != torch.float64 or not t.is_cuda or t.device.index != self.device_index:
                    t = t.to(device=f'cuda:{self.device_index}', dtype=torch.float64)
                if shape != out_shape:
                    t = t.broadcast_to(out_shape)
                if not t.is_contiguous():
                    t = t.contiguous()
                keep.append(t)
                struct.ptr[k] = t.data_ptr()
            else:
                a = v.detach().cpu().numpy() if is_t else v
                if not (type(a) is np.ndarray and a.dtype == np.float64 and a.flags.c_contiguous and shape == out_shape):
                    a = np.asarray(a, dtype=np.float64)
                    if a.shape != out_shape:
                        a = np.broadcast_to(a, out_shape)
                    a = np.ascontiguousarray(a)
                keep.append(a)
                struct.ptr[k] = a.__array_interface__['data'][0]
        self.present = set(vals)


class _PinnedPool:
    """Free-list of page-locked host buffers behind the NumPy outputs of the host path.

    Pinning memory costs about as much as copying into it (cudaHostAlloc: ~15 us per MB), so the buffers must be reused
    across calls.  torch's caching host allocator does that when the previous result is dropped BEFORE the next call, but
    in the usual loop `out = current_density(inputs)` the previous result is still alive while the next one is allocated,
    and every call then pins fresh memory (measured: 11 ms per call for a 728 MB `j_ion`).  Here a buffer returns to the
    free-list when the last NumPy view of it dies (weakref.finalize on the root array), so that loop settles on two
    alternating buffers."""

    def __init__(self, max_cached_bytes: int = 8 << 30):
        self._free: dict[int, list] = {}
        self._cached = 0
        self._max = max_cached_bytes
        self._lock = threading.Lock()

    def _release(self, nbytes: int, tensor) -> None:
        with self._lock:
            if self._cached + nbytes <= self._max:
                self._free.setdefault(nbytes, []).append(tensor)
                self._cached += nbytes

    def take(self, shape: tuple[int, ...], dtype) -> np.ndarray:
        import weakref
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        with self._lock:
            lst = self._free.get(nbytes)
            tensor = lst.pop() if lst else None
            if tensor is not None:
                self._cached -= nbytes
        if tensor is None:
            torch = _torch()
            tensor = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        root = tensor.numpy()                      # every view handed out collapses its `.base` chain onto this array
        weakref.finalize(root, self._release, nbytes, tensor)
        return root.view(dt).reshape(shape)


_pinned_pool = _PinnedPool()


def _alloc_host(shape: tuple[int, ...], dtype=np.float64) -> np.ndarray:
    """Host output buffer; large ones are page-locked (pooled, see _PinnedPool) and the D2H copy DMAs straight into the
    array that is returned."""
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    if nbytes >= _PIN_THRESHOLD_BYTES:
        return _pinned_pool.take(shape, dtype)
    return np.empty(shape, dtype=dtype)


_pool = None
_pool_lock = threading.Lock()


def _thread_pool():
    """Issuing threads of the single-process multi-GPU host path (ctypes releases the GIL inside the C call)."""
    global _pool
    with _pool_lock:
        if _pool is None:
            from concurrent.futures import ThreadPoolExecutor
            _pool = ThreadPoolExecutor(max_workers=16, thread_name_prefix='hpem-dev')
        return _pool


_copy_pool = None


def _copy_threads():
    """Threads for host-side staging copies (separate from the per-device issuing threads, which may call into them)."""
    global _copy_pool
    with _pool_lock:
        if _copy_pool is None:
            from concurrent.futures import ThreadPoolExecutor
            _copy_pool = ThreadPoolExecutor(max_workers=8, thread_name_prefix='hpem-copy')
        return _copy_pool


def host_to_device(arrays: dict, device: int) -> dict:
    """Host float64 arrays -> CUDA tensors on `device` for the entry points whose kernels take device buffers (latents,
    log-likelihood).  A cudaMemcpy from pageable memory is staged by the driver at a few GB/s and blocks the caller; here
    the arrays are copied into pooled page-locked buffers by parallel threads (NumPy releases the GIL) and shipped with
    asynchronous copies on torch's current stream: 72 MB of inputs in ~3 ms instead of ~15 ms.  One-element arrays and
    Python scalars are returned as floats (the kernels broadcast them)."""
    torch = _torch()
    out, jobs = {}, []
    for name, v in arrays.items():
        a = np.asarray(v, dtype=np.float64)
        if a.size <= 1:
            out[name] = float(a.reshape(-1)[0]) if a.size == 1 else a
            continue
        stage = _alloc_host(a.shape, np.float64)
        jobs.append((name, a, stage))
    if jobs:
        if sum(a.nbytes for _, a, _ in jobs) >= (8 << 20) and len(jobs) > 1:
            list(_copy_threads().map(lambda j: np.copyto(j[2], j[1]), jobs))
        else:
            for _, a, stage in jobs:
                np.copyto(stage, a)
        with torch.cuda.device(device):
            for name, _, stage in jobs:
                out[name] = torch.from_numpy(stage).to(f'cuda:{device}', non_blocking=True)
            torch.cuda.current_stream(device).synchronize()      # the staging buffers go back to the pool when `jobs` dies
    return out


def device_to_host(t) -> np.ndarray:
    """CUDA tensor -> NumPy array backed by pooled page-locked memory (the D2H copy is a DMA straight into the result)."""
    torch = _torch()
    host = _alloc_host(tuple(t.shape), np.float64)
    dst = torch.from_numpy(host)
    dst.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device.index).synchronize()
    return host


class PreparedCall:
    """One marshalled request: input struct, output buffers and grid handle.  `run()` issues the C-ABI call and may
    be repeated (same buffers) -- bench.py times exactly this; `evaluate()` is prepare + run + results.

    `device='all'` (or a list of indices) with HOST inputs shards the samples over several GPUs from this ONE process:
    contiguous ranges cut at multiples of 64 samples (so every sample keeps its position modulo 4 / 32 and the outputs
    are bit-identical to a single-GPU call), one `hpem_eval_host` pipeline per device, each writing its slice of the same
    output arrays.  This is how the reference's actual caller -- amisc, one process, one call per batch
    (pem_v0_SPT-100.yml:5-7,215-218) -- reaches all GPUs of the box."""

    def __init__(self, inputs: dict, *, want_cathode: bool, want_plume: bool, sweep_radius=1.0, n_angles: int = 91,
                 torr: float | None = None, device=None, direct: bool = False,
                 want_j_ion: bool = True, extras: bool = False, pin_outputs: bool = True, no_tma: bool = False,
                 lanes1: bool = False, lanes4: bool = False, no_quad: bool = False, no_fastmath: bool = False,
                 no_qtable: bool = False):
        self.lib = _lib.load()
        names: tuple[str, ...] = ()
        if want_cathode:
            names += _lib.CATHODE_INPUTS
        if want_plume:
            names += tuple(k for k in _lib.PLUME_INPUTS if k not in names)
        self.batch = batch = _Batch(inputs, names, optional=('T',) if want_plume else ())
        has_thrust = 'T' in batch.present
        if want_plume and batch.n == 0:
            # the reference fails on an empty sample batch, too: scipy.integrate.simpson (plume.py:122) cannot broadcast
            # its slices of a (0, A, R) integrand.  cathode_coupling alone returns an empty V_cc, here as there.
            raise ValueError('operands could not be broadcast together: current_density needs at least one sample '
                             '(the reference raises the same from scipy.integrate.simpson, plume.py:122)')
        self.torr = torr_2_pa() if torr is None else float(torr)
        self.want_plume = want_plume

        radii = np.atleast_1d(np.asarray(sweep_radius, dtype=np.float64)).reshape(-1)
        n_radii = int(radii.shape[0])
        single = n_radii == 1                                  # plume.py:130 squeezes the radius axis

        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError('hallthrusterpem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        devs = None
        if batch.on_device:
            dev = batch.device_index
        elif device is None:
            dev = torch.cuda.current_device()
        elif isinstance(device, int):
            dev = int(device)
        else:
            devs = resolve_devices(device)
            dev = devs[0]
        self.device = dev
        g_angles, g_radii = (n_angles, radii) if want_plume else (91, np.array([1.0]))
        self.grid = grid = get_grid(dev, g_angles, g_radii)

        loop = batch.out_shape
        rshape = loop if single else loop + (n_radii,)
        jshape = loop + (grid.n_angles,) if single else loop + (grid.n_angles, n_radii)
        self.out = out = _lib.HpemOutputs()
        self.result: dict[str, Any] = {}
        result = self.result

        def new(shape, dtype=np.float64):
            if batch.on_device:
                tdtype = torch.float64 if dtype == np.float64 else torch.uint8
                t = torch.empty(shape, dtype=tdtype, device=f'cuda:{dev}')
                return t, t.data_ptr()
            a = _alloc_host(shape, dtype) if pin_outputs else np.empty(shape, dtype=dtype)
            return a, a.ctypes.data

        if want_cathode:
            result['V_cc'], out.V_cc = new(loop)
        if want_plume:
            if want_j_ion:
                result['j_ion'], out.j_ion = new(jshape)
            result['div_angle'], out.div_angle = new(rshape)
            if has_thrust:
                result['T_c'], out.T_c = new(rshape)
            if extras:
                result['cos_div'], out.cos_div = new(rshape)
                result['invalid'], out.invalid = new(loop, np.uint8)
        self.flags = (_lib.FLAG_FORCE_DIRECT if direct else 0) | (_lib.FLAG_NO_TMA if no_tma else 0) \
            | (_lib.FLAG_LANES1 if lanes1 else 0) | (_lib.FLAG_LANES4 if lanes4 else 0) | (_lib.FLAG_NO_QUAD if no_quad else 0) \
            | (_lib.FLAG_NO_FASTMATH if no_fastmath else 0) | (_lib.FLAG_NO_QTABLE if no_qtable else 0)
        self.h2d_bytes = 0 if batch.on_device else 8 * batch.n * sum(1 for k in range(_lib.N_INPUTS)
                                                                      if batch.struct.ptr[k])
        self.d2h_bytes = 0 if batch.on_device else sum(v.nbytes for v in result.values())
        # single-process multi-GPU: per-device views of the same host buffers
        self.shards = []
        if devs is not None and len(devs) > 1 and batch.n >= 64 * len(devs):
            from .synthetic import shard_bounds
            row = grid.n_angles * n_radii
            per_out = {'V_cc': 8, 'j_ion': 8 * row, 'div_angle': 8 * n_radii, 'T_c': 8 * n_radii, 'cos_div': 8 * n_radii, 'invalid': 1}
            for r, d in enumerate(devs):
                lo, hi = shard_bounds(batch.n, len(devs), r)
                if hi <= lo:
                    continue
                si, so = _lib.HpemInputs(), _lib.HpemOutputs()
                for k in range(_lib.N_INPUTS):
                    si.ptr[k] = (batch.struct.ptr[k] + 8 * lo) if batch.struct.ptr[k] else None
                    si.scalar[k] = batch.struct.scalar[k]
                for name, stride in per_out.items():
                    base = getattr(out, name)
                    setattr(so, name, (base + stride * lo) if base else None)
                self.shards.append((get_grid(d, g_angles, g_radii), hi - lo, si, so))

    def run(self, stream: int | None = None) -> None:
        """Device inputs: asynchronous launch on `stream` (default: torch's current stream).
        Host inputs: returns when every output has landed in host memory."""
        b = self.batch
        if b.n == 0:          # empty batch (cathode only): nothing to launch, and empty buffers have no address to pass
            return
        if b.on_device:
            torch = _torch()
            if stream is None:
                stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self.lib.hpem_eval(self.grid.handle, b.n, ctypes.byref(b.struct), ctypes.byref(self.out),
                                          self.torr, self.flags, ctypes.c_void_p(stream)))
        elif self.shards:
            def one(shard):
                g, count, si, so = shard
                return self.lib.hpem_eval_host(g.handle, count, ctypes.byref(si), ctypes.byref(so), self.torr, self.flags), \
                    self.lib.hpem_last_error().decode('utf-8', 'replace')      # thread-local text, read on the issuing thread
            for status, msg in list(_thread_pool().map(one, self.shards)):
                if status != _lib.HPEM_OK:
                    raise _lib.HpemError(f'libhpem status {status}: {msg}')
        else:
            _lib.check(self.lib.hpem_eval_host(self.grid.handle, b.n, ctypes.byref(b.struct), ctypes.byref(self.out),
                                               self.torr, self.flags))

    def results(self) -> dict:
        res = dict(self.result)
        if self.want_plume:
            # plume.py:152-157: an object array of loop shape whose every element is the SAME angle-grid ndarray.  The
            # reference fills it with a Python loop (O(n)); here it is a zero-stride (read-only) broadcast view of one
            # 0-d object array -- same shape, dtype, element identity and indexing behaviour, O(1) to build.
            cell = np.empty((), dtype=object)
            cell[()] = self.grid.alpha
            res['j_ion_coords'] = np.broadcast_to(cell, self.batch.out_shape)
        return res


def evaluate(inputs: dict, **kwargs) -> dict:
    """Run the fused kernel for the requested output groups.  Returns outputs keyed like the reference."""
    call = PreparedCall(inputs, **kwargs)
    call.run()
    return call.results()
