"""ctypes binding of libhpem.so (the C ABI declared in include/hpem.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded this module raises, and every
model entry point raises with it.  `python __graft_entry__.py` (or `build_library()`) compiles it in-tree with
`nvcc -gencode arch=compute_100a,code=sm_100a`.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / 'csrc'
LIB_PATH = PKG_DIR / 'lib' / 'libhpem.so'
HEADER = REPO_ROOT / 'include' / 'hpem.h'

INPUT_NAMES = ('P_b', 'V_a', 'T_e', 'V_vac', 'Pstar', 'P_T',
               'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0', 'T')   # enum hpem_input order
N_INPUTS = len(INPUT_NAMES)
CATHODE_INPUTS = INPUT_NAMES[:6]
PLUME_INPUTS = ('P_b',) + INPUT_NAMES[6:14]

HPEM_OK = 0
ABI_VERSION = 3
FLAG_FORCE_DIRECT = 1
FLAG_NO_TMA = 2
FLAG_LANES1 = 4
FLAG_LANES4 = 8
FLAG_NO_QUAD = 16
FLAG_NO_FASTMATH = 32
FLAG_NO_QTABLE = 64

EXPORTED_SYMBOLS = (
    'hpem_abi_version', 'hpem_source_hash', 'hpem_last_error', 'hpem_grid_create', 'hpem_grid_destroy', 'hpem_grid_is_uniform',
    'hpem_eval', 'hpem_eval_host', 'hpem_launch_count',
    'hpem_moments_layout_query', 'hpem_moments_accumulate', 'hpem_sample_inputs', 'hpem_moments_accumulate_sampled',
    'hpem_moments_merge', 'hpem_quadrature_table_eval',
    'hpem_measurements_create', 'hpem_measurements_destroy', 'hpem_loglike', 'hpem_logsumexp',
    'hpem_basis_create', 'hpem_basis_destroy', 'hpem_compress', 'hpem_compress_field', 'hpem_reconstruct',
)


class HpemInputs(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p * N_INPUTS), ('scalar', ctypes.c_double * N_INPUTS)]


class HpemOutputs(ctypes.Structure):
    _fields_ = [('V_cc', ctypes.c_void_p), ('j_ion', ctypes.c_void_p), ('div_angle', ctypes.c_void_p),
                ('T_c', ctypes.c_void_p), ('cos_div', ctypes.c_void_p), ('invalid', ctypes.c_void_p)]


class HpemMomentsSpec(ctypes.Structure):
    _fields_ = [('hist_angle_stride', ctypes.c_int32), ('hist_sub_bits', ctypes.c_int32),
                ('hist_min_exp2', ctypes.c_int32), ('hist_max_exp2', ctypes.c_int32),
                ('want_cathode', ctypes.c_int32), ('want_thrust', ctypes.c_int32), ('scalar_shift', ctypes.c_double * 3)]


class HpemMomentsLayout(ctypes.Structure):
    _fields_ = [('n_sums', ctypes.c_int64), ('off_angle_sum', ctypes.c_int64), ('off_angle_sumsq', ctypes.c_int64),
                ('off_hist', ctypes.c_int64), ('n_hist_angles', ctypes.c_int32), ('n_bins', ctypes.c_int32),
                ('n_minmax', ctypes.c_int32), ('reserved', ctypes.c_int32)]


class HpemPrior(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('reserved', ctypes.c_int32), ('a', ctypes.c_double), ('b', ctypes.c_double)]


PRIOR_CONST, PRIOR_UNIFORM, PRIOR_LOGUNIFORM, PRIOR_NORMAL = 0, 1, 2, 3


class HpemError(RuntimeError):
    """Raised for a non-zero status from the C ABI (API misuse or a CUDA error)."""


class LibraryMissing(ImportError):
    pass


HASH_MARKER = b'HPEM_SOURCE_HASH='


def _sources() -> list[Path]:
    return sorted(list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.inc')) + [HEADER])


def source_hash() -> str:
    """SHA-256 (first 16 hex digits) over the CUDA sources and the C header -- compiled into the library
    (`hpem_source_hash()`), so a binary that does not belong to this tree is detected whatever its mtime says."""
    import hashlib
    h = hashlib.sha256()
    for p in _sources():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def embedded_hash(path: Path) -> str | None:
    """The source hash compiled into a built library, read from the file (no dlopen); None if absent."""
    try:
        data = Path(path).read_bytes()
    except OSError:
        return None
    i = data.find(HASH_MARKER)
    if i < 0:
        return None
    return data[i + len(HASH_MARKER): i + len(HASH_MARKER) + 16].decode('ascii', 'replace')


def nvcc_command(out: Path = LIB_PATH) -> list[str]:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    # -fmad=false: every fused multiply-add in the kernels is written as fma(); the compiler never contracts a * b + c on
    # its own, so one expression gives the same bits in every kernel and template instantiation it is inlined into
    # (K2 from arrays == K2 with on-device sampling == K1u, bit for bit) and the reference's separately-rounded
    # multiply-adds stay separately rounded.  Measured cost: none (the hot loops are explicit fma / mul / add already).
    return [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-fmad=false',
            f'-DHPEM_SOURCE_HASH={source_hash()}', '-Xcompiler', '-fPIC', '-shared', '-o', str(out), str(CSRC / 'hpem_api.cu')]


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into lib/libhpem.so for sm_100a (cross-compiles without a GPU) unless the library on disk already
    carries the hash of the current sources.  Concurrent callers (the ranks of a multi-GPU launch) serialise on a lock file:
    one compiles, the others find the fresh library when they get the lock."""
    import fcntl
    want = source_hash()
    if LIB_PATH.exists() and not force and embedded_hash(LIB_PATH) == want:
        return LIB_PATH
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    with open(LIB_PATH.with_name('.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if LIB_PATH.exists() and not force and embedded_hash(LIB_PATH) == want:
            return LIB_PATH
        tmp = LIB_PATH.with_name(f'{LIB_PATH.name}.tmp.{os.getpid()}')
        cmd = nvcc_command(tmp)
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
        os.replace(tmp, LIB_PATH)
        if verbose:
            print(res.stderr)
    return LIB_PATH


_lock = threading.Lock()
_lib = None


def load() -> ctypes.CDLL:
    """Load libhpem.so once; raise LibraryMissing if it is not built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = Path(os.environ.get('HPEM_LIBRARY', LIB_PATH))
        if 'HPEM_LIBRARY' not in os.environ:
            # The library must be built from THIS tree's sources (hash compiled in).  Stale or missing: rebuild when nvcc is
            # here -- this IS the product, not a fallback -- and fail loudly otherwise; a failed rebuild never falls back to
            # the old binary.
            if embedded_hash(path) != source_hash():
                if not (shutil.which('nvcc') or Path('/usr/local/cuda/bin/nvcc').exists()):
                    raise LibraryMissing(f'{path} is missing or was not built from the sources in {CSRC} and nvcc is not '
                                         'available: build it with `python __graft_entry__.py`. No CPU fallback exists.')
                try:
                    build_library()
                except Exception as exc:  # noqa: BLE001
                    raise LibraryMissing(f'building {path} failed: {exc}') from exc
        if not path.exists():
            raise LibraryMissing(f'{path} not found: build it with `python __graft_entry__.py` '
                                 '(nvcc, sm_100a). hallthrusterpem_b200 has no CPU fallback.')
        lib = ctypes.CDLL(str(path))
        vp, i32, i64, dbl, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_uint32
        dptr = ctypes.POINTER(ctypes.c_double)
        lib.hpem_abi_version.restype = i32
        lib.hpem_source_hash.restype = ctypes.c_char_p
        lib.hpem_last_error.restype = ctypes.c_char_p
        lib.hpem_launch_count.restype = i64
        lib.hpem_grid_create.argtypes = [i32, i32, dptr, dptr, dptr, i32, dptr, ctypes.POINTER(vp)]
        lib.hpem_grid_create.restype = i32
        lib.hpem_grid_destroy.argtypes = [vp]
        lib.hpem_grid_destroy.restype = i32
        lib.hpem_grid_is_uniform.argtypes = [vp]
        lib.hpem_grid_is_uniform.restype = i32
        lib.hpem_eval.argtypes = [vp, i64, ctypes.POINTER(HpemInputs), ctypes.POINTER(HpemOutputs), dbl, u32, vp]
        lib.hpem_eval.restype = i32
        lib.hpem_eval_host.argtypes = [vp, i64, ctypes.POINTER(HpemInputs), ctypes.POINTER(HpemOutputs), dbl, u32]
        lib.hpem_eval_host.restype = i32
        lib.hpem_moments_layout_query.argtypes = [vp, ctypes.POINTER(HpemMomentsSpec), ctypes.POINTER(HpemMomentsLayout)]
        lib.hpem_moments_layout_query.restype = i32
        lib.hpem_moments_accumulate.argtypes = [vp, i64, ctypes.POINTER(HpemInputs), dbl, ctypes.POINTER(HpemMomentsSpec),
                                                vp, vp, vp]
        lib.hpem_moments_accumulate.restype = i32
        lib.hpem_moments_merge.argtypes = [i32, ctypes.POINTER(HpemMomentsLayout), i32, vp, i64, vp, vp, vp]
        lib.hpem_moments_merge.restype = i32
        lib.hpem_quadrature_table_eval.argtypes = [i32, dptr, dptr, i64, dptr, dptr, dptr]
        lib.hpem_quadrature_table_eval.restype = i32
        u64 = ctypes.c_uint64
        lib.hpem_sample_inputs.argtypes = [i32, i64, u64, u64, ctypes.POINTER(HpemPrior), ctypes.POINTER(vp), vp]
        lib.hpem_sample_inputs.restype = i32
        lib.hpem_moments_accumulate_sampled.argtypes = [vp, i64, u64, u64, ctypes.POINTER(HpemPrior), dbl,
                                                        ctypes.POINTER(HpemMomentsSpec), vp, vp, vp]
        lib.hpem_moments_accumulate_sampled.restype = i32
        lib.hpem_measurements_create.argtypes = [vp, i32, dptr, dptr, dptr, ctypes.POINTER(vp)]
        lib.hpem_measurements_create.restype = i32
        lib.hpem_measurements_destroy.argtypes = [vp]
        lib.hpem_measurements_destroy.restype = i32
        lib.hpem_loglike.argtypes = [vp, vp, i64, ctypes.POINTER(HpemInputs), dbl, vp, vp, vp]
        lib.hpem_loglike.restype = i32
        lib.hpem_logsumexp.argtypes = [i32, i64, i32, vp, vp, vp]
        lib.hpem_logsumexp.restype = i32
        lib.hpem_basis_create.argtypes = [i32, i32, i32, dptr, i32, ctypes.POINTER(vp)]
        lib.hpem_basis_create.restype = i32
        lib.hpem_basis_destroy.argtypes = [vp]
        lib.hpem_basis_destroy.restype = i32
        lib.hpem_compress.argtypes = [vp, vp, i64, ctypes.POINTER(HpemInputs), dbl, vp, vp]
        lib.hpem_compress.restype = i32
        lib.hpem_compress_field.argtypes = [vp, i64, vp, vp, vp]
        lib.hpem_compress_field.restype = i32
        lib.hpem_reconstruct.argtypes = [vp, i64, vp, vp, vp]
        lib.hpem_reconstruct.restype = i32
        if lib.hpem_abi_version() != ABI_VERSION:
            raise HpemError(f'libhpem ABI version {lib.hpem_abi_version()} != {ABI_VERSION}: rebuild with `python __graft_entry__.py`')
        if 'HPEM_LIBRARY' not in os.environ and lib.hpem_source_hash().decode() != source_hash():
            raise HpemError(f'{path} reports source hash {lib.hpem_source_hash().decode()}, the tree has {source_hash()}')
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != HPEM_OK:
        msg = load().hpem_last_error().decode('utf-8', 'replace')
        raise HpemError(f'libhpem status {status}: {msg}')
