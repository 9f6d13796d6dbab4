"""On-device sampler for the PEM v0 input priors (the step before the hot path: amisc's `system.sample_inputs`,
/root/reference/scripts/gen_data.py:238, over the priors of scripts/pem_v0/pem_v0_SPT-100.yml).

Counter-based Philox4x32-10: the 15 inputs of global sample index i are a pure function of (seed, i), so any sharding
of an index range over GPUs / chunks reproduces the unsharded draw.  The draws happen in libhpem; the NumPy statement of the
same stream that the tests check them against lives with the other checkers, in oracle/sampler_oracle.py."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

# (kind, a, b) per input; priors of pem_v0_SPT-100.yml, test ranges where the YAML has only a domain (see synthetic.py)
SPT100_PRIORS = {
    'P_b': ('loguniform', 1e-8, 1e-4),      # yml:15 domain, log10-normalised
    'V_a': ('uniform', 200.0, 400.0),       # yml:24
    'T_e': ('uniform', 1.0, 5.0),           # yml:31
    'V_vac': ('uniform', 0.0, 60.0),        # yml:38
    'Pstar': ('uniform', 10e-6, 100e-6),    # yml:45
    'P_T': ('uniform', 10e-6, 100e-6),      # yml:53
    'c0': ('uniform', 0.1, 0.9),            # tests/test_plume.py:21 (yml:226 is U(0,1))
    'c1': ('uniform', 0.1, 0.9),            # yml:232
    'c2': ('uniform', -15.0, 15.0),         # yml:239
    'c3': ('uniform', 0.2, 1.570796),       # yml:246
    'c4': ('loguniform', 1e18, 1e22),       # yml:253
    'c5': ('loguniform', 1e14, 1e18),       # yml:261
    'sigma_cex': ('uniform', 51e-20, 58e-20),   # yml:269
    'I_B0': ('uniform', 2.0, 8.0),          # tests/test_plume.py:28
    'T': ('uniform', 0.02, 0.12),           # yml:183-184
}
_KINDS = {'const': _lib.PRIOR_CONST, 'uniform': _lib.PRIOR_UNIFORM, 'loguniform': _lib.PRIOR_LOGUNIFORM,
          'normal': _lib.PRIOR_NORMAL}


def priors_struct(priors: dict):
    arr = (_lib.HpemPrior * _lib.N_INPUTS)()
    for k, name in enumerate(_lib.INPUT_NAMES):
        kind, a, b = priors.get(name, ('const', 0.0, 0.0))
        arr[k] = _lib.HpemPrior(_KINDS[kind], 0, float(a), float(b))
    return arr


def sample_inputs(n: int, seed: int, first_index: int = 0, priors: dict = SPT100_PRIORS, device: int | None = None) -> dict:
    """Draw samples [first_index, first_index + n) of every input named in `priors` into torch CUDA float64 tensors."""
    import torch
    lib = _lib.load()
    dev = torch.cuda.current_device() if device is None else int(device)
    out = {name: torch.empty(n, dtype=torch.float64, device=f'cuda:{dev}') for name in _lib.INPUT_NAMES if name in priors}
    ptrs = (ctypes.c_void_p * _lib.N_INPUTS)()
    for k, name in enumerate(_lib.INPUT_NAMES):
        ptrs[k] = out[name].data_ptr() if name in out else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.hpem_sample_inputs(dev, n, seed, first_index, priors_struct(priors), ptrs, ctypes.c_void_p(stream)))
    return out
