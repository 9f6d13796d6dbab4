"""Cathode coupling model -- drop-in for `hallmd.models.cathode.cathode_coupling`
(/root/reference/src/hallmd/models/cathode.py:16-38), evaluated by libhpem's CUDA kernel."""
from __future__ import annotations

from ..engine import evaluate

__all__ = ['cathode_coupling']


def cathode_coupling(inputs: dict, *, torr_2_pa: float | None = None, device=None) -> dict:
    """Computes cathode coupling voltage dependence on background pressure.

    :param inputs: input arrays - `P_b`, `V_a`, `T_e`, `V_vac`, `Pstar`, `P_T` for background pressure (Torr),
                   discharge voltage (V), electron temperature (eV), vacuum coupling voltage (V), and model
                   parameters P* (Torr) and P_T (Torr).  Python scalars, NumPy arrays of a common (broadcastable)
                   loop shape, or torch CUDA float64 tensors (zero-copy, output stays on the device).
    :param torr_2_pa: value of `pem_core.constants.TORR_2_PA` (keyword-only extra; default 133.322 or pem_core's).
    :param device: CUDA device index for host inputs, or 'all' / a list of indices to shard the samples over several GPUs
                   (keyword-only extra; default: torch's current device).
    :returns outputs: `V_cc` for cathode coupling voltage (V), always at least 1-D (cathode.py:34).

    Where the reference raises IndexError (scalar `V_a` with array inputs and an active upper clamp,
    cathode.py:37) this implementation broadcasts `V_a`; everywhere else results are identical.
    """
    out = evaluate(inputs, want_cathode=True, want_plume=False, torr=torr_2_pa, device=device)
    return {'V_cc': out['V_cc']}
