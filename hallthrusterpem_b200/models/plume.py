"""Plume model -- drop-in for `hallmd.models.plume.current_density`
(/root/reference/src/hallmd/models/plume.py:21-159), evaluated by libhpem's CUDA kernels."""
from __future__ import annotations

from ..engine import evaluate

__all__ = ['current_density', 'plume_cathode']


def current_density(inputs: dict, sweep_radius=1.0, *, n_angles: int = 91, torr_2_pa: float | None = None,
                    device=None, direct: bool = False, extras: bool = False, no_tma: bool = False,
                    lanes1: bool = False, lanes4: bool = False, no_quad: bool = False, no_fastmath: bool = False,
                    no_qtable: bool = False) -> dict:
    """Semi-empirical ion current density (j_ion) plume model over a 90 deg sweep (0 deg = thruster centerline),
    plus the plume divergence angle and, if `T` is given, the divergence-corrected thrust.

    :param inputs: `P_b`, `c0`, `c1`, `c2`, `c3`, `c4`, `c5`, `sigma_cex`, `I_B0` (and optionally `T`): Python
                   scalars, NumPy arrays of a common (broadcastable) loop shape, or torch CUDA float64 tensors.
    :param sweep_radius: radius/radii (m) of the sweep; with several radii `j_ion` gains a trailing radius axis.
    :param n_angles: number of sweep angles (keyword-only extra; the reference hard-codes 91, plume.py:53).
    :param torr_2_pa: value of `pem_core.constants.TORR_2_PA` (keyword-only extra).
    :param device: CUDA device index for host inputs; 'all' or a list of indices shards the samples over several GPUs from
                   this one process, results bit-identical to one GPU (keyword-only extra).
    :param direct: force the reference-operation-order kernel instead of the recurrence kernel (diagnostics).
    :param lanes1: force the one-lane-per-sample sweep kernel (K1u); `lanes4` forces the four-lane one (K1v).
                   By default the library picks per angle count (diagnostics).
    :param no_tma: stage `j_ion` with plain global stores instead of TMA tensor stores (diagnostics).
    :param no_quad: angle counts that are not a multiple of 4: skip the quad-row tensor stores (diagnostics).
    :param no_fastmath: per-sample part through libdevice exp/log/acos and IEEE division for every warp (diagnostics; the
                   default uses the branch-free functions of csrc/hpem_fastmath.cuh for warps in the nominal range).
    :param no_qtable: accumulate the two Simpson sums of plume.py:121-122 angle by angle instead of taking them from the
                   grid's table (diagnostics; csrc/hpem_qtable.cuh, the two agree to ~4e-16).
    :param extras: also return `cos_div` and the whole-sample `invalid` mask (plume.py:105,124).
    :returns outputs: `j_ion` (..., A[, R]), `div_angle` (...[, R]), optionally `T_c`, and `j_ion_coords`
                      (object array of loop shape whose elements are the angle grid in radians).
    """
    return evaluate(inputs, want_cathode=False, want_plume=True, sweep_radius=sweep_radius, n_angles=n_angles,
                    torr=torr_2_pa, device=device, direct=direct, extras=extras, no_tma=no_tma,
                    lanes1=lanes1, lanes4=lanes4, no_quad=no_quad, no_fastmath=no_fastmath, no_qtable=no_qtable)


def plume_cathode(inputs: dict, sweep_radius=1.0, *, n_angles: int = 91, torr_2_pa: float | None = None,
                  device=None, direct: bool = False, extras: bool = False,
                  want_j_ion: bool = True, no_tma: bool = False, lanes1: bool = False,
                  lanes4: bool = False, no_quad: bool = False, no_fastmath: bool = False, no_qtable: bool = False) -> dict:
    """The PEM v0 chain Cathode -> (Thruster, external) -> Plume in ONE fused launch over the same samples
    (pem_v0_SPT-100.yml:5,62,215): returns `V_cc` together with the plume outputs.  `P_b` is loaded once."""
    return evaluate(inputs, want_cathode=True, want_plume=True, sweep_radius=sweep_radius, n_angles=n_angles,
                    torr=torr_2_pa, device=device, direct=direct, extras=extras, want_j_ion=want_j_ion,
                    no_tma=no_tma, lanes1=lanes1, lanes4=lanes4, no_quad=no_quad, no_fastmath=no_fastmath, no_qtable=no_qtable)
