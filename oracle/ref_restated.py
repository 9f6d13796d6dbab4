"""NumPy/SciPy restatement of the reference hot path (TEST INFRASTRUCTURE -- the CPU oracle).

Restates, operation for operation, what the reference computes:

* `cathode_coupling_oracle`  follows /root/reference/src/hallmd/models/cathode.py:24-38
* `current_density_oracle`   follows /root/reference/src/hallmd/models/plume.py:38-159

The only deliberate differences are the two things the reference hard-codes but BASELINE.json's
configs need as parameters:

* `n_angles`  -- plume.py:53 is `np.linspace(0, np.pi / 2, 91)`;
* `torr_2_pa` -- plume.py:12,40 / cathode.py:10,26-31 read `pem_core.constants.TORR_2_PA`, a value that
  is not present under /root/reference (un-vendored dependency).

At `n_angles=91` and the shim's `TORR_2_PA` this module is asserted BIT-IDENTICAL to the imported
reference (tests/test_oracle_vs_reference.py; oracle/make_golden.py refuses to write vectors
otherwise).  Every NumPy expression below therefore keeps the reference's evaluation order
(left-to-right products, `(alpha/a)**2`, `1 - decay`, `log(1 + x)` ...) -- reordering anything here
changes low-order bits.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import simpson
from scipy.special import erfi

DEFAULT_TORR_2_PA = 133.322  # see oracle/_shim/pem_core/constants.py
HALF_PI = np.pi / 2


def angle_grid(n_angles: int = 91) -> np.ndarray:
    """plume.py:53 with the angle count as a parameter."""
    return np.linspace(0, np.pi / 2, n_angles)


def cathode_coupling_oracle(inputs: dict, torr_2_pa: float = DEFAULT_TORR_2_PA) -> dict:
    """cathode.py:24-38.  V_cc = V_vac + T_e*log(1 + PB/PT) - (T_e/(PT+P*))*PB, clamped to [0, V_a]."""
    pb = inputs['P_b'] * torr_2_pa                      # cathode.py:26
    va = inputs['V_a']                                  # cathode.py:27
    te = inputs['T_e']                                  # cathode.py:28
    v_vac = inputs['V_vac']                             # cathode.py:29
    p_star = inputs['Pstar'] * torr_2_pa                # cathode.py:30
    p_t = inputs['P_T'] * torr_2_pa                     # cathode.py:31

    v = np.atleast_1d(v_vac + te * np.log(1 + pb / p_t) - (te / (p_t + p_star)) * pb)   # cathode.py:34
    v[v < 0] = 0                                        # cathode.py:35
    over = np.where(v > va)                             # cathode.py:36
    # cathode.py:37 indexes atleast_1d(Va) with `over`, which only works when V_a already has V_cc's shape;
    # broadcasting first is identical wherever the reference does not raise IndexError.
    v[over] = np.broadcast_to(np.atleast_1d(va), v.shape)[over]
    return {'V_cc': v}


def _beam_integral(alpha):
    """Denominator of A1/A2, plume.py:65-75 (and :77-85): the complex-erfi closed form of
    2*pi * int_0^{pi/2} exp(-(t/alpha)^2) sin(t) dt.  Returns complex128 exactly like the reference."""
    return (
        (np.pi ** (3 / 2))
        / 2
        * alpha
        * np.exp(-((alpha / 2) ** 2))
        * (
            2 * erfi(alpha / 2)
            + erfi((np.pi * 1j - (alpha**2)) / (2 * alpha))
            - erfi((np.pi * 1j + (alpha**2)) / (2 * alpha))
        )
    )


def current_density_oracle(inputs: dict, sweep_radius=1.0, n_angles: int = 91,
                           torr_2_pa: float = DEFAULT_TORR_2_PA, with_coords: bool = True,
                           return_internals: bool = False) -> dict:
    """plume.py:38-159.  Output layout is the reference's: j_ion (..., A) when one radius is given,
    (..., A, R) otherwise; div_angle/T_c (...,) or (..., R); j_ion_coords an object array of loop shape."""
    p_pa = inputs['P_b'] * torr_2_pa                    # plume.py:40
    c0, c1, c2, c3 = inputs['c0'], inputs['c1'], inputs['c2'], inputs['c3']
    c4, c5 = inputs['c4'], inputs['c5']
    sigma = inputs['sigma_cex']
    beam_current = inputs['I_B0']
    thrust = inputs.get('T', None)                      # plume.py:49
    radii = np.atleast_1d(sweep_radius)                 # plume.py:50

    theta = angle_grid(n_angles)                        # plume.py:53

    density = c4 * p_pa + c5                            # plume.py:56
    a_main = np.atleast_1d(c2 * p_pa + c3)              # plume.py:59
    a_main[a_main > np.pi / 2] = np.pi / 2              # plume.py:60 (upper clip only)
    a_scat = a_main / c1                                # plume.py:61

    with np.errstate(invalid='ignore', divide='ignore', over='ignore', under='ignore'):
        amp_main = (1 - c0) / _beam_integral(a_main)    # plume.py:64-76
        amp_scat = c0 / _beam_integral(a_scat)          # plume.py:77-85

        lift = (-1, -2)                                 # plume.py:87-93: (..., 1, 1)
        amp_main = np.expand_dims(amp_main, axis=lift)
        amp_scat = np.expand_dims(amp_scat, axis=lift)
        a_main = np.expand_dims(a_main, axis=lift)
        a_scat = np.expand_dims(a_scat, axis=lift)
        beam_current = np.expand_dims(beam_current, axis=lift)
        density = np.expand_dims(density, axis=lift)
        sigma = np.expand_dims(sigma, axis=lift)

        decay = np.exp(-radii * density * sigma)                               # plume.py:95
        j_cex = beam_current * (1 - decay) / (2 * np.pi * radii**2)            # plume.py:96
        base = beam_current * decay / radii**2                                 # plume.py:98
        th = theta[..., np.newaxis]
        j_main = base * amp_main * np.exp(-((th / a_main) ** 2))               # plume.py:99
        j_scat = base * amp_scat * np.exp(-((th / a_scat) ** 2))               # plume.py:100
        j_ion = j_main + j_scat + j_cex                                        # plume.py:102

    # plume.py:105-107 -- whole-sample invalidation; note div_angle / T_c below are NOT masked
    invalid = np.logical_or(np.any(a_main <= 0, axis=(-1, -2)), np.any(j_ion <= 0, axis=(-1, -2)))
    j_ion[invalid, ...] = 1e-20
    max_imag = float(np.max(np.abs(j_ion.imag))) if j_ion.size else 0.0        # plume.py:109-110 (warning only)
    j_ion = j_ion.real                                                         # plume.py:111

    # plume.py:117-119 -- integrands on the flipped profile, re-summed without j_cex
    beam_only = np.flip((j_main + j_scat).real, axis=-2)
    den_f = beam_only * np.cos(theta[..., np.newaxis])
    num_f = den_f * np.sin(theta[..., np.newaxis])

    with np.errstate(divide='ignore', invalid='ignore'):
        num = simpson(num_f, x=theta, axis=-2)                                 # plume.py:122
        den = simpson(den_f, x=theta, axis=-2)                                 # plume.py:123
        cos_div = np.atleast_1d(num / den)                                     # plume.py:124
        cos_div[cos_div == np.inf] = np.nan                                    # plume.py:125
        div_angle = np.arccos(cos_div)                                         # plume.py:127

    single_radius = radii.shape[0] == 1
    if single_radius:                                                          # plume.py:130-132
        j_ion = np.squeeze(j_ion, axis=-1)
        div_angle = np.squeeze(div_angle, axis=-1)

    out = {'j_ion': j_ion, 'div_angle': div_angle}
    if thrust is not None:                                                     # plume.py:136-140
        t_c = np.expand_dims(thrust, axis=-1) * cos_div
        out['T_c'] = np.squeeze(t_c, axis=-1) if single_radius else t_c

    if with_coords:                                                            # plume.py:152-157
        loop_shape = j_ion.shape[:-1] if single_radius else j_ion.shape[:-2]
        coords = np.empty(loop_shape, dtype=object)
        for idx in np.ndindex(loop_shape):
            coords[idx] = theta
        out['j_ion_coords'] = coords

    if return_internals:  # extras for tests only (never part of the reference's return value)
        out['_cos_div'] = np.squeeze(cos_div, axis=-1) if single_radius else cos_div
        out['_invalid'] = invalid
        out['_max_imag'] = max_imag
    return out


def simpson_weights(theta: np.ndarray) -> np.ndarray:
    """W such that simpson(y, x=theta) == W @ y up to rounding (SciPy's non-uniform branch incl. the
    even-N last-interval correction).  Host-side helper; mirrors what the product precomputes."""
    return simpson(np.eye(theta.shape[0]), x=theta, axis=-1)
