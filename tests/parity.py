"""Parity rules shared by the tests (the tolerances of BASELINE.json's north_star, written down once).

north_star: "Outputs must match the reference NumPy implementation on identical inputs to within rel 1e-12 in fp64."

* `j_ion`   : |got - ref| <= 1e-12*|ref| + FLOOR,  FLOOR = 2*ulp(1)*I_B0/(2*pi*r^2).
              The floor is the reference's own rounding noise, not slack for the kernel: j_cex is proportional to
              `1 - exp(-r*n*sigma)` (plume.py:95-96), so ONE ulp of difference in that exp -- NumPy's SIMD exp is
              itself only faithful to +-1 ulp and differs between CPUs -- moves j_cex by ulp(1)/2 * I_B0/(2 pi r^2)
              in absolute terms, which exceeds 1e-12 relative whenever r*n*sigma < ~1e-4.
              tests report the fraction of elements that meet the PURE 1e-12 rule as well.
              The 1e-20 invalid fill, NaN positions and the invalid mask must match EXACTLY.
* `cos_div`, `T_c`, `V_cc` : rel 1e-12 (V_cc: relative to the largest term of cathode.py:34's sum, since the
              clamp at 0 makes a pure relative test meaningless when the terms cancel).
* `div_angle` = arccos(cos_div): arccos amplifies a relative error e in cos_div to e*cot(theta) in theta, so
              |got - ref| <= 1e-12*(|ref| + |cot(ref)|); for degenerate needle beams (theta < 1e-5) the comparison
              is made on cos_div instead (arccos is singular at 1).
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-12
ULP1 = np.finfo(np.float64).eps


def _same_nan(got, ref, name):
    gn, rn = np.isnan(got), np.isnan(ref)
    assert np.array_equal(gn, rn), f'{name}: NaN positions differ ({int(gn.sum())} vs {int(rn.sum())})'
    return ~rn


def check_j_ion(got, ref, i_b0, radii, invalid_ref, name='j_ion'):
    """got/ref: (..., A) or (..., A, R); i_b0 broadcastable to loop shape; returns fraction meeting pure rel 1e-12."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f'{name}: shape {got.shape} != {ref.shape}'
    radii = np.atleast_1d(np.asarray(radii, dtype=np.float64))
    single = radii.shape[0] == 1
    loop = ref.shape[:-1] if single else ref.shape[:-2]
    ok = _same_nan(got, ref, name)
    inv = np.broadcast_to(np.asarray(invalid_ref, dtype=bool), loop)
    # invalid rows: exact fill
    assert np.all(got[inv] == 1e-20), f'{name}: invalid rows are not exactly 1e-20'
    assert np.all(ref[inv] == 1e-20)
    ib = np.broadcast_to(np.abs(np.asarray(i_b0, dtype=np.float64)), loop)
    floor = 2 * ULP1 * ib[..., None] / (2 * np.pi * (radii ** 2 if not single else radii[0] ** 2))
    if single:
        floor = np.broadcast_to(floor, ref.shape)
    else:
        floor = np.broadcast_to(floor[..., None, :], ref.shape)
    with np.errstate(invalid='ignore'):
        err = np.abs(got - ref)
        lim = RTOL * np.abs(ref) + floor
        bad = ok & ~(err <= lim)
        pure = ok & (err <= RTOL * np.abs(ref))
    if bad.any():
        idx = np.unravel_index(np.argmax(np.where(ok, err / np.maximum(lim, 1e-300), 0)), ref.shape)
        raise AssertionError(f'{name}: {int(bad.sum())} of {ref.size} outside rel 1e-12 + floor; worst at {idx}: '
                             f'got {got[idx]!r} ref {ref[idx]!r} err {err[idx]:.3e} lim {lim[idx]:.3e}')
    return float(pure.sum()) / max(1, int(ok.sum()))


def check_rel(got, ref, name, scale=None, rtol=RTOL):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f'{name}: shape {got.shape} != {ref.shape}'
    ok = _same_nan(got, ref, name)
    ref_scale = np.abs(ref) if scale is None else np.maximum(np.abs(ref), np.abs(scale))
    with np.errstate(invalid='ignore'):
        err = np.abs(got - ref)
        bad = ok & ~(err <= rtol * ref_scale)
    if bad.any():
        i = np.argmax(np.where(ok, err / np.maximum(ref_scale, 1e-300), 0))
        idx = np.unravel_index(i, ref.shape)
        raise AssertionError(f'{name}: {int(bad.sum())} of {ref.size} outside rel {rtol}; worst at {idx}: '
                             f'got {got[idx]!r} ref {ref[idx]!r}')
    with np.errstate(invalid='ignore', divide='ignore'):
        r = np.where(ok & (ref_scale > 0), err / ref_scale, 0.0)
    return float(r.max()) if r.size else 0.0


def check_div_angle(got, ref, cos_got=None, cos_ref=None, name='div_angle'):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape
    ok = _same_nan(got, ref, name)
    needle = ok & (np.abs(ref) < 1e-5)
    reg = ok & ~needle
    with np.errstate(invalid='ignore', divide='ignore'):
        lim = RTOL * (np.abs(ref) + np.abs(np.cos(ref) / np.sin(ref)))
        err = np.abs(got - ref)
        bad = reg & ~(err <= lim)
    if bad.any():
        i = np.argmax(np.where(reg, err / lim, 0))
        idx = np.unravel_index(i, ref.shape)
        raise AssertionError(f'{name}: {int(bad.sum())} outside tolerance; worst at {idx}: got {got[idx]!r} ref {ref[idx]!r}')
    if needle.any():
        if cos_got is not None and cos_ref is not None:
            check_rel(np.asarray(cos_got)[needle], np.asarray(cos_ref)[needle], name + '(needle, via cos_div)')
        else:
            assert np.all(np.abs(got[needle] - ref[needle]) <= 2e-6), f'{name}: needle-beam samples differ'
    with np.errstate(invalid='ignore', divide='ignore'):
        r = np.where(reg, err / np.abs(ref), 0.0)
    return float(r.max()) if r.size else 0.0


def cathode_scale(inp, torr):
    """Largest term of cathode.py:34's sum (the natural scale of V_cc's rounding error)."""
    pb = np.asarray(inp['P_b'], dtype=np.float64) * torr
    pt = np.asarray(inp['P_T'], dtype=np.float64) * torr
    ps = np.asarray(inp['Pstar'], dtype=np.float64) * torr
    te = np.asarray(inp['T_e'], dtype=np.float64)
    with np.errstate(all='ignore'):
        t1 = np.abs(te * np.log(1 + pb / pt))
        t2 = np.abs(te / (pt + ps) * pb)
    return np.maximum(np.maximum(np.abs(np.asarray(inp['V_vac'], dtype=np.float64)), t1), t2)


def load_golden(path):
    import json
    g = np.load(path, allow_pickle=False)
    meta = json.loads(str(g['meta']))
    inputs = {k[3:]: g[k] for k in g.files if k.startswith('in_')}
    return g, meta, inputs
