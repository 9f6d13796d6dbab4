"""Exporter of the path's outputs to the reference's comparison format (host side, no kernel).

Mirror of `hallmd.data.pem_to_xarray` (/root/reference/src/hallmd/data.py:239-279): one entry per operating condition
with the quantities this path produces -- cathode coupling voltage, (divergence-corrected) thrust and the ion current
density as an `(r, theta)` field -- plus the pass-through quantities of the external thruster solve (`I_d`, `u_ion`)
when the caller supplies them.  `xarray` and `pem_core.types` are not vendored with the reference; when `xarray` is
importable the values are `xarray.DataArray`s exactly as in the reference, otherwise a minimal stand-in with the same
`.values / .dims / .coords` attributes is used, so downstream code that only reads those keeps working.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

try:  # pragma: no cover - depends on the environment
    import xarray as _xr
except Exception:  # noqa: BLE001
    _xr = None


@dataclass
class FieldArray:
    """Stand-in for `xarray.DataArray` (values + named dims + coordinate arrays)."""
    values: np.ndarray
    dims: tuple = ()
    coords: dict = field(default_factory=dict)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)


def _data_array(values, coords=None, dims=None):
    values = np.asarray(values)
    if _xr is not None:
        return _xr.DataArray(values) if coords is None else _xr.DataArray(values, coords=coords, dims=dims)
    return FieldArray(values, tuple(dims or ()), dict(zip(dims or (), coords or ())))


def pem_to_xarray(operating_conditions: list[dict], outputs: dict, sweep_radii, use_corrected_thrust: bool = True) -> list[dict]:
    """data.py:239-279.  `outputs` is the merged output dict of the PEM chain (`V_cc`, `j_ion`, `j_ion_coords`, `T_c`
    from this package; `T`, `I_d`, `u_ion`, `u_ion_coords` from the external thruster model when present).
    Returns a list of `{'operating_condition': ..., 'data': {name: {'val': DataArray, 'unit': str}}}` entries."""
    r = np.atleast_1d(np.asarray(sweep_radii, dtype=np.float64))
    j_all = np.atleast_3d(np.asarray(outputs['j_ion']))                       # data.py:265: (n, A, R)
    entries = []
    for i, opcond in enumerate(operating_conditions):
        data: dict[str, dict] = {}
        if use_corrected_thrust:
            # with several radii there are several corrected thrusts; the reference keeps the last (radii sorted), :250
            data['thrust'] = {'val': _data_array(np.atleast_1d(outputs['T_c'][i])[-1]), 'unit': 'N'}
        elif 'T' in outputs:
            data['thrust'] = {'val': _data_array(outputs['T'][i]), 'unit': 'N'}
        if 'I_d' in outputs:
            data['discharge current'] = {'val': _data_array(outputs['I_d'][i]), 'unit': 'A'}
        data['cathode coupling voltage'] = {'val': _data_array(outputs['V_cc'][i]), 'unit': 'V'}
        if 'u_ion' in outputs:
            z = outputs['u_ion_coords'][i]
            data['ion velocity'] = {'val': _data_array(outputs['u_ion'][i], coords=[z], dims=['z']), 'unit': 'm/s'}
        theta = outputs['j_ion_coords'][i]
        jion = j_all[i, :, :].T                                               # (R, A), data.py:265
        data['ion current density'] = {'val': _data_array(jion, coords=[r, theta], dims=['r', 'theta']), 'unit': 'A/m^2'}
        entries.append({'operating_condition': opcond, 'data': data})
    return entries
