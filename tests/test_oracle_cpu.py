"""CPU tests (no GPU): the oracle against the committed golden vectors and -- where /root/reference exists -- against
the UNMODIFIED reference.  This is what pins the oracle (the reference's own tests hold no
golden vectors; the committed files were minted from the reference by oracle/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_import
from oracle.ref_restated import (angle_grid, cathode_coupling_oracle, current_density_oracle, simpson_weights,
                                 _beam_integral)
from tests import parity

GOLDEN = sorted((Path(__file__).parent / 'golden').glob('*.npz'))


def test_golden_files_present():
    names = {p.stem for p in GOLDEN}
    assert {'ref91_spt100_n256', 'ref91_testrange_n64_r25', 'cfg1_spt100_n256_a100', 'cfg2_spt100_n128_a200',
            'cfg3_h9_n128_a256', 'cfg5_spt100_n64_a512', 'edge_a91', 'edge_a100', 'edge_a91_r3'} <= names


@pytest.mark.parametrize('path', GOLDEN, ids=[p.stem for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    """Same NumPy/SciPy/CPU as the generator -> bit-identical; other machines (NumPy's SIMD exp/log differ by an ulp
    between CPU families) -> the parity tolerances of tests/parity.py."""
    g, meta, inputs = parity.load_golden(path)
    radii = g['sweep_radius']
    with np.errstate(all='ignore'):
        out = current_density_oracle(inputs, radii if radii.shape[0] > 1 else float(radii[0]), meta['n_angles'],
                                     meta['torr_2_pa'], with_coords=False, return_internals=True)
        v = cathode_coupling_oracle(inputs, meta['torr_2_pa'])['V_cc']
    assert np.array_equal(out['_invalid'], g['invalid'])
    parity.check_j_ion(out['j_ion'], g['j_ion'], inputs['I_B0'], radii, g['invalid'])
    parity.check_rel(out['_cos_div'], g['cos_div'], 'cos_div')
    parity.check_rel(out['T_c'], g['T_c'], 'T_c')
    parity.check_div_angle(out['div_angle'], g['div_angle'], out['_cos_div'], g['cos_div'])
    parity.check_rel(v, g['V_cc'], 'V_cc', scale=parity.cathode_scale(inputs, meta['torr_2_pa']))
    assert out['_max_imag'] == 0.0 or np.isnan(out['_max_imag'])     # plume.py:109 never warns on these inputs


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference is only present in the build container')
@pytest.mark.parametrize('seed,n', [(1, 300), (2, 2000)])
def test_restatement_is_bit_identical_to_reference(seed, n):
    from hallthrusterpem_b200.synthetic import spt100_batch
    ref_plume, ref_cathode, torr = ref_import.load()
    b = spt100_batch(n, seed, c3_test_range=bool(seed % 2))
    radii = np.random.default_rng(seed).uniform(1.0, 1.2, 5)
    with np.errstate(all='ignore'):
        for sweep in (1.0, radii):
            r = ref_plume(dict(b), sweep)
            o = current_density_oracle(b, sweep, 91, torr)
            for key in ('j_ion', 'div_angle', 'T_c'):
                assert np.array_equal(r[key], o[key], equal_nan=True), key
            assert r['j_ion_coords'].shape == o['j_ion_coords'].shape and r['j_ion_coords'].dtype == object
            assert np.array_equal(r['j_ion_coords'].flat[0], o['j_ion_coords'].flat[0])
        assert np.array_equal(ref_cathode(dict(b))['V_cc'], cathode_coupling_oracle(b, torr)['V_cc'])


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference is only present in the build container')
def test_reference_sanity_values():
    """SURVEY.md appendix B known-answer values, from the unmodified reference."""
    ref_plume, ref_cathode, torr = ref_import.load()
    v = ref_cathode({'P_b': 10e-6, 'V_a': 300, 'T_e': 3, 'V_vac': 30, 'Pstar': 20e-6, 'P_T': 50e-6})['V_cc']
    assert abs(v[0] - 30.11839324) < 1e-8
    o = ref_plume({'P_b': 1e-5, 'c0': .1, 'c1': .7, 'c2': -8., 'c3': .2, 'c4': 1e20, 'c5': 1e16, 'sigma_cex': 55e-20,
                   'I_B0': 3, 'T': .08}, 1.0)
    assert abs(o['j_ion'][0, 0] - 23.5475585) < 1e-6 and abs(o['j_ion'][0, 90] - 0.036191975) < 1e-8
    assert abs(o['div_angle'][0] - 0.19785857) < 1e-8 and abs(o['T_c'][0] - 0.07843918) < 1e-8


def test_oracle_properties():
    """Known-answer facts of SURVEY.md section 8c that need no reference import."""
    from scipy.integrate import quad, simpson
    # the complex-erfi closed form is the beam integral 2 pi int exp(-(t/a)^2) sin t dt, and is exactly real
    for a in (0.05, 0.3, 1.0, 1.5707963, 5.0, 15.0):
        d = _beam_integral(np.array([a]))[0]
        q = 2 * np.pi * quad(lambda t: np.exp(-(t / a) ** 2) * np.sin(t), 0, np.pi / 2, epsabs=1e-15, epsrel=1e-14)[0]
        assert d.imag == 0.0 and abs(d.real / q - 1) < 1e-12
    # Simpson weight equivalence incl. the even-count tail
    for A in (91, 100, 200, 256, 512):
        th = angle_grid(A)
        y = np.cos(3 * th) + th ** 2
        assert abs(simpson_weights(th) @ y / simpson(y, x=th) - 1) < 1e-14
    # invalid-sample asymmetry: j_ion is masked, div_angle is not (plume.py:104-107 vs 117-127)
    o = current_density_oracle({'P_b': 1e-4, 'c0': .3, 'c1': .5, 'c2': -15., 'c3': .1, 'c4': 1e20, 'c5': 1e16,
                                'sigma_cex': 55e-20, 'I_B0': 3.}, 1.0, 91, 133.322, return_internals=True)
    assert o['_invalid'][0] and np.all(o['j_ion'] == 1e-20) and np.isfinite(o['div_angle'][0])
    # scalar inputs -> (1, 91), (1,), object coords of shape (1,)
    assert o['j_ion'].shape == (1, 91) and o['div_angle'].shape == (1,) and o['j_ion_coords'].shape == (1,)


@pytest.mark.skipif(not ref_import.available(), reason='reference sources not present (GPU box)')
def test_drop_in_signatures_match_the_reference():
    """Same callable names, same positional parameters with the same defaults; every extra parameter of the drop-in is
    keyword-only with a default, so any call that is valid for the reference is valid for the drop-in."""
    import inspect
    from hallthrusterpem_b200.models import cathode_coupling, current_density
    ref_plume, ref_cathode, _ = ref_import.load()
    for ours, ref in ((current_density, ref_plume), (cathode_coupling, ref_cathode)):
        assert ours.__name__ == ref.__name__
        po, pr = inspect.signature(ours).parameters, inspect.signature(ref).parameters
        assert list(po)[:len(pr)] == list(pr)
        for name, prm in pr.items():
            assert po[name].kind == prm.kind and po[name].default == prm.default
        for name in list(po)[len(pr):]:
            assert po[name].kind is inspect.Parameter.KEYWORD_ONLY and po[name].default is not inspect.Parameter.empty
