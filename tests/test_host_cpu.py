"""CPU tests (no GPU) of the host-side logic: quadrature weights, beam-integral table, input marshalling,
C-ABI library surface, sharding, and the multi-rank merge over gloo."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_simpson_weights_match_scipy():
    from scipy.integrate import simpson
    from hallthrusterpem_b200.quadrature import angle_grid, fused_weights, simpson_weights
    rng = np.random.default_rng(3)
    for A in (2, 3, 4, 5, 91, 100, 200, 256, 512, 513):
        x = angle_grid(A)
        W = simpson(np.eye(A), x=x, axis=-1)
        assert np.max(np.abs(simpson_weights(x) - W)) <= 4e-16 * np.max(np.abs(W))
        xr = np.sort(rng.uniform(0, 2, A))
        assert np.max(np.abs(simpson_weights(xr) - simpson(np.eye(A), x=xr, axis=-1))) < 1e-15
        if A >= 3:
            wd, wn = fused_weights(x)
            f = np.exp(-(x / 0.4) ** 2) + 0.01
            flipped = f[::-1]
            assert abs(wd @ f / simpson(flipped * np.cos(x), x=x) - 1) < 1e-14
            assert abs(wn @ f / simpson(flipped * np.cos(x) * np.sin(x), x=x) - 1) < 1e-14


def _device_table():
    txt = (ROOT / 'hallthrusterpem_b200' / 'csrc' / 'hpem_dtable.inc').read_text()
    rows = re.findall(r'\{([^{}]+)\},', txt)
    tab = np.array([[float.fromhex(v.strip()) for v in r.split(',')] for r in rows])
    m = int(re.search(r'#define HPEM_DTAB_M (\d+)', txt).group(1))
    umax = float(re.search(r'#define HPEM_DTAB_UMAX ([0-9.]+)', txt).group(1))
    return tab, m, umax


def test_beam_integral_table_matches_reference_closed_form():
    """NumPy emulation of the table lookup in csrc/hpem_device.cuh::beam_amplitude (same Horner order) vs the oracle's complex-erfi form
    (plume.py:64-85) over the whole domain the reference can evaluate."""
    from oracle.ref_restated import _beam_integral
    tab, m, umax = _device_table()
    assert tab.shape == (64, 11) and m == 64

    def d_tab(a):
        aa = np.abs(a)
        v = aa / (aa + 2.0) * (m / umax)
        idx = np.minimum(v.astype(int), m - 1)
        t = 2 * (v - idx) - 1
        q = tab[idx, 0].copy()
        for j in range(1, tab.shape[1]):
            q = q * t + tab[idx, j]
        return q * ((2 * np.pi) * aa * aa / (aa * aa + 2.0))

    rng = np.random.default_rng(1)
    a = np.concatenate([10 ** rng.uniform(-6, np.log10(53.28), 100000), rng.uniform(0.01, 16, 100000)])
    with np.errstate(all='ignore'):
        ref = _beam_integral(a).real
    rel = np.abs(d_tab(a) / ref - 1)
    assert rel.max() < 5e-14 and np.sqrt(np.mean(rel ** 2)) < 3e-15      # the reference itself is only good to ~1e-14
    assert np.allclose(d_tab(-a[:100]), d_tab(a[:100]), rtol=0, atol=0)   # even in alpha, like the reference
    mp = pytest.importorskip('mpmath')
    mp.mp.dps = 40
    for x in (1e-4, 0.021, 0.2, 0.9, 1.5707963, 7.0, 30.0, 53.0):
        z = mp.mpc(x / 2, mp.pi / (2 * x))
        truth = mp.pi ** mp.mpf(1.5) / 2 * x * mp.exp(-(mp.mpf(x) / 2) ** 2) * (2 * mp.erfi(mp.mpf(x) / 2) - 2 * mp.re(mp.erfi(z)))
        assert abs(d_tab(np.array([x]))[0] / float(truth) - 1) < 1e-15


def test_erfi_overflow_threshold_constant():
    """csrc/hpem_device.cuh::kExpOverflow reproduces where the reference's normalisation turns non-finite."""
    from oracle.ref_restated import _beam_integral
    k = float.fromhex('0x1.62e42fefa39efp+9')
    lo, hi = 53.28349511409265, 53.28349511409266
    assert (lo / 2) ** 2 <= k < (hi / 2) ** 2
    with np.errstate(all='ignore'):
        v = _beam_integral(np.array([lo, hi]))
    assert np.isfinite(v[0].real) and not np.isfinite(v[1].real)


def test_c_abi_library_exports_declared_symbols():
    """The shared library loads without a GPU and exports every function include/hpem.h declares."""
    from hallthrusterpem_b200 import _lib
    path = _lib.build_library()
    header = (ROOT / 'include' / 'hpem.h').read_text()
    declared = set(re.findall(r'^\s*(?:const\s+)?[A-Za-z_0-9]+\s*\*?\s*(hpem_[a-z_0-9]+)\s*\(', header, flags=re.M))
    assert {'hpem_abi_version', 'hpem_source_hash', 'hpem_last_error', 'hpem_grid_create', 'hpem_grid_destroy', 'hpem_grid_is_uniform',
            'hpem_eval', 'hpem_eval_host', 'hpem_launch_count', 'hpem_moments_layout_query',
            'hpem_moments_accumulate', 'hpem_sample_inputs', 'hpem_moments_accumulate_sampled', 'hpem_moments_merge', 'hpem_quadrature_table_eval',
            'hpem_measurements_create', 'hpem_measurements_destroy', 'hpem_loglike', 'hpem_logsumexp',
            'hpem_basis_create', 'hpem_basis_destroy', 'hpem_compress', 'hpem_compress_field', 'hpem_reconstruct'} == declared
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    lib = ctypes.CDLL(str(path))
    for name in declared:
        assert hasattr(lib, name), name
    lib.hpem_abi_version.restype = ctypes.c_int
    assert lib.hpem_abi_version() == _lib.ABI_VERSION == 3
    lib.hpem_source_hash.restype = ctypes.c_char_p
    assert lib.hpem_source_hash().decode() == _lib.source_hash() == _lib.embedded_hash(path)
    assert ctypes.sizeof(_lib.HpemInputs) == 15 * 8 * 2 and ctypes.sizeof(_lib.HpemOutputs) == 6 * 8
    assert ctypes.sizeof(_lib.HpemMomentsSpec) == 48 and ctypes.sizeof(_lib.HpemMomentsLayout) == 48
    assert ctypes.sizeof(_lib.HpemPrior) == 24


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from hallthrusterpem_b200.models import cathode_coupling, current_density
    scal = {'P_b': 1e-5, 'c0': .1, 'c1': .7, 'c2': -8., 'c3': .2, 'c4': 1e20, 'c5': 1e16, 'sigma_cex': 55e-20, 'I_B0': 3}
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        current_density(scal)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        cathode_coupling({'P_b': 10e-6, 'V_a': 300, 'T_e': 3, 'V_vac': 30, 'Pstar': 20e-6, 'P_T': 50e-6})


def test_product_does_not_import_the_oracle():
    for py in (ROOT / 'hallthrusterpem_b200').rglob('*.py'):
        assert not re.search(r'^\s*(from|import)\s+oracle\b', py.read_text(), flags=re.M), py


def test_input_marshalling_broadcasts_like_numpy():
    from hallthrusterpem_b200 import _lib
    from hallthrusterpem_b200.engine import _Batch
    rng = np.random.default_rng(0)
    inputs = {'P_b': rng.random((4, 3)), 'c0': 0.3, 'c1': np.float64(0.5), 'c2': rng.random((4, 1)), 'c3': np.array([0.2]),
              'c4': 1e20, 'c5': 1e16, 'sigma_cex': 55e-20, 'I_B0': rng.random(3), 'T': None, 'extra': 'ignored'}
    b = _Batch(inputs, _lib.PLUME_INPUTS, optional=('T',))
    assert b.out_shape == (4, 3) and b.n == 12 and not b.on_device and 'T' not in b.present
    k = _lib.INPUT_NAMES.index
    assert b.struct.ptr[k('c0')] is None and b.struct.scalar[k('c0')] == 0.3
    assert b.struct.ptr[k('c3')] is None and b.struct.scalar[k('c3')] == 0.2          # size-1 array = scalar
    c2 = np.ctypeslib.as_array(ctypes.cast(b.struct.ptr[k('c2')], ctypes.POINTER(ctypes.c_double)), shape=(12,))
    assert np.array_equal(c2.reshape(4, 3), np.broadcast_to(inputs['c2'], (4, 3)))
    ib = np.ctypeslib.as_array(ctypes.cast(b.struct.ptr[k('I_B0')], ctypes.POINTER(ctypes.c_double)), shape=(12,))
    assert np.array_equal(ib.reshape(4, 3), np.broadcast_to(inputs['I_B0'], (4, 3)))
    scal = _Batch({n: 1.0 for n in _lib.PLUME_INPUTS}, _lib.PLUME_INPUTS)
    assert scal.out_shape == (1,) and scal.n == 1                                       # np.atleast_1d (plume.py:59)
    with pytest.raises(KeyError):
        _Batch({'P_b': 1.0}, _lib.PLUME_INPUTS)
    with pytest.raises(ValueError):
        _Batch(dict({n: 1.0 for n in _lib.PLUME_INPUTS}, P_b=np.ones(3), c0=np.ones(4)), _lib.PLUME_INPUTS)


def test_synthetic_batches_and_sharding():
    from hallthrusterpem_b200.synthetic import ALL_KEYS, h9_sweep_batch, shard_bounds, spt100_batch
    b = spt100_batch(1000, 7)
    assert tuple(b) == ALL_KEYS and all(v.shape == (1000,) and v.dtype == np.float64 for v in b.values())
    assert np.array_equal(b['c1'], spt100_batch(1000, 7)['c1']) and not np.array_equal(b['c1'], spt100_batch(1000, 8)['c1'])
    assert 1e-8 <= b['P_b'].min() and b['P_b'].max() <= 1e-4 and 0.1 <= b['c1'].min() and b['c1'].max() <= 0.9
    h = h9_sweep_batch(101)
    assert h['P_b'][0] == 0.0 and h['P_b'][-1] == 1e-4
    for n, w in ((10, 3), (1_000_000, 8), (7, 8), (100_000_001, 8), (8191, 8), (8192, 8)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b_[0] for a, b_ in zip(spans, spans[1:]))
        big = n // w >= 1024
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= (128 if big else 1)
        assert not big or all(lo % 64 == 0 for lo, _ in spans)      # rows keep their position modulo 4 / 64 across shards


def test_philox_known_answers_and_uniforms():
    """The NumPy statement of the sampler's stream reproduces the published Philox4x32-10 known-answer vectors
    (Random123 kat_vectors); the device code is the same function compiled by nvcc (tests/test_sampler_gpu.py)."""
    from hallthrusterpem_b200.sampler import SPT100_PRIORS
    from oracle.sampler_oracle import apply_priors_numpy, philox4x32_10, philox_uniforms
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(x[0]) for x in got) == want
    u = philox_uniforms(123, 10, 50000)
    assert u.shape == (50000, 15) and 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 2e-3
    assert np.array_equal(u * 2.0 ** 42, np.floor(u * 2.0 ** 42)) and abs(np.corrcoef(u[:, 0], u[:, 2])[0, 1]) < 0.02   # 42-bit draws
    assert np.array_equal(philox_uniforms(123, 1010, 100), u[1000:1100])       # index-addressed: shards line up
    x = apply_priors_numpy(u, SPT100_PRIORS)
    assert 1e-8 <= x['P_b'].min() and x['P_b'].max() <= 1e-4 and 200 <= x['V_a'].min() and x['V_a'].max() <= 400
    assert abs(np.corrcoef(x['c0'], x['c1'])[0, 1]) < 0.02


@pytest.mark.parametrize('n_angles', [2, 3, 17, 91, 200, 256, 512])
def test_quadrature_table_matches_long_double_sums(n_angles):
    """The tabulated Simpson sums the reduce-only kernel uses (csrc/hpem_qtable.cuh), evaluated on the host with the
    kernel's arithmetic, against long-double sums of plume.py:117-123's integrands over the whole range of x = (h/alpha)^2
    (clamped bins, needle beams and x = 0 included): <= 5e-16 relative, the level of a double-precision sum."""
    from hallthrusterpem_b200 import _lib
    from hallthrusterpem_b200.quadrature import angle_grid, fused_weights
    lib = ctypes.CDLL(str(_lib.build_library()))
    dptr = ctypes.POINTER(ctypes.c_double)
    lib.hpem_quadrature_table_eval.argtypes = [ctypes.c_int, dptr, dptr, ctypes.c_int64, dptr, dptr, dptr]
    wd, wn = fused_weights(angle_grid(n_angles))
    rng = np.random.default_rng(n_angles)
    x = np.concatenate([2.0 ** rng.uniform(-75, 12, 6000), 2.0 ** rng.uniform(-12, 2, 6000),
                        [0.0, 1e-300, 2.0 ** -48, 15.999999, 16.0, 700.0, 745.0, 1e30]])
    nd, nn = np.empty_like(x), np.empty_like(x)
    assert lib.hpem_quadrature_table_eval(n_angles, wd.ctypes.data_as(dptr), wn.ctypes.data_as(dptr), len(x),
                                          x.ctypes.data_as(dptr), nd.ctypes.data_as(dptr), nn.ctypes.data_as(dptr)) == 0
    i2 = np.arange(n_angles, dtype=np.longdouble) ** 2
    e = np.exp(-x.astype(np.longdouble)[:, None] * i2[None, :])
    rd, rn = (e * wd.astype(np.longdouble)).sum(axis=1), (e * wn.astype(np.longdouble)).sum(axis=1)
    assert np.all(np.abs(nd - rd) <= 5e-16 * np.abs(rd)) and np.all(np.abs(nn - rn) <= 5e-16 * np.abs(rn))
    # misuse is reported through the status code + hpem_last_error, like every other entry point
    lib.hpem_last_error.restype = ctypes.c_char_p
    assert lib.hpem_quadrature_table_eval(1, wd.ctypes.data_as(dptr), wn.ctypes.data_as(dptr), 0, None, None, None) != 0
    assert b'n_angles' in lib.hpem_last_error()
    assert lib.hpem_quadrature_table_eval(n_angles, None, wn.ctypes.data_as(dptr), 0, None, None, None) != 0
    assert lib.hpem_quadrature_table_eval(n_angles, wd.ctypes.data_as(dptr), wn.ctypes.data_as(dptr), 0, None, None, None) == 0


def _gloo_worker(rank, world, port, n, n_angles, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from hallthrusterpem_b200.mc import HistogramSpec, Layout, merge_across_ranks
    from hallthrusterpem_b200.synthetic import shard_bounds, spt100_batch
    from oracle.moments_oracle import packed_moments
    from oracle.ref_restated import cathode_coupling_oracle, current_density_oracle
    layout = Layout(n_angles, HistogramSpec(angle_stride=8, sub_bits=2))
    full = spt100_batch(n, 31)
    lo, hi = shard_bounds(n, world, rank)
    mine = {k: v[lo:hi] for k, v in full.items()}
    with np.errstate(all='ignore'):
        o = current_density_oracle(mine, 1.0, n_angles, with_coords=False, return_internals=True)
        v = cathode_coupling_oracle(mine)['V_cc']
    sums, minmax = packed_moments(layout, o['j_ion'], v, o['div_angle'], o['T_c'], o['_invalid'])
    packed = torch.from_numpy(np.concatenate([sums, minmax]))
    merge_across_ranks(layout, packed)                 # ONE all-gather + the fixed-order merge, on every rank
    q.put((rank, packed.numpy().copy()))
    dist.destroy_process_group()


def test_two_rank_merge_over_gloo_equals_single_rank():
    """world_size 2 on CPU: shard -> per-rank packed moments -> merge (one all-gather + pairwise merge in rank order) ==
    unsharded result, and both ranks hold the same bits."""
    import torch.multiprocessing as mp
    from hallthrusterpem_b200.mc import HistogramSpec, Layout, MomentsResult
    from hallthrusterpem_b200.synthetic import spt100_batch
    from oracle.moments_oracle import packed_moments
    from oracle.ref_restated import cathode_coupling_oracle, current_density_oracle
    n, n_angles, world = 601, 100, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (np.random.default_rng().integers(0, 2000))
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, int(port), n, n_angles, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    layout = Layout(n_angles, HistogramSpec(angle_stride=8, sub_bits=2))
    assert np.array_equal(got[0], got[1]), 'ranks disagree after the merge'
    sums, minmax = got[0][:layout.n_sums], got[0][layout.n_sums:]
    full = spt100_batch(n, 31)
    with np.errstate(all='ignore'):
        o = current_density_oracle(full, 1.0, n_angles, with_coords=False, return_internals=True)
        v = cathode_coupling_oracle(full)['V_cc']
    ref_s, ref_m = packed_moments(layout, o['j_ion'], v, o['div_angle'], o['T_c'], o['_invalid'])
    np.testing.assert_allclose(sums, ref_s, rtol=1e-12)
    assert np.array_equal(sums[layout.off_hist:], ref_s[layout.off_hist:]) and np.array_equal(minmax, ref_m)
    np.testing.assert_allclose(res_var(layout, sums), o['j_ion'].var(axis=0), rtol=1e-12)
    res = MomentsResult(layout, sums, minmax)
    assert res.n_samples == n and abs(res.scalar('V_cc')['mean'] - v.mean()) < 1e-12
    np.testing.assert_allclose(res.j_mean, o['j_ion'].mean(axis=0), rtol=1e-12)
    p = res.j_percentile([5, 50, 95])
    ref_p = np.percentile(o['j_ion'][:, layout.hist_angle_index], [5, 50, 95], axis=0)
    assert np.all(np.abs(p / ref_p - 1) < 0.10)         # 601 samples, 4 bins per octave (25 % wide), interpolated inside the bin


def res_var(layout, sums):
    from hallthrusterpem_b200.mc import MomentsResult
    return MomentsResult(layout, sums, np.full(6, -np.inf)).j_var


def test_percentile_interpolation_against_numpy():
    """mc.MomentsResult.j_percentile on synthetic log-normal populations: the in-bin interpolation brings the 8-per-octave
    histogram (bins 9-12 % wide) within 2 % of np.percentile (tests/test_plume.py:50-52 takes exactly those percentiles)."""
    from hallthrusterpem_b200.mc import HistogramSpec, Layout, MomentsResult
    from oracle.moments_oracle import hist_bins
    rng = np.random.default_rng(3)
    spec = HistogramSpec(angle_stride=1, sub_bits=3)
    A, n = 4, 200_000
    layout = Layout(A, spec)
    j = np.exp(rng.normal([[-3.0, 0.0, 2.0, 5.0]], [[0.3, 1.0, 2.0, 0.6]], (n, A)))
    sums = np.zeros(layout.n_sums)
    sums[0] = n
    for a in range(A):
        sums[layout.off_hist + a * layout.n_bins: layout.off_hist + (a + 1) * layout.n_bins] = np.bincount(
            hist_bins(j[:, a], spec.sub_bits, spec.min_exp2, spec.max_exp2), minlength=layout.n_bins)
    res = MomentsResult(layout, sums, np.full(6, -np.inf))
    q = [1, 5, 25, 50, 75, 95, 99]
    got, ref = res.j_percentile(q), np.percentile(j, q, axis=0)
    assert np.all(np.abs(got / ref - 1) < 0.02), np.abs(got / ref - 1).max()


def test_merge_packed_host_is_order_stable_and_matches_single_pass():
    """Chan merge of shard vectors == the vector of the whole population (sums, centred second moments, histograms)."""
    from hallthrusterpem_b200.mc import HistogramSpec, Layout, merge_packed_host
    from oracle.moments_oracle import packed_moments
    rng = np.random.default_rng(5)
    layout = Layout(24, HistogramSpec(angle_stride=8, sub_bits=2))
    n = 5000
    j = np.exp(rng.normal(0, 1.5, (n, 24))) + 1e3           # large mean / small spread: the raw-sum formula would lose digits
    v, d, t = rng.normal(30, 1e-3, n), rng.uniform(0.1, 1.0, n), rng.normal(0.08, 1e-6, n)
    inv = np.zeros(n, bool)
    parts = []
    for lo, hi in ((0, 64), (64, 1000), (1000, 1000), (1000, n)):
        s, m = packed_moments(layout, j[lo:hi], v[lo:hi], d[lo:hi], t[lo:hi], inv[lo:hi])
        parts.append(np.concatenate([s, m]))
    merged = merge_packed_host(layout, np.stack(parts))
    ref_s, ref_m = packed_moments(layout, j, v, d, t, inv)
    np.testing.assert_allclose(merged[:layout.n_sums], ref_s, rtol=1e-11)
    assert np.array_equal(merged[layout.n_sums:], ref_m)
    assert abs(merged[5] / n - v.var()) < 1e-12 * v.var() * 1e3      # var 1e-6 on a mean of 30: centred merge keeps it


def test_pem_to_xarray_layout_matches_reference_text():
    """data.py:239-279: (r, theta) field orientation, last-radius thrust, coords taken from j_ion_coords."""
    from hallthrusterpem_b200.export import pem_to_xarray
    n, A, R = 3, 7, 2
    rng = np.random.default_rng(0)
    alpha = np.linspace(0, np.pi / 2, A)
    cell = np.empty((), dtype=object)
    cell[()] = alpha
    out = {'j_ion': rng.uniform(1, 2, (n, A, R)), 'j_ion_coords': np.broadcast_to(cell, (n,)), 'T_c': rng.uniform(0, 1, (n, R)),
           'V_cc': rng.uniform(0, 50, n), 'T': rng.uniform(0, 1, n)}
    ops = [{'P_b': 1e-5 * k} for k in range(n)]
    entries = pem_to_xarray(ops, out, [1.0, 1.5])
    assert len(entries) == n and entries[1]['operating_condition'] is ops[1]
    f = entries[2]['data']['ion current density']
    assert f['unit'] == 'A/m^2' and tuple(f['val'].dims) == ('r', 'theta')
    assert np.array_equal(np.asarray(f['val'].values), out['j_ion'][2].T)
    assert np.array_equal(np.asarray(f['val'].coords['theta']), alpha) and np.array_equal(np.asarray(f['val'].coords['r']), [1.0, 1.5])
    assert float(np.asarray(entries[0]['data']['thrust']['val'].values)) == out['T_c'][0, -1]
    assert float(np.asarray(entries[0]['data']['cathode coupling voltage']['val'].values)) == out['V_cc'][0]
    single = pem_to_xarray(ops, {**out, 'j_ion': out['j_ion'][..., 0], 'T_c': out['T_c'][:, 0]}, 1.0, use_corrected_thrust=False)
    assert np.asarray(single[0]['data']['ion current density']['val'].values).shape == (1, A)
    assert float(np.asarray(single[1]['data']['thrust']['val'].values)) == out['T'][1]


def test_pinned_output_pool_reuses_buffers(monkeypatch):
    """engine._PinnedPool: a buffer returns to the free-list only when its last NumPy view dies; the usual loop
    `out = f(x)` (previous result alive during the next call) settles on two buffers instead of pinning fresh memory."""
    import gc
    import torch
    from hallthrusterpem_b200 import engine
    real_empty = torch.empty
    allocations = []

    def fake_empty(n, dtype=None, pin_memory=False):
        assert pin_memory
        allocations.append(n)
        return real_empty(n, dtype=dtype)

    monkeypatch.setattr(torch, 'empty', fake_empty)
    pool = engine._PinnedPool(max_cached_bytes=10 * 1000 * 91 * 8)
    a = pool.take((1000, 91), np.float64)
    assert a.shape == (1000, 91) and a.dtype == np.float64 and a.flags.c_contiguous
    view = a[10:20]
    del a
    gc.collect()
    assert pool._cached == 0                                   # a view is still alive: the buffer must not be recycled
    pool.take((1000, 91), np.float64)
    del view
    gc.collect()
    assert len(allocations) == 2
    out = None
    for _ in range(10):
        out = pool.take((1000, 91), np.float64)                # noqa: F841  previous result alive while the next is allocated
    assert len(allocations) <= 4
    m = pool.take((7, 3), np.uint8)
    assert m.shape == (7, 3) and m.dtype == np.uint8


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle timed on the host cores) runs without a GPU and prints ONE JSON line with the
    keys the driver reads; under torchrun only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    cmd = [sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK='0', WORLD_SIZE='1'))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'evals/s' and d['higher_is_better'] is True and d['value'] > 1e6
    assert d['metric'] == 'plume+cathode fp64 sample x angle evals/s' and d['config']['n_angles'] == 200
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK='1', WORLD_SIZE='2'))
    assert res.returncode == 0 and not [ln for ln in res.stdout.splitlines() if ln.startswith('{')]


def test_empty_batch_raises_like_the_reference():
    """plume.py:122: scipy's simpson cannot handle a (0, A, R) integrand, so the reference raises ValueError for an empty
    sample batch; the drop-in raises the same error class (before touching the GPU)."""
    from hallthrusterpem_b200.models import current_density
    from oracle.ref_restated import current_density_oracle
    empty = {k: np.zeros(0) for k in ('P_b', 'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0')}
    with pytest.raises(ValueError):
        current_density_oracle(empty, 1.0, 91)
    with pytest.raises(ValueError):
        current_density(empty)


def test_branch_free_elementary_functions_meet_their_ulp_bounds(tmp_path):
    """csrc/hpem_fastmath.cuh compiled for the HOST (same source the kernels inline; the hardware reciprocal / rsqrt seeds are
    emulated by 20-bit truncations) against x87 long double libm: exp < 1 ulp and correctly rounded for |x| << 1, log < 1 ulp,
    acos < 1.25 ulp, div / sqrt correctly rounded, special values (tools/fastmath_check.cpp)."""
    import shutil
    import subprocess
    gxx = shutil.which('g++')
    if gxx is None:
        pytest.skip('no g++')
    exe = tmp_path / 'fastmath_check'
    subprocess.run([gxx, '-O2', '-mfma', '-ffp-contract=off', '-DHPEM_CHECK_N=200000', '-o', str(exe),
                    str(ROOT / 'tools' / 'fastmath_check.cpp')], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith('ok'), res.stdout
