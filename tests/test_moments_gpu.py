"""GPU tests of the reduce-only Monte-Carlo pass (K2) against the materialising path and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_moments(batch, n_angles, hist, chunks=1, want_thrust=True):
    import torch
    from hallthrusterpem_b200.mc import MonteCarloMoments
    mc = MonteCarloMoments(n_angles=n_angles, hist=hist, device=0, torr=133.322, want_thrust=want_thrust)
    n = len(batch['P_b'])
    bounds = np.linspace(0, n, chunks + 1).astype(int)
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        mc.accumulate({k: torch.as_tensor(v[lo:hi], device='cuda:0') for k, v in batch.items()})
    torch.cuda.synchronize()
    return mc


@pytest.mark.parametrize('n,n_angles,chunks', [(5000, 91, 1), (20000, 200, 3), (3333, 256, 2), (1000, 512, 1), (77, 17, 1)])
def test_moments_match_materialised_path(n, n_angles, chunks, cuda_device):
    from hallthrusterpem_b200.mc import HistogramSpec
    from hallthrusterpem_b200.models import plume_cathode
    from hallthrusterpem_b200.synthetic import spt100_batch
    from oracle.moments_oracle import packed_moments
    hist = HistogramSpec(angle_stride=8, sub_bits=3)
    b = spt100_batch(n, 4242 + n)
    mc = _run_moments(b, n_angles, hist, chunks)
    res = mc.result()
    # K2 shares the one-lane recurrence with K1u, so the histograms (exact integer counts) must agree with K1u's j_ion
    out = plume_cathode(b, 1.0, n_angles=n_angles, torr_2_pa=133.322, extras=True, lanes1=True)
    sums, minmax = packed_moments(mc.layout, out['j_ion'], out['V_cc'], out['div_angle'], out['T_c'], out['invalid'])
    L = mc.layout
    assert np.array_equal(res.sums[:3], sums[:3])
    assert np.array_equal(res.sums[[3, 6, 9]], sums[[3, 6, 9]])
    # sums to rel 1e-12; centred second moments (Chan merges on the device, two-pass NumPy here) to 1e-11
    np.testing.assert_allclose(res.sums[[4, 7, 10]], sums[[4, 7, 10]], rtol=1e-12)
    np.testing.assert_allclose(res.sums[[5, 8, 11]], sums[[5, 8, 11]], rtol=1e-11)
    np.testing.assert_allclose(res.sums[L.off_angle_sum:L.off_angle_sumsq], sums[L.off_angle_sum:L.off_angle_sumsq], rtol=1e-12)
    np.testing.assert_allclose(res.sums[L.off_angle_sumsq:L.off_hist], sums[L.off_angle_sumsq:L.off_hist], rtol=1e-11)
    assert np.array_equal(res.sums[L.off_hist:], sums[L.off_hist:]), 'histogram counts differ'
    np.testing.assert_allclose(res.minmax, minmax, rtol=1e-12)     # (K1u's quad-row mode sums num/den in another order: arccos amplifies)
    assert res.histograms.sum(axis=1).tolist() == [n] * L.n_hist_angles
    # decoded statistics against NumPy on the materialised arrays
    np.testing.assert_allclose(res.j_mean, out['j_ion'].mean(axis=0), rtol=1e-12)
    assert abs(res.scalar('div_angle')['mean'] - out['div_angle'].mean()) < 1e-12
    np.testing.assert_allclose(res.j_var, out['j_ion'].var(axis=0), rtol=1e-11)
    assert abs(res.scalar('V_cc')['var'] - out['V_cc'].var()) < 1e-11 * out['V_cc'].var()
    # the percentiles the reference's consumers take (tests/test_plume.py:50-52), interpolated inside the 8-per-octave bins
    pq = res.j_percentile([5, 50, 95])
    ref = np.percentile(out['j_ion'][:, L.hist_angle_index], [5, 50, 95], axis=0)
    assert np.all(np.abs(pq / ref - 1) < (0.02 if n >= 5000 else 0.12)), np.abs(pq / ref - 1).max()


@pytest.mark.parametrize('n_angles,stride,sub_bits', [(100, 4, 2), (224, 8, 3), (288, 8, 3)])
def test_moments_against_oracle_with_edge_cases(n_angles, stride, sub_bits, cuda_device):
    """Edge batch (invalid samples, NaN rows, late-invalid rows) + random samples, checked against the CPU oracle: the
    two-samples-per-thread kernel (100 angles), the three-samples one (224) and its restart variant (288)."""
    from hallthrusterpem_b200.mc import HistogramSpec
    from hallthrusterpem_b200.synthetic import spt100_batch
    from oracle.make_golden import edge_batch
    from oracle.moments_oracle import packed_moments
    from oracle.ref_restated import cathode_coupling_oracle, current_density_oracle
    e, r = edge_batch(), spt100_batch(500, 99)
    b = {k: np.concatenate([e[k], r[k]]) for k in e}
    hist = HistogramSpec(angle_stride=stride, sub_bits=sub_bits, min_exp2=-70, max_exp2=20)
    mc = _run_moments(b, n_angles, hist)
    res = mc.result()
    with np.errstate(all='ignore'):
        ref = current_density_oracle(b, 1.0, n_angles, 133.322, with_coords=False, return_internals=True)
        v = cathode_coupling_oracle(b, 133.322)['V_cc']
    sums, minmax = packed_moments(mc.layout, ref['j_ion'], v, ref['div_angle'], ref['T_c'], ref['_invalid'])
    L = mc.layout
    assert np.array_equal(res.sums[:3], sums[:3]), (res.sums[:3], sums[:3])
    assert np.array_equal(res.sums[[3, 6, 9]], sums[[3, 6, 9]])
    np.testing.assert_allclose(res.sums[3:12], sums[3:12], rtol=1e-10)
    np.testing.assert_allclose(res.sums[L.off_angle_sum:L.off_hist], sums[L.off_angle_sum:L.off_hist], rtol=1e-10)
    # histogram counts may differ from the oracle only where a value sits within rounding of a bin edge
    diff = np.abs(res.sums[L.off_hist:] - sums[L.off_hist:]).sum()
    assert diff <= 2 * 4 * (1 << (sub_bits - 2)), diff
    # (-min, max) of V_cc, div_angle, T_c; div_angle by the parity rule of tests/parity.py: arccos amplifies a relative error
    # in cos_div by cot(theta) (the edge batch holds needle beams with theta ~ 8e-3 rad)
    np.testing.assert_allclose(res.minmax[[0, 1, 4, 5]], minmax[[0, 1, 4, 5]], rtol=1e-12)
    th = np.abs(minmax[2:4])
    assert np.all(np.abs(res.minmax[2:4] - minmax[2:4]) <= 1e-12 * (th + np.abs(np.cos(th) / np.sin(th))))


def test_moments_without_histograms_and_thrust(cuda_device):
    from hallthrusterpem_b200.mc import HistogramSpec
    from hallthrusterpem_b200.synthetic import spt100_batch
    b = spt100_batch(1000, 5, with_thrust=False)
    mc = _run_moments(b, 91, HistogramSpec(angle_stride=0), want_thrust=False)
    res = mc.result()
    assert res.n_samples == 1000 and res.layout.n_sums == 12 + 2 * 91
    assert res.scalar('T_c')['n'] == 0 and res.scalar('V_cc')['n'] == 1000


def test_merge_kernel_matches_host_merge_and_sharded_sampling_is_exact(cuda_device):
    """(i) hpem_moments_merge == the NumPy statement of the same pairwise update; (ii) four shards of one sampled index
    range merged == the unsharded pass: counts and histograms bit for bit (index-addressed sampler), sums to 1e-12."""
    import torch
    from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments, merge_packed, merge_packed_host
    from hallthrusterpem_b200.synthetic import shard_bounds
    n, A, seed = 300_000, 91, 11
    whole = MonteCarloMoments(n_angles=A, hist=HistogramSpec(), device=0, torr=133.322)
    whole.accumulate_sampled(n, seed, 1000)
    parts = []
    for r in range(4):
        lo, hi = shard_bounds(n, 4, r)
        m = MonteCarloMoments(n_angles=A, hist=HistogramSpec(), device=0, torr=133.322)
        m.accumulate_sampled(hi - lo, seed, 1000 + lo)
        parts.append(m.packed.clone())
    parts = torch.stack(parts)
    merged = merge_packed(whole.layout, parts).cpu().numpy()
    host = merge_packed_host(whole.layout, parts.cpu().numpy())
    L = whole.layout
    np.testing.assert_allclose(merged[:L.n_sums], host[:L.n_sums], rtol=1e-14)
    assert np.array_equal(merged[L.n_sums:], host[L.n_sums:])
    ref = whole.packed.cpu().numpy()
    assert np.array_equal(merged[:3], ref[:3]) and np.array_equal(merged[[3, 6, 9]], ref[[3, 6, 9]])
    assert np.array_equal(merged[L.off_hist:L.n_sums], ref[L.off_hist:L.n_sums]), 'histograms differ between shardings'
    assert np.array_equal(merged[L.n_sums:], ref[L.n_sums:])
    np.testing.assert_allclose(merged[:L.off_hist], ref[:L.off_hist], rtol=1e-12)


def test_multi_device_reducer_matches_single_device(cuda_device):
    """MonteCarloMoments(devices='all'): one process, the index range split over every visible GPU, merged on the first."""
    import torch
    from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    n, A = 400_000, 200
    one = MonteCarloMoments(n_angles=A, hist=HistogramSpec(), device=0, torr=133.322)
    one.accumulate_sampled(n, 5, 0)
    many = MonteCarloMoments(n_angles=A, hist=HistogramSpec(), devices='all', torr=133.322)
    many.accumulate_sampled(n, 5, 0)
    a, b = one.result(), many.result()
    L = one.layout
    assert np.array_equal(a.sums[:3], b.sums[:3]) and np.array_equal(a.sums[L.off_hist:], b.sums[L.off_hist:])
    np.testing.assert_allclose(a.sums[:L.off_hist], b.sums[:L.off_hist], rtol=1e-12)
    assert np.array_equal(a.minmax, b.minmax)
    # host arrays split over the devices (pinned, threaded staging) == the same arrays on one device
    from hallthrusterpem_b200.synthetic import spt100_batch
    host = spt100_batch(100_000, 21)
    one2 = MonteCarloMoments(n_angles=91, hist=HistogramSpec(), device=0, torr=133.322, scalar_shift=(30.0, 0.8, 0.05))
    one2.accumulate({k: torch.as_tensor(v, device='cuda:0') for k, v in host.items()})
    many2 = MonteCarloMoments(n_angles=91, hist=HistogramSpec(), devices='all', torr=133.322, scalar_shift=(30.0, 0.8, 0.05))
    many2.accumulate(host)
    c, d = one2.result(), many2.result()
    L2 = one2.layout
    assert np.array_equal(c.sums[:3], d.sums[:3]) and np.array_equal(c.sums[L2.off_hist:], d.sums[L2.off_hist:])
    np.testing.assert_allclose(c.sums[:L2.off_hist], d.sums[:L2.off_hist], rtol=1e-12)
