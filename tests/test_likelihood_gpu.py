"""GPU tests of K3 (interpolation to probe angles + Gaussian log-likelihood) against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,n_angles,m', [(3000, 100, 37), (513, 91, 5), (2000, 256, 300), (64, 17, 1)])
def test_loglike_and_interpolation_match_oracle(n, n_angles, m, cuda_device):
    import torch
    from hallthrusterpem_b200.likelihood import JionMeasurements, jion_log_likelihood
    from hallthrusterpem_b200.synthetic import spt100_batch
    from oracle.likelihood_oracle import jion_log_likelihood_oracle
    from oracle.make_golden import edge_batch
    rng = np.random.default_rng(n + m)
    b = spt100_batch(n, 55 + n)
    if n == 3000:                                   # mix in the edge rows (invalid samples, NaN rows, needle beams)
        e = edge_batch()
        b = {k: np.concatenate([e[k], b[k]]) for k in b}
        b['c2'][40], b['c3'][40], b['c1'][40] = -15.0, -0.5, np.nan      # alpha1 <= 0 AND a NaN elsewhere: still the 1e-20 row
    theta = rng.uniform(-np.pi / 2, np.pi / 2, m)
    theta[0] = 0.0
    if m > 3:
        theta[1], theta[2] = np.pi / 2, -np.pi / 2                      # end points and an exact grid node
        theta[3] = np.linspace(0, np.pi / 2, n_angles)[n_angles // 3]
    y = 10 ** rng.uniform(-2, 1.5, m)
    sigma = 0.2 * y + 0.01
    meas = JionMeasurements(theta, y, sigma, n_angles=n_angles, device=0)
    dev = {k: torch.as_tensor(v, device='cuda:0') for k, v in b.items()}
    ll, pred = jion_log_likelihood(dev, meas, torr=133.322, return_pred=True)
    ll2 = jion_log_likelihood(dev, meas, torr=133.322)
    assert torch.equal(ll, ll2) or torch.equal(torch.isnan(ll), torch.isnan(ll2))
    ref_ll, ref_pred = jion_log_likelihood_oracle(b, theta, y, sigma, n_angles, 133.322)
    got_pred, got_ll = pred.cpu().numpy(), ll.cpu().numpy()
    assert got_pred.shape == ref_pred.shape == (len(b['P_b']), m)
    assert np.array_equal(np.isnan(got_pred), np.isnan(ref_pred))
    ok = ~np.isnan(ref_pred)
    # interpolated predictions: rel 1e-12 plus the j_cex cancellation floor of tests/parity.py
    floor = 2 * np.finfo(float).eps * np.abs(b['I_B0'])[:, None] / (2 * np.pi)
    assert np.all(np.abs(got_pred - ref_pred)[ok] <= (1e-12 * np.abs(ref_pred) + floor)[ok])
    # log-likelihood: a sum of squared residuals -> error bound from the per-term conditioning |dL| <= sum |r| * |dy_hat| / sigma
    okl = ~np.isnan(ref_ll)
    assert np.array_equal(np.isnan(got_ll), np.isnan(ref_ll))
    resid = np.abs((y - ref_pred) / sigma)
    bound = np.nansum(resid * (1e-12 * np.abs(ref_pred) + floor) / sigma, axis=-1) + 1e-13 * np.abs(ref_ll)
    assert np.all(np.abs(got_ll - ref_ll)[okl] <= bound[okl])


def test_loglike_rejects_out_of_range_angles(cuda_device):
    from hallthrusterpem_b200.likelihood import JionMeasurements
    with pytest.raises(ValueError):
        JionMeasurements([0.1, 1.6], [1.0, 1.0], [0.1, 0.1], n_angles=91, device=0)


def test_loglike_accepts_host_inputs(cuda_device):
    """NumPy arrays and plain scalars (one MCMC step) go through the same kernel and come back as NumPy."""
    import torch
    from hallthrusterpem_b200.likelihood import JionMeasurements, jion_log_likelihood
    from hallthrusterpem_b200.synthetic import spt100_batch
    b = {k: v for k, v in spt100_batch(257, 3).items() if k in ('P_b', 'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0')}
    meas = JionMeasurements(np.linspace(-1.2, 1.2, 9), np.full(9, 0.5), np.full(9, 0.1), n_angles=91, device=0)
    dev = jion_log_likelihood({k: torch.as_tensor(v, device='cuda:0') for k, v in b.items()}, meas, torr=133.322)
    host, pred = jion_log_likelihood(b, meas, torr=133.322, return_pred=True)
    assert isinstance(host, np.ndarray) and host.shape == (257,) and pred.shape == (257, 9)
    assert np.array_equal(host, dev.cpu().numpy())
    one = jion_log_likelihood({k: float(v[5]) for k, v in b.items()}, meas, torr=133.322)
    assert one.shape == (1,) and one[0] == host[5]


def test_marginal_log_likelihood_matches_oracle(cuda_device):
    """log-sum-exp over the M draws (mcmc.py:101-102): wide dynamic range, -inf entries, NaN propagation, M = 1."""
    import torch
    from hallthrusterpem_b200.likelihood import marginal_log_likelihood
    from oracle.likelihood_oracle import marginal_log_likelihood_oracle
    rng = np.random.default_rng(4)
    for shape in ((1000, 64), (7, 3, 33), (5, 1), (2000, 500)):
        ll = -10.0 ** rng.uniform(-1, 4, shape)
        ll.reshape(-1)[::97] = -np.inf
        ll.reshape(-1, shape[-1])[3, :] = -np.inf            # a whole group at -inf: nan, as in NumPy
        ll.reshape(-1, shape[-1])[4, 0] = np.nan
        ref = marginal_log_likelihood_oracle(ll)
        got = marginal_log_likelihood(torch.as_tensor(ll, device='cuda:0')).cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.all(np.abs(got[ok] - ref[ok]) <= 1e-12 * np.abs(ref[ok]) + 1e-13)
        assert np.array_equal(marginal_log_likelihood(ll), got, equal_nan=True)     # NumPy in -> NumPy out
