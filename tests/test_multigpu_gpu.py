"""GPU tests that need at least two devices (skipped on a single-GPU box): the single-process multi-GPU host path and the
multi-rank merge of the reduce-only pass over NCCL."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _need_two():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    return torch.cuda.device_count()


def test_current_density_device_all_is_bit_identical_to_one_gpu(cuda_device):
    """`device='all'`: ONE process (the reference's actual caller, amisc, pem_v0_SPT-100.yml:5-7,215-218) shards the
    samples at multiples of 64 over every visible GPU; every output equals the single-GPU call bit for bit."""
    _need_two()
    from hallthrusterpem_b200.models import cathode_coupling, current_density, plume_cathode
    from hallthrusterpem_b200.synthetic import spt100_batch
    for n, A in ((100_003, 91), (50_000, 200)):
        b = spt100_batch(n, 5 + A)
        one = plume_cathode(b, 1.0, n_angles=A, device=0, extras=True)
        many = plume_cathode(b, 1.0, n_angles=A, device='all', extras=True)
        for k in ('V_cc', 'j_ion', 'div_angle', 'T_c', 'cos_div', 'invalid'):
            assert np.array_equal(one[k], many[k], equal_nan=True), (n, A, k)
        assert many['j_ion_coords'].shape == (n,)
    b = spt100_batch(4097, 3)
    r = np.linspace(1.0, 1.2, 5)
    assert np.array_equal(current_density(b, r, device=0)['j_ion'], current_density(b, r, device=[0, 1])['j_ion'])
    assert np.array_equal(cathode_coupling(b, device=0)['V_cc'], cathode_coupling(b, device='all')['V_cc'])
    tiny = spt100_batch(100, 1)     # fewer than 64 samples per device: one device takes the call
    assert np.array_equal(current_density(tiny, device=0)['j_ion'], current_density(tiny, device='all')['j_ion'])


def test_log_likelihood_device_all_matches_one_gpu(cuda_device):
    _need_two()
    from hallthrusterpem_b200.likelihood import JionMeasurements, jion_log_likelihood
    from hallthrusterpem_b200.synthetic import spt100_batch
    rng = np.random.default_rng(0)
    m = 40
    theta, y, sg = rng.uniform(-1.5, 1.5, m), 10 ** rng.uniform(-2, 1, m), np.full(m, 0.1)
    b = {k: v for k, v in spt100_batch(70_001, 8).items() if k in ('P_b', 'c0', 'c1', 'c2', 'c3', 'c4', 'c5', 'sigma_cex', 'I_B0')}
    one = jion_log_likelihood(b, JionMeasurements(theta, y, sg, device=0), return_pred=True)
    many = jion_log_likelihood(b, JionMeasurements(theta, y, sg, device='all'), return_pred=True)
    assert np.array_equal(one[0], many[0]) and np.array_equal(one[1], many[1])


def _nccl_worker(rank, world, port, n, n_angles, seed, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
    from hallthrusterpem_b200.synthetic import shard_bounds
    lo, hi = shard_bounds(n, world, rank)
    mc = MonteCarloMoments(n_angles=n_angles, hist=HistogramSpec(), device=rank, torr=133.322)
    for first in range(lo, hi, 50_000):                               # several chunks per rank
        mc.accumulate_sampled(min(50_000, hi - first), seed, first)
    mc.merge()                                                        # ONE all-gather + fixed-order merge
    torch.cuda.synchronize()
    q.put((rank, mc.packed.cpu().numpy().copy()))
    dist.destroy_process_group()


def test_nccl_merge_over_all_gpus_equals_single_gpu(cuda_device):
    """Shard -> K2 -> merge() over N real ranks (NCCL) == the single-GPU packed vector: counts, histograms and min/max bit
    for bit, sums / centred second moments to 1e-12; all ranks end with identical bits."""
    world = _need_two()
    import torch
    import torch.multiprocessing as mp
    from hallthrusterpem_b200.mc import HistogramSpec, MonteCarloMoments
    n, A, seed = 1_000_000, 256, 99
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + int(np.random.default_rng().integers(0, 2000))
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, n, A, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(1, world):
        assert np.array_equal(got[0], got[r]), f'rank {r} holds different bits than rank 0'
    ref = MonteCarloMoments(n_angles=A, hist=HistogramSpec(), device=0, torr=133.322)
    ref.accumulate_sampled(n, seed, 0)
    torch.cuda.synchronize()
    r0, L = ref.packed.cpu().numpy(), ref.layout
    a = got[0]
    assert np.array_equal(a[:3], r0[:3]) and np.array_equal(a[[3, 6, 9]], r0[[3, 6, 9]])
    assert np.array_equal(a[L.off_hist:L.n_sums], r0[L.off_hist:L.n_sums]), 'histograms differ'
    assert np.array_equal(a[L.n_sums:], r0[L.n_sums:]), 'min/max differ'
    np.testing.assert_allclose(a[:L.off_hist], r0[:L.off_hist], rtol=1e-12)
