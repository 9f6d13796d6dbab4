"""CPU tests of the host side of the SVD field compression (compute_map) against the restated amisc algorithm."""
import numpy as np
import pytest


def _compression_matrix(n=300, n_angles=91, seed=11):
    from hallthrusterpem_b200.synthetic import spt100_batch
    from oracle.compression_oracle import normalize_log10
    from oracle.ref_restated import current_density_oracle
    b = spt100_batch(n, seed)
    with np.errstate(all='ignore'):
        j = current_density_oracle(b, 1.0, n_angles, 133.322, with_coords=False)['j_ion']
    return normalize_log10(j).T          # (dof, num_samples)


@pytest.mark.parametrize('kw', [dict(reconstruction_tol=0.01), dict(rank=5), dict(energy_tol=0.999), dict()])
def test_compute_map_matches_restated_amisc(kw):
    from hallthrusterpem_b200.compression import SVD
    from oracle.compression_oracle import compute_map_oracle
    dm = _compression_matrix()
    proj, rank, energy, rec = compute_map_oracle(dm, **kw)
    c = SVD(norm='log10', **kw)
    c.compute_map(dm)
    assert c.rank == rank == c.latent_size()
    assert c.projection_matrix.shape == (dm.shape[0], rank)
    assert np.allclose(c.projection_matrix, proj, rtol=0, atol=1e-13)
    assert abs(c.energy_tol - energy) < 1e-14 and abs(c.reconstruction_tol - rec) < 1e-12
    # orthonormal columns; the tolerance actually holds on the compression set
    assert np.allclose(c.projection_matrix.T @ c.projection_matrix, np.eye(rank), atol=1e-13)
    if 'reconstruction_tol' in kw:
        assert c.reconstruction_tol <= kw['reconstruction_tol']
        if rank > 1:   # and rank - 1 would not have met it
            p1 = proj[:, :rank - 1]
            assert np.sqrt(np.sum((p1 @ p1.T @ dm - dm) ** 2) / np.sum(dm ** 2)) > kw['reconstruction_tol']
    lo_hi = c.estimate_latent_ranges()
    z = proj.T @ dm
    assert len(lo_hi) == rank and np.allclose([v[0] for v in lo_hi], z.min(axis=1)) and np.allclose([v[1] for v in lo_hi], z.max(axis=1))


def test_compute_map_accepts_the_field_dict_and_drops_nan_samples():
    from hallthrusterpem_b200.compression import SVD
    dm = _compression_matrix(n=120)
    dm_nan = dm.copy()
    dm_nan[3, 7] = np.nan
    a = SVD(reconstruction_tol=0.01)
    a.compute_map({'j_ion': dm_nan.T})                       # gen_data.py:288-290 passes {field: (num_samples, dof)}
    b = SVD(reconstruction_tol=0.01)
    b.compute_map(np.delete(dm, 7, axis=1))
    assert a.rank == b.rank and np.allclose(a.projection_matrix, b.projection_matrix, atol=1e-13)
    with pytest.raises(ValueError):
        SVD(norm='sqrt')
    with pytest.raises(RuntimeError):
        SVD()._basis(0, True)
