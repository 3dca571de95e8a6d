"""The oracle's metrics pinned against the reference's closed-form tests
(tests/test_metrics.py:8-88, 235-252, 259-405 of the reference)."""
import numpy as np
import pytest

from oracle import exponax_np as ox


@pytest.mark.parametrize("D", [1, 2, 3])
def test_constant_offset(D):
    L, N = 5.0, 40
    g = ox.make_grid(D, L, N)
    u0 = 2.0 * np.ones_like(g[0:1])
    u1 = 4.0 * np.ones_like(g[0:1])
    assert ox.MSE(u1, u0, domain_extent=1.0) == pytest.approx(4.0)
    assert ox.MSE(u1, u0, domain_extent=L) == pytest.approx(L**D * 4.0)
    assert ox.MSE(u0, u1) == ox.MSE(u1, u0)
    assert ox.nMSE(u1, u0) == pytest.approx(1.0)
    assert ox.nMSE(u0, u1) == pytest.approx(0.25)
    assert ox.sMSE(u1, u0) == pytest.approx(0.4)
    assert ox.sMSE(u0, u1) == ox.sMSE(u1, u0)
    assert ox.RMSE(u1, u0, domain_extent=1.0) == pytest.approx(2.0)
    assert ox.RMSE(u1, u0, domain_extent=L) == pytest.approx(np.sqrt(L**D * 4.0))
    assert ox.nRMSE(u1, u0) == pytest.approx(1.0)
    assert ox.nRMSE(u0, u1) == pytest.approx(0.5)
    assert ox.sRMSE(u1, u0) == pytest.approx(2 / 3)


@pytest.mark.parametrize("k,l", [(1, 2), (2, 1), (3, 4)])
def test_analytical_solution_1d(k, l):
    L = 2 * np.pi
    g = ox.make_grid(1, L, 100)
    u0, u1 = np.sin(k * g), np.sin(l * g)
    assert ox.MSE(u1, u0, domain_extent=L) == pytest.approx(2 * np.pi)
    assert ox.nMSE(u1, u0) == pytest.approx(2.0)
    assert ox.nMSE(u1, u0, domain_extent=L) == pytest.approx(2.0)


def test_correlation():
    rng = np.random.default_rng(0)
    u = rng.standard_normal((1, 64)).astype(np.float32)
    assert ox.correlation(u, u) == pytest.approx(1.0, abs=1e-5)
    assert ox.correlation(u, -u) == pytest.approx(-1.0, abs=1e-5)
    g = ox.make_grid(1, 2 * np.pi, 128)
    assert ox.correlation(np.sin(g), np.cos(g)) == pytest.approx(0.0, abs=1e-4)
    u3 = rng.standard_normal((3, 64)).astype(np.float32)
    assert ox.correlation(u3, u3) == pytest.approx(1.0, abs=1e-5)
    u2d = rng.standard_normal((1, 32, 32)).astype(np.float32)
    assert ox.correlation(u2d, u2d) == pytest.approx(1.0, abs=1e-5)


def test_mean_metric_and_mae_family():
    rng = np.random.default_rng(1)
    up, ur = rng.standard_normal((5, 1, 64)), rng.standard_normal((5, 1, 64))
    manual = np.mean([ox.MSE(up[i], ur[i], domain_extent=2.0) for i in range(5)])
    assert ox.mean_metric(ox.MSE, up, ur, domain_extent=2.0) == pytest.approx(manual, abs=1e-6)
    assert np.ndim(ox.mean_metric(ox.RMSE, up, ur)) == 0
    a, b = 4.0 * np.ones((1, 64)), 2.0 * np.ones((1, 64))
    assert ox.MAE(a, b) == pytest.approx(2.0, abs=1e-5)
    assert ox.MAE(3.0 * np.ones((1, 64))) == pytest.approx(3.0, abs=1e-5)
    assert ox.nMAE(a, b) == pytest.approx(1.0, abs=1e-5)
    assert ox.sMAE(a, b) == pytest.approx(2 / 3, abs=1e-5)
    assert ox.sMAE(up[0], ur[0]) == pytest.approx(ox.sMAE(ur[0], up[0]), abs=1e-6)
    assert ox.MAE(a, b, domain_extent=5.0) == pytest.approx(5.0 * ox.MAE(a, b), abs=1e-5)
    with pytest.raises(ValueError, match="normalized.*requires"):
        ox.spatial_norm(a, mode="normalized")
    with pytest.raises(ValueError, match="symmetric.*requires"):
        ox.spatial_norm(a, mode="symmetric")
    assert ox.spatial_norm(np.ones((1, 64)), inner_exponent=2.0) == pytest.approx(1.0, abs=1e-5)
