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


# ---- Fourier / Sobolev families (tests/test_metrics.py:95-224, 405-425 of the reference) -------------------
def _series(D, N, seed, offset):
    noise = np.random.default_rng(seed).standard_normal((1,) + (N,) * D).astype(np.float32)
    return ox.ic_truncated_fourier_series(noise, offset=offset * N**D, zero_mean=False)


@pytest.mark.parametrize("D", [1, 2, 3])
def test_fourier_equals_spatial_aggregation(D):
    L = 5.0
    u0, u1 = _series(D, 40, 0, 0.3), _series(D, 40, 1, -0.6)
    assert ox.fourier_MSE(u1, u0, domain_extent=L) == pytest.approx(ox.MSE(u1, u0, domain_extent=L), rel=1e-5)
    assert ox.fourier_RMSE(u1, u0, domain_extent=L) == pytest.approx(ox.RMSE(u1, u0, domain_extent=L), rel=1e-5)


@pytest.mark.parametrize("D", [1, 2, 3])
@pytest.mark.parametrize("N", [40, 41])
def test_fourier_metric_filtering(D, N):
    g = ox.make_grid(D, 2 * np.pi, N)
    u = np.sin(4 * g[0:1])
    if D > 1:
        u = u * np.sin(4 * g[1:2])
    if D > 2:
        u = u * np.sin(4 * g[2:3])
    u = u.astype(np.float32)
    nonzero = lambda **kw: abs(float(ox.fourier_MSE(u, **kw))) > 1e-6
    assert nonzero()
    assert not nonzero(low=8)
    assert nonzero(high=8)
    assert not nonzero(high=2)
    assert nonzero(low=2)
    assert nonzero(low=2, high=8)
    assert not nonzero(low=8, high=16)
    assert not nonzero(low=0, high=2)
    assert nonzero(low=4, high=8)
    assert nonzero(low=0, high=4)


@pytest.mark.parametrize("D", [1, 2, 3])
@pytest.mark.parametrize("name", ["MAE", "nMAE", "MSE", "nMSE", "RMSE", "nRMSE"])
def test_sobolev_vs_manual(D, name):
    L = 5.0
    u0, u1 = _series(D, 40, 0, 0.3), _series(D, 40, 1, -0.6)
    f = getattr(ox, "fourier_" + name)
    want = f(u0, u1, domain_extent=L) + f(u0, u1, domain_extent=L, derivative_order=1)
    assert getattr(ox, "H1_" + name)(u0, u1, domain_extent=L) == pytest.approx(want)


def test_fourier_norm_edge_cases():
    with pytest.raises(ValueError, match="normalized"):
        ox.fourier_norm(np.ones((1, 64), np.float32), mode="normalized")
    g = ox.make_grid(1, 2 * np.pi, 64)
    r = float(ox.fourier_norm(np.sin(g).astype(np.float32), inner_exponent=2.0))
    assert r > 0 and np.isfinite(r)
    # first derivative of sin(3x): |d/dx| = 3 -> derivative-weighted RMSE = 3 x plain RMSE
    u = np.sin(3 * g).astype(np.float32)
    assert ox.fourier_RMSE(u, derivative_order=1, domain_extent=2 * np.pi) == pytest.approx(
        3.0 * ox.fourier_RMSE(u, domain_extent=2 * np.pi), rel=1e-5)
