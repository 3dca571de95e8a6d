"""The oracle's get_spectrum / stack_sub_trajectories pinned against the reference's own closed-form
tests (tests/test_spectrum.py:12-330 of the reference; jax.random replaced by NumPy PCG64 -- the
assertions are identities that hold for any field)."""
import numpy as np
import pytest

from oracle import exponax_np as ox

PI2 = 2 * np.pi


def f32(x):
    return np.asarray(x, np.float32)


def test_amplitude_spectrum_1d():
    g = ox.make_grid(1, PI2, 128)
    assert ox.get_spectrum(f32(3.0 * np.sin(g)), power=False)[0, 1] == pytest.approx(3.0)
    assert ox.get_spectrum(f32(3.0 * np.cos(2 * g)), power=False)[0, 2] == pytest.approx(3.0)
    s = ox.get_spectrum(f32(3.0 * np.sin(3 * g) + 4.0 * np.cos(3 * g)), power=False)
    assert s[0, 3] == pytest.approx(5.0)
    assert ox.get_spectrum(f32(3.0 * np.ones_like(g)), power=False)[0, 0] == pytest.approx(3.0)


def test_amplitude_spectrum_2d():
    g = ox.make_grid(2, PI2, 48)
    for ax in (0, 1):
        assert ox.get_spectrum(f32(3.0 * np.sin(g[ax:ax + 1])), power=False)[0, 1] == pytest.approx(3.0)
        assert ox.get_spectrum(f32(3.0 * np.cos(2 * g[ax:ax + 1])), power=False)[0, 2] == pytest.approx(3.0)
    s = ox.get_spectrum(f32(3.0 * np.sin(g[0:1]) * np.cos(g[1:2])), power=False)
    assert s[0, 1] == pytest.approx(3.0)
    # |(2, 2)| = 2.83 falls into the 3-bin [2.5, 3.5), not the 2-bin
    s = ox.get_spectrum(f32(3.0 * np.sin(2 * g[0:1]) * np.cos(2 * g[1:2])), power=False)
    assert s[0, 3] == pytest.approx(3.0)
    assert s[0, 2] == pytest.approx(0.0, abs=1e-5)


def test_amplitude_spectrum_3d():
    g = ox.make_grid(3, PI2, 16)
    for ax in range(3):
        assert ox.get_spectrum(f32(3.0 * np.sin(g[ax:ax + 1])), power=False)[0, 1] == pytest.approx(3.0, abs=1e-4)
    u = f32(3.0 * np.sin(g[0:1]) * np.cos(g[1:2]) * np.sin(g[2:3]))     # |k| = sqrt(3) -> bin 2
    assert ox.get_spectrum(u, power=False)[0, 2] == pytest.approx(3.0, abs=1e-4)
    assert ox.get_spectrum(f32(5.0 * np.ones_like(g[0:1])), power=False)[0, 0] == pytest.approx(5.0, abs=1e-4)


def _parseval(u, rel):
    s = ox.get_spectrum(u, power=True, radial_binning="sum")
    assert float(np.sum(s)) == pytest.approx(float(0.5 * np.mean(u.astype(np.float64) ** 2)), rel=rel)


def test_power_spectrum_and_parseval_1d():
    N = 16
    g = ox.make_grid(1, PI2, N)
    _parseval(f32(5.0 * np.cos(3 * g)), 1e-5)
    _parseval(f32(3.0 * np.ones_like(g)), 1e-5)
    _parseval(f32(2.0 * (-1.0) ** np.arange(N))[None, :], 1e-5)          # Nyquist mode
    g = ox.make_grid(1, PI2, 64)
    _parseval(f32(2.0 * np.cos(g) + 3.0 * np.sin(5 * g) + 1.0), 1e-5)


def test_power_spectrum_and_parseval_2d_3d():
    g = ox.make_grid(2, PI2, 32)
    _parseval(f32(4.0 * np.cos(3 * g[0:1])), 1e-4)
    _parseval(f32(4.0 * np.cos(3 * g[1:2])), 1e-4)
    _parseval(f32(3.0 * np.ones_like(g[0:1])), 1e-4)
    _parseval(f32(2.0 * np.cos(g[0:1]) + 3.0 * np.sin(5 * g[1:2]) + 1.0), 1e-4)
    _parseval(f32(4.0 * np.sin(2 * g[0:1]) * np.cos(3 * g[1:2])), 1e-4)
    g = ox.make_grid(3, PI2, 16)
    _parseval(f32(2.0 * np.cos(g[0:1]) + 3.0 * np.sin(2 * g[2:3]) + 1.0), 1e-3)


def _random_in_nyquist_sphere(D, N, seed):
    rng = np.random.default_rng(seed)
    shape = (1,) + ox.wavenumber_shape(D, N)
    uh = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    uh = uh * ox.low_pass_filter_mask(D, N, cutoff=N // 2, axis_separate=False)
    return ox.ifft(uh, num_spatial_dims=D, num_points=N).astype(np.float32)


@pytest.mark.parametrize("D,N,rel", [(2, 32, 1e-4), (3, 16, 1e-3)])
def test_parseval_random(D, N, rel):
    _parseval(_random_in_nyquist_sphere(D, N, 0), rel)


def test_binning_average_vs_sum():
    g = ox.make_grid(1, PI2, 64)
    u = f32(3.0 * np.sin(g) + 2.0 * np.cos(5 * g))
    assert np.allclose(ox.get_spectrum(u, radial_binning="sum"), ox.get_spectrum(u, radial_binning="average"))
    for D, N in ((2, 32), (3, 16)):
        wn = ox.build_wavenumbers(D, N)
        norm = np.sqrt(np.sum(wn * wn, axis=0))
        count = np.array([np.sum((norm >= k - 0.5) & (norm < k + 0.5)) for k in range(N // 2 + 1)])
        u = _random_in_nyquist_sphere(D, N, 1)
        ssum = ox.get_spectrum(u, radial_binning="sum")
        savg = ox.get_spectrum(u, radial_binning="average")
        for k in range(N // 2 + 1):
            if count[k] > 0 and savg[0, k] > 1e-10:
                assert ssum[0, k] / savg[0, k] == pytest.approx(count[k], rel=1e-4)
        if D == 2:   # half of the 2 pi k shell: the rfft grid stores k_last >= 0 only
            assert np.allclose(count[5:N // 4], np.pi * np.arange(5, N // 4), rtol=0.2)


def test_stack_sub_trajectories():
    trj = np.arange(7 * 2 * 3, dtype=np.float32).reshape(7, 2, 3)
    s = ox.stack_sub_trajectories(trj, 3)
    assert s.shape == (5, 3, 2, 3)
    for i in range(5):
        assert np.array_equal(s[i], trj[i:i + 3])
    with pytest.raises(ValueError):
        ox.stack_sub_trajectories(trj, 8)


# ---- exponax.derivative / wrap_bc (tests/test_spectral.py:36-86, 172-196, 447-480; test_utils.py:8-24) ----
@pytest.mark.parametrize("D,axis", [(1, 0), (2, 0), (2, 1)])
def test_derivative_of_sine(D, axis):
    L, k = 3.0, 3
    g = ox.make_grid(D, L, 64)
    u = np.sin(k * 2 * np.pi * g[axis:axis + 1] / L).astype(np.float32)
    want = k * 2 * np.pi / L * np.cos(k * 2 * np.pi * g[axis] / L)
    assert np.allclose(ox.derivative(u, L, order=1)[axis], want, atol=1e-4)


def test_higher_order_and_multi_channel_derivatives():
    g = ox.make_grid(1, 2 * np.pi, 64)
    assert np.allclose(ox.derivative(np.sin(3 * g).astype(np.float32), 2 * np.pi, order=2), -9 * np.sin(3 * g), atol=1e-3)
    g = ox.make_grid(1, 2 * np.pi, 32)
    u = np.sin(g).astype(np.float32)
    assert np.allclose(ox.derivative(u, 2 * np.pi, order=3), -np.cos(g), atol=1e-3)
    assert np.allclose(ox.derivative(u, 2 * np.pi, order=4), np.sin(g), atol=0.01)
    u2 = np.concatenate([np.sin(g), np.cos(2 * g)]).astype(np.float32)
    d = ox.derivative(u2, 2 * np.pi)
    assert d.shape == (2, 1, 32)
    assert np.allclose(d[0, 0], np.cos(g[0]), atol=1e-4) and np.allclose(d[1, 0], -2 * np.sin(2 * g[0]), atol=1e-4)
    g2 = ox.make_grid(2, 1.0, 16)
    u = np.concatenate([np.sin(2 * np.pi * g2[0:1]), np.cos(2 * np.pi * g2[1:2])]).astype(np.float32)
    d = ox.derivative(u, 1.0)
    assert d.shape == (2, 2, 16, 16) and np.all(np.isfinite(d))


@pytest.mark.parametrize("D", [1, 2, 3])
def test_wrap_bc(D):
    L = 3.0
    u = np.sin(2 * np.pi * ox.make_grid(D, L, 10)[0:1] / L)
    full = np.sin(2 * np.pi * np.stack(np.meshgrid(*([np.linspace(0, L, 11)] * D), indexing="ij"))[0:1] / L)
    assert np.allclose(ox.wrap_bc(u), full, atol=1e-5)
