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
