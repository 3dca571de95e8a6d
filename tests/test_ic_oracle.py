"""The oracle's IC restatements pinned against the property tests of the reference
(tests/test_ic.py:187-245, 505-530, 549-570 of the reference: normalisation options, Nyquist-spike
regression, band limit) -- the reference's generators are seeded with jax.random, so only properties,
not values, are available."""
import numpy as np
import pytest

from oracle import exponax_np as ox


def _noise(D, N, seed=0, dtype=np.float32):
    return np.random.default_rng(seed).standard_normal((1,) + (N,) * D).astype(dtype)


@pytest.mark.parametrize("D,N", [(1, 64), (2, 32), (3, 16)])
def test_diffused_noise_and_grf_normalisation(D, N):
    n = _noise(D, N)
    for fn, kw in ((ox.ic_diffused_noise, dict(intensity=0.001)), (ox.ic_gaussian_random_field, dict())):
        assert abs(float(np.mean(fn(n, zero_mean=True, **kw)))) < 1e-5
        assert float(np.std(fn(n, std_one=True, **kw))) == pytest.approx(1.0, abs=1e-4)
        assert float(np.max(np.abs(fn(n, max_one=True, **kw)))) == pytest.approx(1.0, abs=1e-5)
        assert fn(n, **kw).shape == (1,) + (N,) * D


def test_truncated_series_band_limit_and_offset():
    n = _noise(1, 64, 3)
    u = ox.ic_truncated_fourier_series(n, cutoff=3)
    s = ox.get_spectrum(u, power=False)
    assert np.all(s[0, 4:] < 1e-6) and np.all(s[0, 1:4] > 1e-3) and s[0, 0] < 1e-6
    # the offset is written as an UNNORMALISED DC coefficient (reference line 87): mean = offset / N
    u = ox.ic_truncated_fourier_series(n, cutoff=3, offset=32.0, zero_mean=False)
    assert float(np.mean(u)) == pytest.approx(0.5, rel=1e-5)
    u = ox.ic_truncated_fourier_series(_noise(2, 32), cutoff=2, max_one=True)
    assert float(np.max(np.abs(u))) == pytest.approx(1.0, abs=1e-6)
    uh = np.abs(ox.fft(u))
    k = ox.build_wavenumbers(2, 32)
    assert np.all(uh[0][(np.abs(k[0]) > 2) | (np.abs(k[1]) > 2)] < 1e-3)


@pytest.mark.parametrize("fn", [ox.ic_gaussian_random_field, ox.ic_truncated_fourier_series])
def test_no_nyquist_spike_2d(fn):
    spectrum = ox.get_spectrum(fn(_noise(2, 32)))
    assert float(spectrum[0, -1]) < 10 * float(spectrum[0, -3]) + 1e-12


def test_grf_power_law_slope():
    # power spectrum ~ k^-alpha per mode: the radially AVERAGED spectrum of many realisations has slope -alpha
    acc = 0
    for seed in range(16):
        acc = acc + ox.get_spectrum(ox.ic_gaussian_random_field(_noise(2, 64, seed, np.float64), powerlaw_exponent=3.0,
                                                                std_one=True), radial_binning="average")[0]
    k = np.arange(2, 20)
    slope = np.polyfit(np.log(k), np.log(acc[2:20]), 1)[0]
    assert slope == pytest.approx(-3.0, abs=0.15)


def test_normalize_ic_options():
    x = _noise(1, 50, 1) + 3.0
    assert abs(np.mean(ox.normalize_ic(x))) < 1e-6
    assert np.std(ox.normalize_ic(x, std_one=True)) == pytest.approx(1.0, abs=1e-5)
    assert np.max(np.abs(ox.normalize_ic(x, zero_mean=False, max_one=True))) == pytest.approx(1.0, abs=1e-6)
