"""`exponax.ic`, the spectral random generators, on the device.

    white noise -> exb_fft -> exb_ic_shape (per-mode factor) -> exb_ifft -> exb_ic_normalize

with the noise drawn on the GPU.  Class names, constructor arguments, validation and the arithmetic follow
exponax/ic/_white_noise.py, _truncated_fourier_series.py, _gaussian_random_field.py, _diffused_noise.py and
_base_ic.py.  Differences forced by the missing JAX: `key` is an integer seed or a `torch.Generator` on the
CUDA device (the random stream is torch's Philox, not jax.random's threefry -- same distribution, other
values), and every generator takes an optional `noise=` array so that the deterministic part can be checked
against the oracle.  `__call__(num_points, key=...)` returns `(1, N, .., N)`; `batch(num_points, num_samples,
key=...)` returns `(S, 1, N, .., N)` from ONE batched pipeline (what `ex.build_ic_set` uses).  The closed-form
generators of the reference (sine waves, Gaussian blobs, discontinuities, clamping / scaling / multi-channel
wrappers) are not mirrored here."""
from __future__ import annotations

import numpy as np

from .. import _array as A
from .. import _native as nat
from .. import _spectral as sp
from .._config import real_dtype

__all__ = ["BaseRandomICGenerator", "WhiteNoise", "RandomTruncatedFourierSeries", "GaussianRandomField",
           "DiffusedNoise", "normalize_ic", "validate_normalization_options"]


def validate_normalization_options(*, zero_mean: bool, std_one: bool, max_one: bool):
    """exponax/ic/_base_ic.py:9-13."""
    if not zero_mean and std_one:
        raise ValueError("Cannot have `zero_mean=False` and `std_one=True`.")
    if std_one and max_one:
        raise ValueError("Cannot have `std_one=True` and `max_one=True`.")


def _generator(key):
    torch = A.torch
    if isinstance(key, torch.Generator):
        return key
    g = torch.Generator(device="cuda")
    g.manual_seed(int(key))
    return g


def normalize_ic(ic, *, zero_mean: bool = True, std_one: bool = False, max_one: bool = False, _lead: int = 0,
                 _inplace: bool = False):
    """exponax/ic/_base_ic.py:16-33: statistics over the WHOLE array, as the reference takes them.
    (`_lead`: number of leading axes that index independent arrays -- the batched generators use 1.)"""
    rd = real_dtype()
    t, kind = A.to_device(ic, rd)
    if not _inplace and kind == "torch" and t.data_ptr() == getattr(ic, "data_ptr", lambda: 0)():
        t = t.clone()
    nfields = int(np.prod(t.shape[:_lead])) if _lead else 1
    npoints = t.numel() // nfields
    stats = A.torch.empty(nfields * 4, dtype=A.torch.float64, device="cuda")
    for f0 in range(0, nfields, 65535):
        n = min(65535, nfields - f0)
        nat.check(nat.lib().exb_ic_normalize(
            A.stream_ptr(), nat.EXB_F32 if rd == np.float32 else nat.EXB_F64, n, npoints,
            t.data_ptr() + f0 * npoints * t.element_size(), int(zero_mean), int(std_one), int(max_one),
            stats.data_ptr() + f0 * 32))
    return A.from_device(t, kind)


class BaseRandomICGenerator:
    """Random generators draw `(S, 1, N, .., N)` white noise and push it through `_from_noise`."""
    num_spatial_dims: int

    def _noise(self, num_points, num_samples, key, noise):
        rd = real_dtype()
        shape = (num_samples, 1) + sp.spatial_shape(self.num_spatial_dims, num_points)
        if noise is not None:
            t, _ = A.to_device(noise, rd)
            return t.reshape(shape).clone()
        return A.torch.randn(shape, generator=_generator(key), device="cuda", dtype=A.real_t(rd))

    def _from_noise(self, noise, num_points, gen):
        raise NotImplementedError

    def batch(self, num_points: int, num_samples: int, *, key=0, noise=None, **kw):
        gen = _generator(key)
        out = self._from_noise(self._noise(num_points, num_samples, gen, noise), num_points, gen, **kw)
        return A.from_device(out, "numpy" if isinstance(noise, np.ndarray) else "torch")

    def __call__(self, num_points: int, *, key=0, noise=None, **kw):
        return self.batch(num_points, 1, key=key, noise=noise, **kw)[0]


class WhiteNoise(BaseRandomICGenerator):
    """exponax/ic/_white_noise.py:9-30."""

    def __init__(self, num_spatial_dims: int, *, std: float = 1.0):
        self.num_spatial_dims = num_spatial_dims
        self.std = std

    def _from_noise(self, noise, num_points, gen):
        return noise * self.std if self.std != 1.0 else noise


def _shape_and_back(noise, D, N, kind, param, L, dc_values=None):
    """fft -> exb_ic_shape -> (per-sample DC) -> ifft on a (S, 1, N..N) device tensor."""
    rd = real_dtype()
    S = noise.shape[0]
    uh = sp.fft(noise, num_spatial_dims=D)
    plan = sp._plain_plan(D, N, rd)
    nat.check(nat.lib().exb_ic_shape(plan.handle, A.stream_ptr(), S, A.ptr(uh), kind, float(param), float(L), 0.0))
    if dc_values is not None:
        uh.view(S, -1)[:, 0] = dc_values.to(uh.dtype)
    return sp.ifft(uh, num_spatial_dims=D, num_points=N)


class RandomTruncatedFourierSeries(BaseRandomICGenerator):
    """exponax/ic/_truncated_fourier_series.py:19-100."""

    def __init__(self, num_spatial_dims: int, *, cutoff: int = 5, offset_range=(0.0, 0.0), std_one: bool = False,
                 max_one: bool = False):
        zero_mean = tuple(offset_range) == (0.0, 0.0)
        validate_normalization_options(zero_mean=zero_mean, std_one=std_one, max_one=max_one)
        self.num_spatial_dims = num_spatial_dims
        self.cutoff = cutoff
        self.offset_range = tuple(offset_range)
        self.std_one = std_one
        self.max_one = max_one
        self.white_noise = WhiteNoise(num_spatial_dims)

    def _from_noise(self, noise, num_points, gen, offsets=None):
        S = noise.shape[0]
        zero_mean = self.offset_range == (0.0, 0.0)
        dc = None
        if offsets is not None:
            dc, _ = A.to_device(np.asarray(offsets, np.float64).reshape(S), np.float64)
        elif not zero_mean:
            lo, hi = self.offset_range
            dc = lo + (hi - lo) * A.torch.rand(S, generator=gen, device="cuda", dtype=A.torch.float64)
        ic = _shape_and_back(noise, self.num_spatial_dims, num_points, 0, self.cutoff, 1.0, dc)
        return normalize_ic(ic, zero_mean=zero_mean, std_one=self.std_one, max_one=self.max_one, _lead=1, _inplace=True)


class GaussianRandomField(BaseRandomICGenerator):
    """exponax/ic/_gaussian_random_field.py:18-93: power spectrum ~ |k|^(-powerlaw_exponent)."""

    def __init__(self, num_spatial_dims: int, *, domain_extent: float = 1.0, powerlaw_exponent: float = 3.0,
                 zero_mean: bool = True, std_one: bool = False, max_one: bool = False):
        validate_normalization_options(zero_mean=zero_mean, std_one=std_one, max_one=max_one)
        self.num_spatial_dims = num_spatial_dims
        self.domain_extent = domain_extent
        self.powerlaw_exponent = powerlaw_exponent
        self.zero_mean = zero_mean
        self.std_one = std_one
        self.max_one = max_one
        self.white_noise = WhiteNoise(num_spatial_dims)

    def _from_noise(self, noise, num_points, gen):
        ic = _shape_and_back(noise, self.num_spatial_dims, num_points, 1, self.powerlaw_exponent, self.domain_extent)
        return normalize_ic(ic, zero_mean=self.zero_mean, std_one=self.std_one, max_one=self.max_one, _lead=1,
                            _inplace=True)


class DiffusedNoise(BaseRandomICGenerator):
    """exponax/ic/_diffused_noise.py:13-77: one exact diffusion step (dt = 1, nu = intensity) of white noise."""

    def __init__(self, num_spatial_dims: int, *, domain_extent: float = 1.0, intensity=0.001, zero_mean: bool = True,
                 std_one: bool = False, max_one: bool = False):
        validate_normalization_options(zero_mean=zero_mean, std_one=std_one, max_one=max_one)
        self.num_spatial_dims = num_spatial_dims
        self.domain_extent = domain_extent
        self.intensity = intensity
        self.zero_mean = zero_mean
        self.std_one = std_one
        self.max_one = max_one
        self.white_noise = WhiteNoise(num_spatial_dims)

    def _from_noise(self, noise, num_points, gen):
        from ..stepper import Diffusion
        stepper = Diffusion(self.num_spatial_dims, self.domain_extent, num_points, 1.0, diffusivity=self.intensity)
        ic = stepper._step_batched(noise)
        return normalize_ic(ic, zero_mean=self.zero_mean, std_one=self.std_one, max_one=self.max_one, _lead=1,
                            _inplace=True)
