"""
ORACLE -- test infrastructure only.  NOT part of the product path.

NumPy CPU restatement of the exponax forward ETDRK hot path (reference:
Ceyron/exponax, mounted read-only at /root/reference while this was written).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this module; the package `exponax_b200` never does.

Parity pinning: the reference is pure Python on JAX + Equinox, neither of which
is installable in this image (no network), so the reference itself cannot be
executed here.  The arithmetic below follows the reference file:line cited at
each function, and the oracle is pinned (tests/test_oracle_*.py) against every
known-answer the reference's own tests hold for this path:
  * the dealiasing-mask golden arrays of tests/test_filter_masks.py
    (extracted verbatim into tests/golden/filter_masks.json by
    tests/golden/make_filter_mask_golden.py),
  * the analytical nonlinear-function / linear-stepper / ETDRK-order tests
    (tests/test_nonlinear_funs.py, test_validation.py, test_etdrk.py),
  * the Taylor-Green errors recorded in validation/validate_taylor_green.ipynb.
Trajectory-level golden vectors do not exist in the reference; chaotic runs are
"parity unpinned" beyond those checks.

The third-party arithmetic the reference delegates to (jax.numpy.fft.rfftn /
irfftn, jax>=0.4.13 unpinned in pyproject.toml:12) is restated with
numpy.fft / scipy.fft (pocketfft family, same as jaxlib's CPU backend), which
keep float32 -> complex64 under NumPy >= 2.

Dtype convention mirrors JAX's weak typing: every array is created in the
working real dtype `dtype` (float32 default, float64 == jax_enable_x64) and
Python scalars never promote it.
"""
from __future__ import annotations

import os

import numpy as np

try:  # scipy's pocketfft supports multi-threading (used by the CPU baseline)
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None

_WORKERS = 1


def set_fft_workers(n: int | None):
    """Number of host threads the FFTs may use (bench CPU baseline)."""
    global _WORKERS
    _WORKERS = int(n) if n else (os.cpu_count() or 1)


def _cdtype(dtype):
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


# --------------------------------------------------------------------------
# _spectral.py
# --------------------------------------------------------------------------
def build_wavenumbers(D, N, dtype=np.float32):
    """exponax/_spectral.py:13-49 (rfftfreq on last axis, fftfreq on others, 'ij')."""
    right = np.fft.rfftfreq(N, 1.0 / N).astype(dtype)
    other = np.fft.fftfreq(N, 1.0 / N).astype(dtype)
    lst = [other] * (D - 1) + [right]
    return np.stack(np.meshgrid(*lst, indexing="ij"))


def build_scaled_wavenumbers(D, L, N, dtype=np.float32):
    """exponax/_spectral.py:52-83; scale is a Python double, weakly typed."""
    scale = 2 * np.pi / L
    return (dtype(scale) * build_wavenumbers(D, N, dtype)).astype(dtype)


def build_derivative_operator(D, L, N, dtype=np.float32):
    """exponax/_spectral.py:86-115."""
    return (1j * build_scaled_wavenumbers(D, L, N, dtype)).astype(_cdtype(dtype))


def build_laplace_operator(derivative_operator, order=2):
    """exponax/_spectral.py:118-155."""
    if order % 2 != 0:
        raise ValueError("Order must be even.")
    if order == 0:
        return np.ones((1, *derivative_operator.shape[1:]), dtype=derivative_operator.dtype)
    return np.sum(derivative_operator**order, axis=0, keepdims=True)


def build_gradient_inner_product_operator(derivative_operator, velocity, order=1):
    """exponax/_spectral.py:158-209."""
    if order % 2 != 1:
        raise ValueError("Order must be odd.")
    velocity = np.asarray(velocity, dtype=derivative_operator.real.dtype)
    if velocity.shape != (derivative_operator.shape[0],):
        raise ValueError(
            f"Expected velocity shape to be {derivative_operator.shape[0]}, got {velocity.shape}."
        )
    op = np.einsum("i,i...->...", velocity, derivative_operator**order)
    return op[None, ...].astype(derivative_operator.dtype)


def spatial_shape(D, N):
    """exponax/_spectral.py:230-251."""
    return (N,) * D


def wavenumber_shape(D, N):
    """exponax/_spectral.py:254-275."""
    return (N,) * (D - 1) + (N // 2 + 1,)


def low_pass_filter_mask(D, N, *, cutoff, axis_separate=True, dtype=np.float32):
    """exponax/_spectral.py:278-342; `cutoff` compared in the working precision."""
    wn = build_wavenumbers(D, N, dtype)
    cut = dtype(cutoff)
    if axis_separate:
        mask = True
        for g in wn:
            mask = mask & (np.abs(g) <= cut)
    else:
        mask = np.linalg.norm(wn, axis=0) <= cut
    return mask[None, ...]


def build_scaling_array(D, N, *, mode, dtype=np.float32):
    """exponax/_spectral.py:403-527."""
    den = {"norm_compensation": (1, 1), "reconstruction": (2, 1), "coef_extraction": (2, 2)}
    if mode not in den:
        raise ValueError("Invalid mode.")
    rden, oden = den[mode]
    right_wn = np.fft.rfftfreq(N, 1.0 / N)
    other_wn = np.fft.fftfreq(N, 1.0 / N)
    right = np.where(right_wn == 0, N, N / rden)
    other = np.where(other_wn == 0, N, N / oden)
    if N % 2 == 0:
        right = np.where(right_wn == N // 2, N, right)
        other = np.where(other_wn == -N // 2, N, other)
    lst = [other] * (D - 1) + [right]
    return np.prod(np.stack(np.meshgrid(*lst, indexing="ij")), axis=0, keepdims=True).astype(dtype)


def fft(field, *, num_spatial_dims=None):
    """exponax/_spectral.py:614-656: rfftn over the last D axes, unnormalised."""
    if num_spatial_dims is None:
        num_spatial_dims = field.ndim - 1
    axes = tuple(range(-num_spatial_dims, 0))
    if _sfft is not None:
        return _sfft.rfftn(field, axes=axes, workers=_WORKERS)
    return np.fft.rfftn(field, axes=axes)


def ifft(field_hat, *, num_spatial_dims=None, num_points=None):
    """exponax/_spectral.py:659-721: irfftn(s=(N,)*D), scaled by 1/N^D."""
    if num_spatial_dims is None:
        num_spatial_dims = field_hat.ndim - 1
    if num_points is None:
        if num_spatial_dims >= 2:
            num_points = field_hat.shape[-2]
        else:
            raise ValueError("num_points must be provided if num_spatial_dims == 1.")
    axes = tuple(range(-num_spatial_dims, 0))
    s = spatial_shape(num_spatial_dims, num_points)
    if _sfft is not None:
        return _sfft.irfftn(field_hat, s=s, axes=axes, workers=_WORKERS)
    return np.fft.irfftn(field_hat, s=s, axes=axes)


def make_grid(D, L, N, dtype=np.float32):
    """exponax/_utils.py:11-66 (full=False, zero_centered=False)."""
    g = np.linspace(0, L, N, endpoint=False).astype(dtype)
    return np.stack(np.meshgrid(*([g] * D), indexing="ij"))


def derivative(field, domain_extent, *, order=1):
    """exponax/_spectral.py:724-792."""
    D, N = field.ndim - 1, field.shape[-1]
    dop = build_derivative_operator(D, domain_extent, N, field.dtype.type) ** order
    fh = fft(field, num_spatial_dims=D)
    dh = dop * fh if field.shape[0] == 1 else fh[:, None] * dop[None]
    return ifft(dh, num_spatial_dims=D, num_points=N).astype(field.dtype)


def wrap_bc(u):
    """exponax/_utils.py:69-89."""
    return np.pad(u, ((0, 0),) + ((0, 1),) * (u.ndim - 1), mode="wrap")


def get_spectrum(state, *, power=True, radial_binning="sum"):
    """exponax/_spectral.py:866-1030: power / amplitude spectrum (C, N//2+1), radially binned for D > 1
    with the reference's bucket masks  k - dk/2 <= |k| < k + dk/2  evaluated in the state's precision."""
    D = state.ndim - 1
    N = state.shape[-1]
    dt = state.dtype.type
    a = np.abs(fft(state, num_spatial_dims=D))
    magnitude = a / build_scaling_array(D, N, mode="reconstruction", dtype=dt)
    if power:
        quantity = dt(0.5) * magnitude * (a / build_scaling_array(D, N, mode="norm_compensation", dtype=dt))
    else:
        quantity = magnitude
    if D == 1:
        return quantity
    wn = build_wavenumbers(D, N, dtype=dt)
    wn1 = build_wavenumbers(1, N, dtype=dt)[0]
    norm = np.sqrt(np.sum(wn * wn, axis=0)).astype(dt)           # jnp.linalg.norm(..., axis=0)
    dk = wn1[1] - wn1[0]
    out = np.empty((state.shape[0], wn1.shape[0]), dt)
    for i, k in enumerate(wn1):
        mask = (norm >= k - dk / 2) & (norm < k + dk / 2)
        for c in range(state.shape[0]):
            vals = quantity[c][mask]
            if radial_binning == "average":
                out[c, i] = vals.mean(dtype=dt) if vals.size else np.nan
            else:
                out[c, i] = vals.sum(dtype=dt)
    return out


def stack_sub_trajectories(trj, sub_len):
    """exponax/_utils.py:257-313 for one array: windows trj[i : i + sub_len], i = 0 .. T - sub_len."""
    n = trj.shape[0]
    if sub_len > n:
        raise ValueError("n must be smaller than or equal to the number of time steps in trj")
    return np.stack([trj[i:i + sub_len] for i in range(n - sub_len + 1)])


def get_spectrum_1d(u, dtype=np.float32):
    """Amplitude spectrum |u_hat|/scaling for 1-D states (test harness only;
    follows exponax/_spectral.py:866-1030 for D=1, power=False)."""
    N = u.shape[-1]
    uh = fft(u, num_spatial_dims=1)
    return np.abs(uh) / build_scaling_array(1, N, mode="reconstruction", dtype=dtype)


# --------------------------------------------------------------------------
# nonlin_fun/
# --------------------------------------------------------------------------
class BaseNonlinearFun:
    """exponax/nonlin_fun/_base.py:9-159."""

    def __init__(self, D, N, *, dealiasing_fraction=None, dtype=np.float32):
        self.num_spatial_dims = D
        self.num_points = N
        self.dtype = dtype
        if dealiasing_fraction is None:
            self.dealiasing_mask = None
        else:
            nyquist_mode = N // 2 + 1
            highest_resolved_mode = nyquist_mode - 1
            start_of_aliased_modes = dealiasing_fraction * highest_resolved_mode
            self.dealiasing_mask = low_pass_filter_mask(
                D, N, cutoff=start_of_aliased_modes - 1, dtype=dtype
            )

    def dealias(self, u_hat):
        if self.dealiasing_mask is None:
            raise ValueError("Nonlinear function was set up without dealiasing")
        return self.dealiasing_mask * u_hat

    def fft(self, u):
        u_hat = fft(u, num_spatial_dims=self.num_spatial_dims)
        if self.dealiasing_mask is not None:
            u_hat = self.dealiasing_mask * u_hat
        return u_hat

    def ifft(self, u_hat):
        if self.dealiasing_mask is not None:
            u_hat = self.dealiasing_mask * u_hat
        return ifft(u_hat, num_spatial_dims=self.num_spatial_dims, num_points=self.num_points)


class ZeroNonlinearFun(BaseNonlinearFun):
    """exponax/nonlin_fun/_zero.py:34-38."""

    def __init__(self, D, N, dtype=np.float32):
        super().__init__(D, N, dtype=dtype)

    def __call__(self, u_hat):
        return np.zeros_like(u_hat)


class ConvectionNonlinearFun(BaseNonlinearFun):
    """exponax/nonlin_fun/_convection.py:104-245."""

    def __init__(self, D, N, *, derivative_operator, dealiasing_fraction=2 / 3, scale=1.0,
                 single_channel=False, conservative=False, dtype=np.float32):
        self.derivative_operator = derivative_operator
        self.scale = scale
        self.single_channel = single_channel
        self.conservative = conservative
        super().__init__(D, N, dealiasing_fraction=dealiasing_fraction, dtype=dtype)

    def __call__(self, u_hat):
        t = self.dtype
        dop = self.derivative_operator
        if self.single_channel:
            if self.conservative:  # :192-205
                u = self.ifft(u_hat)
                u_square_hat = self.fft(u**2)
                sum_d = np.sum(dop, axis=0, keepdims=True)
                return -t(self.scale) * (t(0.5) * sum_d * u_square_hat)
            u = self.ifft(u_hat)  # :207-217
            nabla_u = self.ifft(dop * u_hat)
            conv_u = np.sum(u * nabla_u, axis=0, keepdims=True)
            return -t(self.scale) * self.fft(conv_u)
        if u_hat.shape[0] != self.num_spatial_dims:
            raise ValueError(
                "Number of channels in u_hat should match number of spatial dimensions"
            )
        if self.conservative:  # :140-163
            u = self.ifft(u_hat)
            outer = u[None, :] * u[:, None]
            outer_hat = self.fft(outer)
            conv = t(0.5) * np.sum(dop[None, :] * outer_hat, axis=1)
            return -t(self.scale) * conv
        u = self.ifft(u_hat)  # :165-190
        nabla_u = self.ifft(dop[None, :] * u_hat[:, None])
        conv_u = np.sum(u[None, :] * nabla_u, axis=1)
        return -t(self.scale) * self.fft(conv_u)


class GradientNormNonlinearFun(BaseNonlinearFun):
    """exponax/nonlin_fun/_gradient_norm.py:78-101."""

    def __init__(self, D, N, *, derivative_operator, dealiasing_fraction, zero_mode_fix=True,
                 scale=1.0, dtype=np.float32):
        super().__init__(D, N, dealiasing_fraction=dealiasing_fraction, dtype=dtype)
        self.derivative_operator = derivative_operator
        self.zero_mode_fix = zero_mode_fix
        self.scale = scale

    def __call__(self, u_hat):
        t = self.dtype
        g_hat = self.derivative_operator[None, :] * u_hat[:, None]
        g = self.ifft(g_hat)
        gn = np.sum(g**2, axis=1)
        if self.zero_mode_fix:
            D = self.num_spatial_dims
            gn = gn - np.mean(gn, axis=tuple(range(-D, 0)), keepdims=True)
        return -t(self.scale) * (t(0.5) * self.fft(gn))


class PolynomialNonlinearFun(BaseNonlinearFun):
    """exponax/nonlin_fun/_polynomial.py:64-76."""

    def __init__(self, D, N, *, dealiasing_fraction, coefficients, dtype=np.float32):
        super().__init__(D, N, dealiasing_fraction=dealiasing_fraction, dtype=dtype)
        self.coefficients = coefficients

    def __call__(self, u_hat):
        t = self.dtype
        u = self.ifft(u_hat)
        u_power = np.ones_like(u)
        u_nonlin = np.zeros_like(u)
        for c in self.coefficients:
            u_nonlin = u_nonlin + t(c) * u_power
            u_power = u_power * u
        return self.fft(u_nonlin)


class VorticityConvection2d(BaseNonlinearFun):
    """exponax/nonlin_fun/_vorticity_convection.py:52-99."""

    def __init__(self, D, N, *, convection_scale=1.0, derivative_operator, dealiasing_fraction,
                 dtype=np.float32):
        if D != 2:
            raise ValueError(f"Expected num_spatial_dims = 2, got {D}.")
        super().__init__(D, N, dealiasing_fraction=dealiasing_fraction, dtype=dtype)
        self.convection_scale = convection_scale
        self.derivative_operator = derivative_operator
        lap = build_laplace_operator(derivative_operator, order=2)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.inv_laplacian = np.where(lap == 0, 1.0, 1 / lap).astype(lap.dtype)

    def __call__(self, u_hat):
        t = self.dtype
        dop = self.derivative_operator
        psi = self.inv_laplacian * u_hat
        u = self.ifft(+dop[1:2] * psi)
        v = self.ifft(-dop[0:1] * psi)
        wx = self.ifft(dop[0:1] * u_hat)
        wy = self.ifft(dop[1:2] * u_hat)
        conv_hat = self.fft(u * wx + v * wy)
        return -t(self.convection_scale) * conv_hat


class VorticityConvection2dKolmogorov(VorticityConvection2d):
    """exponax/nonlin_fun/_vorticity_convection.py:102-182."""

    def __init__(self, D, N, *, convection_scale=1.0, injection_mode=4, injection_scale=1.0,
                 derivative_operator, dealiasing_fraction, dtype=np.float32):
        super().__init__(D, N, convection_scale=convection_scale,
                         derivative_operator=derivative_operator,
                         dealiasing_fraction=dealiasing_fraction, dtype=dtype)
        wn = build_wavenumbers(D, N, dtype)
        m = (wn[0] == 0) & (wn[1] == injection_mode)
        self.injection = np.where(
            m,
            dtype(-injection_mode * injection_scale)
            * build_scaling_array(D, N, mode="coef_extraction", dtype=dtype),
            dtype(0.0),
        ).astype(dtype)

    def __call__(self, u_hat):
        return super().__call__(u_hat) + self.injection


def _cross_product_3d(a, b):
    """exponax/nonlin_fun/_projected_convection.py:9-16."""
    return np.stack([
        a[1] * b[2] - a[2] * b[1],
        a[2] * b[0] - a[0] * b[2],
        a[0] * b[1] - a[1] * b[0],
    ], axis=0)


class Leray(BaseNonlinearFun):
    """exponax/nonlin_fun/_leray.py:84-136 (no dealiasing mask of its own)."""

    def __init__(self, D, N, *, derivative_operator, order=2, dtype=np.float32):
        super().__init__(D, N, dtype=dtype)
        lap = build_laplace_operator(derivative_operator, order=order)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.inv_laplacian = np.where(lap != 0, 1.0 / lap, 0.0).astype(lap.dtype)
        self.derivative_operator = derivative_operator

    def __call__(self, u_hat):
        div = np.sum(self.derivative_operator * u_hat, axis=0, keepdims=True)
        p = -self.inv_laplacian * div
        return u_hat + self.derivative_operator * p


class ProjectedConvection3d(BaseNonlinearFun):
    """exponax/nonlin_fun/_projected_convection.py:19-136."""

    def __init__(self, D, N, *, derivative_operator, dealiasing_fraction=2 / 3, dtype=np.float32):
        if D != 3:
            raise ValueError("ProjectedConvection3d only supports 3 spatial dimensions.")
        super().__init__(D, N, dealiasing_fraction=dealiasing_fraction, dtype=dtype)
        self.derivative_operator = derivative_operator
        self.leray_projection = Leray(D, N, derivative_operator=derivative_operator, dtype=dtype)

    def __call__(self, u_hat):
        curl_hat = _cross_product_3d(self.derivative_operator, u_hat)
        curl = self.ifft(curl_hat)
        vel = self.ifft(u_hat)
        conv_hat = self.fft(_cross_product_3d(vel, curl))
        return self.leray_projection(conv_hat)


class ProjectedConvection3dKolmogorov(ProjectedConvection3d):
    """exponax/nonlin_fun/_projected_convection.py:139-226."""

    def __init__(self, D, N, *, injection_mode=4, injection_scale=1.0, derivative_operator,
                 dealiasing_fraction, dtype=np.float32):
        super().__init__(D, N, derivative_operator=derivative_operator,
                         dealiasing_fraction=dealiasing_fraction, dtype=dtype)
        wn = build_wavenumbers(D, N, dtype)
        m = (wn[0] == 0) & (wn[1] == injection_mode) & (wn[2] == 0)
        single = np.where(
            m[None],
            dtype(injection_scale) * build_scaling_array(D, N, mode="coef_extraction", dtype=dtype),
            dtype(0.0),
        ).astype(dtype)
        zeros = np.zeros_like(single)
        self.injection = np.concatenate([single, zeros, zeros], axis=0)

    def __call__(self, u_hat):
        return super().__call__(u_hat) + self.injection


class GeneralNonlinearFun(BaseNonlinearFun):
    """exponax/nonlin_fun/_general_nonlinear.py:21-119."""

    def __init__(self, D, N, *, derivative_operator, dealiasing_fraction,
                 scale_list=(0.0, -1.0, 0.0), zero_mode_fix=True, dtype=np.float32):
        if len(scale_list) != 3:
            raise ValueError("The scale list must have exactly 3 elements")
        self.square = PolynomialNonlinearFun(
            D, N, dealiasing_fraction=dealiasing_fraction,
            coefficients=[0.0, 0.0, scale_list[0]], dtype=dtype)
        self.convection = ConvectionNonlinearFun(
            D, N, derivative_operator=derivative_operator,
            dealiasing_fraction=dealiasing_fraction, scale=-scale_list[1],
            single_channel=True, conservative=True, dtype=dtype)
        self.gradient_norm = GradientNormNonlinearFun(
            D, N, derivative_operator=derivative_operator,
            dealiasing_fraction=dealiasing_fraction, scale=-scale_list[2],
            zero_mode_fix=zero_mode_fix, dtype=dtype)
        super().__init__(D, N, dealiasing_fraction=dealiasing_fraction, dtype=dtype)

    def __call__(self, u_hat):
        return self.square(u_hat) + self.convection(u_hat) + self.gradient_norm(u_hat)


# --------------------------------------------------------------------------
# etdrk/
# --------------------------------------------------------------------------
def roots_of_unity(M, dtype=np.float32):
    """exponax/etdrk/_utils.py:9-23."""
    return np.exp(2j * np.pi * (np.arange(1, M + 1) - 0.5) / M).astype(_cdtype(dtype))


class BaseETDRK:
    """exponax/etdrk/_base_etdrk.py:15-62."""

    def __init__(self, dt, linear_operator):
        self.dt = dt
        self.cd = linear_operator.dtype
        self.rd = linear_operator.real.dtype.type
        self._exp_term = np.exp(self.rd(dt) * linear_operator).astype(self.cd)

    def _contour(self, linear_operator, fns, num_circle_points, circle_radius):
        """Shared contour mean: sum over roots of Re[f(lr)], then /M, then *dt.
        Accumulation order follows the reference's lax.scan (root by root)."""
        rd, cd = self.rd, self.cd
        roots = roots_of_unity(num_circle_points, rd)
        L_dt = (linear_operator * rd(self.dt)).astype(cd)
        accs = [np.zeros_like(L_dt.real) for _ in fns]
        for root in roots:
            lr = (rd(circle_radius) * root + L_dt).astype(cd)
            e = np.exp(lr)
            eh = np.exp(lr / rd(2))
            for i, f in enumerate(fns):
                accs[i] = accs[i] + f(lr, e, eh).real.astype(rd)
        return [rd(self.dt) * (a / rd(num_circle_points)) for a in accs]


class ETDRK0(BaseETDRK):
    """exponax/etdrk/_etdrk_0.py:30-34."""

    def step_fourier(self, u_hat):
        return self._exp_term * u_hat


class ETDRK1(BaseETDRK):
    """exponax/etdrk/_etdrk_1.py:63-82."""

    def __init__(self, dt, linear_operator, nonlinear_fun, *, num_circle_points=16, circle_radius=1.0):
        super().__init__(dt, linear_operator)
        self._nonlinear_fun = nonlinear_fun
        (self._coef_1,) = self._contour(
            linear_operator, [lambda lr, e, eh: (e - 1) / lr], num_circle_points, circle_radius)

    def step_fourier(self, u_hat):
        return self._exp_term * u_hat + self._coef_1 * self._nonlinear_fun(u_hat)


class ETDRK2(BaseETDRK):
    """exponax/etdrk/_etdrk_2.py:71-102."""

    def __init__(self, dt, linear_operator, nonlinear_fun, *, num_circle_points=16, circle_radius=1.0):
        super().__init__(dt, linear_operator)
        self._nonlinear_fun = nonlinear_fun
        self._coef_1, self._coef_2 = self._contour(
            linear_operator,
            [lambda lr, e, eh: (e - 1) / lr, lambda lr, e, eh: (e - 1 - lr) / lr**2],
            num_circle_points, circle_radius)

    def step_fourier(self, u_hat):
        n0 = self._nonlinear_fun(u_hat)
        a = self._exp_term * u_hat + self._coef_1 * n0
        n1 = self._nonlinear_fun(a)
        return a + self._coef_2 * (n1 - n0)


class ETDRK3(BaseETDRK):
    """exponax/etdrk/_etdrk_3.py:154-212 (code, not docstring, is the spec)."""

    def __init__(self, dt, linear_operator, nonlinear_fun, *, num_circle_points=16, circle_radius=1.0):
        super().__init__(dt, linear_operator)
        self._nonlinear_fun = nonlinear_fun
        self._half_exp_term = np.exp(self.rd(0.5) * self.rd(dt) * linear_operator).astype(self.cd)
        (self._coef_1, self._coef_2, self._coef_3, self._coef_4, self._coef_5) = self._contour(
            linear_operator,
            [
                lambda lr, e, eh: (eh - 1) / lr,
                lambda lr, e, eh: (e - 1) / lr,
                lambda lr, e, eh: (-4 - lr + e * (4 - 3 * lr + lr**2)) / lr**3,
                lambda lr, e, eh: (4.0 * (2.0 + lr + e * (-2 + lr))) / lr**3,
                lambda lr, e, eh: (-4 - 3 * lr - lr**2 + e * (4 - lr)) / lr**3,
            ],
            num_circle_points, circle_radius)

    def step_fourier(self, u_hat):
        n0 = self._nonlinear_fun(u_hat)
        a = self._half_exp_term * u_hat + self._coef_1 * n0
        n1 = self._nonlinear_fun(a)
        b = self._exp_term * u_hat + self._coef_2 * (2 * n1 - n0)
        n2 = self._nonlinear_fun(b)
        return self._exp_term * u_hat + self._coef_3 * n0 + self._coef_4 * n1 + self._coef_5 * n2


class ETDRK4(BaseETDRK):
    """exponax/etdrk/_etdrk_4.py:168-224 (code, not docstring, is the spec)."""

    def __init__(self, dt, linear_operator, nonlinear_fun, *, num_circle_points=16, circle_radius=1.0):
        super().__init__(dt, linear_operator)
        self._nonlinear_fun = nonlinear_fun
        self._half_exp_term = np.exp(self.rd(0.5) * self.rd(dt) * linear_operator).astype(self.cd)
        (self._coef_1, self._coef_4, self._coef_5, self._coef_6) = self._contour(
            linear_operator,
            [
                lambda lr, e, eh: (eh - 1) / lr,
                lambda lr, e, eh: (-4 - lr + e * (4 - 3 * lr + lr**2)) / lr**3,
                lambda lr, e, eh: (2 + lr + e * (-2 + lr)) / lr**3,
                lambda lr, e, eh: (-4 - 3 * lr - lr**2 + e * (4 - lr)) / lr**3,
            ],
            num_circle_points, circle_radius)
        self._coef_2 = self._coef_1
        self._coef_3 = self._coef_1

    def step_fourier(self, u_hat):
        n0 = self._nonlinear_fun(u_hat)
        a = self._half_exp_term * u_hat + self._coef_1 * n0
        n1 = self._nonlinear_fun(a)
        b = self._half_exp_term * u_hat + self._coef_2 * n1
        n2 = self._nonlinear_fun(b)
        c = self._half_exp_term * a + self._coef_3 * (2 * n2 - n0)
        n3 = self._nonlinear_fun(c)
        return (self._exp_term * u_hat + self._coef_4 * n0
                + self._coef_5 * 2 * (n1 + n2) + self._coef_6 * n3)


# --------------------------------------------------------------------------
# _base_stepper.py / _repeated_stepper.py / _utils.py
# --------------------------------------------------------------------------
class BaseStepper:
    """exponax/_base_stepper.py:16-271."""

    def __init__(self, D, L, N, dt, *, num_channels, order, num_circle_points=16,
                 circle_radius=1.0, dtype=np.float32):
        self.num_spatial_dims = D
        self.domain_extent = L
        self.num_points = N
        self.dt = dt
        self.num_channels = num_channels
        self.dx = L / N
        self.dtype = dtype
        dop = build_derivative_operator(D, L, N, dtype)
        lin = np.asarray(self._build_linear_operator(dop)).astype(_cdtype(dtype))
        s1 = (1,) + wavenumber_shape(D, N)
        sc = (num_channels,) + wavenumber_shape(D, N)
        if lin.shape not in (s1, sc):
            raise ValueError(f"Expected linear operator to have shape {s1} or {sc}, got {lin.shape}.")
        self.linear_operator = lin
        nl = self._build_nonlinear_fun(dop)
        kw = dict(num_circle_points=num_circle_points, circle_radius=circle_radius)
        if order == 0:
            self._integrator = ETDRK0(dt, lin)
        elif order == 1:
            self._integrator = ETDRK1(dt, lin, nl, **kw)
        elif order == 2:
            self._integrator = ETDRK2(dt, lin, nl, **kw)
        elif order == 3:
            self._integrator = ETDRK3(dt, lin, nl, **kw)
        elif order == 4:
            self._integrator = ETDRK4(dt, lin, nl, **kw)
        else:
            raise NotImplementedError(f"Order {order} not implemented.")

    def step(self, u):
        D = self.num_spatial_dims
        u_hat = fft(u, num_spatial_dims=D)
        u_next_hat = self.step_fourier(u_hat)
        return ifft(u_next_hat, num_spatial_dims=D, num_points=self.num_points)

    def step_fourier(self, u_hat):
        return self._integrator.step_fourier(u_hat)

    def __call__(self, u):
        expected = (self.num_channels,) + spatial_shape(self.num_spatial_dims, self.num_points)
        if u.shape != expected:
            raise ValueError(
                f"Expected shape {expected}, got {u.shape}. For batched operation use "
                "`jax.vmap` on this function."
            )
        return self.step(u)


def repeat(stepper_fn, n):
    """exponax/_utils.py:189-254 (no-aux variant)."""
    def fn(u0):
        u = u0
        for _ in range(n):
            u = stepper_fn(u)
        return u
    return fn


def rollout(stepper_fn, n, *, include_init=False):
    """exponax/_utils.py:92-186 (no-aux variant)."""
    def fn(u0):
        u = u0
        trj = []
        for _ in range(n):
            u = stepper_fn(u)
            trj.append(u)
        trj = np.stack(trj) if trj else np.zeros((0,) + u0.shape, u0.dtype)
        if include_init:
            trj = np.concatenate([u0[None], trj], axis=0)
        return trj
    return fn


class RepeatedStepper:
    """exponax/_repeated_stepper.py:20-139: k sub-steps with a spectral carry."""

    def __init__(self, stepper, num_sub_steps):
        self.stepper = stepper
        self.num_sub_steps = num_sub_steps
        self.dt = stepper.dt * num_sub_steps
        self.num_spatial_dims = stepper.num_spatial_dims
        self.domain_extent = stepper.domain_extent
        self.num_points = stepper.num_points
        self.num_channels = stepper.num_channels
        self.dx = stepper.dx

    def step_fourier(self, u_hat):
        return repeat(self.stepper.step_fourier, self.num_sub_steps)(u_hat)

    def step(self, u):
        D = self.num_spatial_dims
        u_hat = fft(u, num_spatial_dims=D)
        return ifft(self.step_fourier(u_hat), num_spatial_dims=D, num_points=self.num_points)

    def __call__(self, u):
        expected = (self.num_channels,) + spatial_shape(self.num_spatial_dims, self.num_points)
        if u.shape != expected:
            raise ValueError(f"Expected shape {expected}, got {u.shape}.")
        return self.step(u)


def batched(stepper_fn):
    """Stand-in for jax.vmap over the leading axis (exponax/_base_stepper.py:260-262).
    The FFT helpers above transform the LAST D axes only, so a stepper built on
    them already broadcasts over leading axes; this wrapper merely skips the
    un-batched shape check of `__call__`."""
    step = getattr(stepper_fn, "step", stepper_fn)
    return step


# --------------------------------------------------------------------------
# stepper/ constructors (only the physics -> (L_hat, N) mapping)
# --------------------------------------------------------------------------
class Burgers(BaseStepper):
    """exponax/stepper/_burgers.py:15-155."""

    def __init__(self, D, L, N, dt, *, diffusivity=0.1, convection_scale=1.0, single_channel=False,
                 conservative=False, order=2, dealiasing_fraction=2 / 3, num_circle_points=16,
                 circle_radius=1.0, dtype=np.float32):
        self.diffusivity = diffusivity
        self.convection_scale = convection_scale
        self.single_channel = single_channel
        self.conservative = conservative
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1 if single_channel else D, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        return self.dtype(self.diffusivity) * build_laplace_operator(dop)

    def _build_nonlinear_fun(self, dop):
        return ConvectionNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=dop,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.convection_scale,
            single_channel=self.single_channel, conservative=self.conservative, dtype=self.dtype)


class KuramotoSivashinsky(BaseStepper):
    """exponax/stepper/_kuramoto_sivashinsky.py:8-168 (combustion format)."""

    def __init__(self, D, L, N, dt, *, gradient_norm_scale=1.0, second_order_scale=1.0,
                 fourth_order_scale=1.0, dealiasing_fraction=2 / 3, order=2, num_circle_points=16,
                 circle_radius=1.0, dtype=np.float32):
        self.gradient_norm_scale = gradient_norm_scale
        self.second_order_scale = second_order_scale
        self.fourth_order_scale = fourth_order_scale
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        return (-t(self.second_order_scale) * build_laplace_operator(dop, order=2)
                - t(self.fourth_order_scale) * build_laplace_operator(dop, order=4))

    def _build_nonlinear_fun(self, dop):
        return GradientNormNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=dop,
            dealiasing_fraction=self.dealiasing_fraction, zero_mode_fix=True,
            scale=self.gradient_norm_scale, dtype=self.dtype)


class KuramotoSivashinskyConservative(BaseStepper):
    """exponax/stepper/_kuramoto_sivashinsky.py:171-315."""

    def __init__(self, D, L, N, dt, *, convection_scale=1.0, second_order_scale=1.0,
                 fourth_order_scale=1.0, single_channel=False, conservative=True,
                 dealiasing_fraction=2 / 3, order=2, num_circle_points=16, circle_radius=1.0,
                 dtype=np.float32):
        self.convection_scale = convection_scale
        self.second_order_scale = second_order_scale
        self.fourth_order_scale = fourth_order_scale
        self.single_channel = single_channel
        self.conservative = conservative
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1 if single_channel else D, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        return (-t(self.second_order_scale) * build_laplace_operator(dop, order=2)
                - t(self.fourth_order_scale) * build_laplace_operator(dop, order=4))

    def _build_nonlinear_fun(self, dop):
        return ConvectionNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=dop,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.convection_scale,
            single_channel=self.single_channel, conservative=self.conservative, dtype=self.dtype)


class KortewegDeVries(BaseStepper):
    """exponax/stepper/_korteweg_de_vries.py:16-216."""

    def __init__(self, D, L, N, dt, *, convection_scale=-6.0, diffusivity=0.0, dispersivity=1.0,
                 hyper_diffusivity=0.01, advect_over_diffuse=False, diffuse_over_diffuse=False,
                 single_channel=False, conservative=False, order=2, dealiasing_fraction=2 / 3,
                 num_circle_points=16, circle_radius=1.0, dtype=np.float32):
        self.convection_scale = convection_scale
        self.diffusivity = diffusivity
        self.dispersivity = dispersivity
        self.hyper_diffusivity = hyper_diffusivity
        self.advect_over_diffuse = advect_over_diffuse
        self.diffuse_over_diffuse = diffuse_over_diffuse
        self.single_channel = single_channel
        self.conservative = conservative
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1 if single_channel else D, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        vel = t(self.dispersivity) * np.ones(self.num_spatial_dims, dtype=t)
        lap = build_laplace_operator(dop, order=2)
        diffusion = t(self.diffusivity) * lap
        if self.advect_over_diffuse:
            dispersion = -build_gradient_inner_product_operator(dop, vel, order=1) * lap
        else:
            dispersion = -build_gradient_inner_product_operator(dop, vel, order=3)
        if self.diffuse_over_diffuse:
            hyper = -t(self.hyper_diffusivity) * lap * lap
        else:
            hyper = -t(self.hyper_diffusivity) * build_laplace_operator(dop, order=4)
        return diffusion + dispersion + hyper

    def _build_nonlinear_fun(self, dop):
        return ConvectionNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=dop,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.convection_scale,
            single_channel=self.single_channel, conservative=self.conservative, dtype=self.dtype)


class _NSBase(BaseStepper):
    def _build_linear_operator(self, dop):
        t = self.dtype
        return (t(self.diffusivity) * build_laplace_operator(dop, order=2)
                + t(self.drag) * build_laplace_operator(dop, order=0))


class NavierStokesVorticity(_NSBase):
    """exponax/stepper/_navier_stokes.py:13-153."""

    def __init__(self, D, L, N, dt, *, diffusivity=0.01, vorticity_convection_scale=1.0, drag=0.0,
                 order=2, dealiasing_fraction=2 / 3, num_circle_points=16, circle_radius=1.0,
                 dtype=np.float32):
        if D != 2:
            raise ValueError(f"Expected num_spatial_dims = 2, got {D}. For 3D, use NavierStokesVelocity instead.")
        self.diffusivity = diffusivity
        self.vorticity_convection_scale = vorticity_convection_scale
        self.drag = drag
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_nonlinear_fun(self, dop):
        return VorticityConvection2d(
            self.num_spatial_dims, self.num_points,
            convection_scale=self.vorticity_convection_scale, derivative_operator=dop,
            dealiasing_fraction=self.dealiasing_fraction, dtype=self.dtype)


class KolmogorovFlowVorticity(_NSBase):
    """exponax/stepper/_navier_stokes.py:156-327."""

    def __init__(self, D, L, N, dt, *, diffusivity=0.001, convection_scale=1.0, drag=-0.1,
                 injection_mode=4, injection_scale=1.0, order=2, dealiasing_fraction=2 / 3,
                 num_circle_points=16, circle_radius=1.0, dtype=np.float32):
        if D != 2:
            raise ValueError(f"Expected num_spatial_dims = 2, got {D}. For 3D, use KolmogorovFlowVelocity instead.")
        self.diffusivity = diffusivity
        self.convection_scale = convection_scale
        self.drag = drag
        self.injection_mode = injection_mode
        self.injection_scale = injection_scale
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_nonlinear_fun(self, dop):
        return VorticityConvection2dKolmogorov(
            self.num_spatial_dims, self.num_points, convection_scale=self.convection_scale,
            injection_mode=self.injection_mode, injection_scale=self.injection_scale,
            derivative_operator=dop, dealiasing_fraction=self.dealiasing_fraction, dtype=self.dtype)


class NavierStokesVelocity(_NSBase):
    """exponax/stepper/_navier_stokes.py:330-463."""

    def __init__(self, D, L, N, dt, *, diffusivity=0.01, drag=0.0, order=2,
                 dealiasing_fraction=2 / 3, num_circle_points=16, circle_radius=1.0, dtype=np.float32):
        if D != 3:
            raise ValueError(f"Expected num_spatial_dims = 3, got {D}. For 2D, use NavierStokesVorticity instead.")
        self.diffusivity = diffusivity
        self.drag = drag
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=3, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_nonlinear_fun(self, dop):
        return ProjectedConvection3d(
            self.num_spatial_dims, self.num_points, derivative_operator=dop,
            dealiasing_fraction=self.dealiasing_fraction, dtype=self.dtype)


class KolmogorovFlowVelocity(_NSBase):
    """exponax/stepper/_navier_stokes.py:466-598."""

    def __init__(self, D, L, N, dt, *, diffusivity=0.01, drag=0.0, injection_mode=4,
                 injection_scale=1.0, order=2, dealiasing_fraction=2 / 3, num_circle_points=16,
                 circle_radius=1.0, dtype=np.float32):
        if D != 3:
            raise ValueError(f"Expected num_spatial_dims = 3, got {D}. For 2D, use KolmogorovFlowVorticity instead.")
        self.diffusivity = diffusivity
        self.drag = drag
        self.injection_mode = injection_mode
        self.injection_scale = injection_scale
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=3, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_nonlinear_fun(self, dop):
        return ProjectedConvection3dKolmogorov(
            self.num_spatial_dims, self.num_points, injection_mode=self.injection_mode,
            injection_scale=self.injection_scale, derivative_operator=dop,
            dealiasing_fraction=self.dealiasing_fraction, dtype=self.dtype)


class _LinearBase(BaseStepper):
    def __init__(self, D, L, N, dt, dtype=np.float32):
        super().__init__(D, L, N, dt, num_channels=1, order=0, dtype=dtype)

    def _build_nonlinear_fun(self, dop):
        return ZeroNonlinearFun(self.num_spatial_dims, self.num_points, dtype=self.dtype)


def _vec(x, D, dtype):
    if isinstance(x, (int, float)):
        return np.ones(D, dtype=dtype) * dtype(x)
    return np.asarray(x, dtype=dtype)


def _mat(x, D, dtype):
    if isinstance(x, (int, float)):
        return np.diag(np.ones(D, dtype=dtype)) * dtype(x)
    x = np.asarray(x, dtype=dtype)
    return np.diag(x) if x.ndim == 1 else x


class Advection(_LinearBase):
    """exponax/stepper/_advection.py:13-104."""

    def __init__(self, D, L, N, dt, *, velocity=1.0, dtype=np.float32):
        self.velocity = _vec(velocity, D, dtype)
        super().__init__(D, L, N, dt, dtype)

    def _build_linear_operator(self, dop):
        return -build_gradient_inner_product_operator(dop, self.velocity, order=1)


class Diffusion(_LinearBase):
    """exponax/stepper/_diffusion.py:12-122."""

    def __init__(self, D, L, N, dt, *, diffusivity=0.01, dtype=np.float32):
        self.diffusivity = _mat(diffusivity, D, dtype)
        super().__init__(D, L, N, dt, dtype)

    def _build_linear_operator(self, dop):
        outer = dop[:, None] * dop[None, :]
        return np.einsum("ij,ij...->...", self.diffusivity, outer)[None, ...]


class AdvectionDiffusion(_LinearBase):
    """exponax/stepper/_advection_diffusion.py:13-137."""

    def __init__(self, D, L, N, dt, *, velocity=1.0, diffusivity=0.01, dtype=np.float32):
        self.velocity = _vec(velocity, D, dtype)
        self.diffusivity = _mat(diffusivity, D, dtype)
        super().__init__(D, L, N, dt, dtype)

    def _build_linear_operator(self, dop):
        outer = dop[:, None] * dop[None, :]
        diffusion = np.einsum("ij,ij...->...", self.diffusivity, outer)[None, ...]
        advection = -build_gradient_inner_product_operator(dop, self.velocity, order=1)
        return advection + diffusion


class Dispersion(_LinearBase):
    """exponax/stepper/_dispersion.py:13-126."""

    def __init__(self, D, L, N, dt, *, dispersivity=1.0, advect_on_diffusion=False, dtype=np.float32):
        self.dispersivity = _vec(dispersivity, D, dtype)
        self.advect_on_diffusion = advect_on_diffusion
        super().__init__(D, L, N, dt, dtype)

    def _build_linear_operator(self, dop):
        if self.advect_on_diffusion:
            lap = build_laplace_operator(dop)
            adv = build_gradient_inner_product_operator(dop, self.dispersivity, order=1)
            return adv * lap
        return build_gradient_inner_product_operator(dop, self.dispersivity, order=3)


class HyperDiffusion(_LinearBase):
    """exponax/stepper/_hyper_diffusion.py:8-119."""

    def __init__(self, D, L, N, dt, *, hyper_diffusivity=0.0001, diffuse_on_diffuse=False,
                 dtype=np.float32):
        self.hyper_diffusivity = hyper_diffusivity
        self.diffuse_on_diffuse = diffuse_on_diffuse
        super().__init__(D, L, N, dt, dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        if self.diffuse_on_diffuse:
            lap = build_laplace_operator(dop)
            return -t(self.hyper_diffusivity) * lap * lap
        return -t(self.hyper_diffusivity) * build_laplace_operator(dop, order=4)


class _PolyBase(BaseStepper):
    def _build_nonlinear_fun(self, dop):
        return PolynomialNonlinearFun(
            self.num_spatial_dims, self.num_points, dealiasing_fraction=self.dealiasing_fraction,
            coefficients=self._poly_coefficients(), dtype=self.dtype)


class FisherKPP(_PolyBase):
    """exponax/stepper/reaction/_fisher_kpp.py:8-129."""

    def __init__(self, D, L, N, dt, *, diffusivity=0.01, reactivity=1.0, order=2,
                 dealiasing_fraction=2 / 3, num_circle_points=16, circle_radius=1.0, dtype=np.float32):
        self.diffusivity = diffusivity
        self.reactivity = reactivity
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        return t(self.diffusivity) * build_laplace_operator(dop, order=2) + t(self.reactivity)

    def _poly_coefficients(self):
        return [0.0, 0.0, -self.reactivity]


class AllenCahn(_PolyBase):
    """exponax/stepper/reaction/_allen_cahn.py:8-128."""

    def __init__(self, D, L, N, dt, *, diffusivity=5e-3, first_order_coefficient=1.0,
                 third_order_coefficient=-1.0, order=2, dealiasing_fraction=1 / 2,
                 num_circle_points=16, circle_radius=1.0, dtype=np.float32):
        self.diffusivity = diffusivity
        self.first_order_coefficient = first_order_coefficient
        self.third_order_coefficient = third_order_coefficient
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        return t(self.diffusivity) * build_laplace_operator(dop, order=2) + t(self.first_order_coefficient)

    def _poly_coefficients(self):
        return [0.0, 0.0, 0.0, self.third_order_coefficient]


class SwiftHohenberg(_PolyBase):
    """exponax/stepper/reaction/_swift_hohenberg.py:8-128."""

    def __init__(self, D, L, N, dt, *, reactivity=0.7, critical_number=1.0,
                 polynomial_coefficients=(0.0, 0.0, 1.0, -1.0), order=2, dealiasing_fraction=1 / 2,
                 num_circle_points=16, circle_radius=1.0, dtype=np.float32):
        self.reactivity = reactivity
        self.critical_number = critical_number
        self.polynomial_coefficients = polynomial_coefficients
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        lap = build_laplace_operator(dop, order=2)
        return t(self.reactivity) - (t(self.critical_number) + lap) ** 2

    def _poly_coefficients(self):
        return self.polynomial_coefficients


# --------------------------------------------------------------------------
# synthetic initial conditions (NumPy restatements of exponax/ic; harness only)
# --------------------------------------------------------------------------
def random_truncated_fourier_series(D, N, *, cutoff=5, seed=0, max_one=False, dtype=np.float32):
    """exponax/ic/_truncated_fourier_series.py:65-100, with NumPy's PCG64 in place
    of jax.random (threefry is unavailable without JAX; values are NOT those of
    the reference PRNG, only the distribution is)."""
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal((1,) + (N,) * D).astype(dtype)
    nh = fft(noise, num_spatial_dims=D)
    mask = low_pass_filter_mask(D, N, cutoff=cutoff, dtype=dtype)
    nh = nh * mask
    nh[(0,) + (0,) * D] = 0.0
    u = ifft(nh, num_spatial_dims=D, num_points=N).astype(dtype)
    u = u - np.mean(u)
    if max_one:
        u = u / np.max(np.abs(u))
    return u.astype(dtype)


def gaussian_random_field(D, N, *, L=1.0, powerlaw_exponent=3.0, seed=0, max_one=True,
                          dtype=np.float32):
    """exponax/ic/_gaussian_random_field.py:64-93 (PCG64 instead of threefry)."""
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal((1,) + (N,) * D).astype(dtype)
    noise_hat = fft(noise, num_spatial_dims=D)
    wn = build_scaled_wavenumbers(D, L, N, dtype)
    norm = np.linalg.norm(wn, axis=0, keepdims=True)
    with np.errstate(divide="ignore"):
        amp = np.power(norm, -powerlaw_exponent / 2.0)
    amp[(0,) + (0,) * D] = 1.0
    u = ifft(noise_hat * amp, num_spatial_dims=D, num_points=N).astype(dtype)
    u = (u - np.mean(u)) / np.std(u)
    if max_one:
        u = u / np.max(np.abs(u))
    return u.astype(dtype)


# --------------------------------------------------------------------------
# stepper/reaction/_gray_scott.py, _cahn_hilliard.py
# --------------------------------------------------------------------------
class GrayScottNonlinearFun(BaseNonlinearFun):
    """exponax/stepper/reaction/_gray_scott.py:12-45 (the mask is applied twice there; idempotent)."""

    def __init__(self, D, N, *, dealiasing_fraction, feed_rate, kill_rate, dtype=np.float32):
        super().__init__(D, N, dealiasing_fraction=dealiasing_fraction, dtype=dtype)
        self.feed_rate, self.kill_rate = feed_rate, kill_rate

    def __call__(self, u_hat):
        if u_hat.shape[0] != 2:
            raise ValueError("num_channels must be 2")
        t = self.dtype
        u = self.ifft(self.dealias(u_hat))
        f, k = t(self.feed_rate), t(self.kill_rate)
        up = np.stack([f * (1 - u[0]) - u[0] * u[1] ** 2, -(f + k) * u[1] + u[0] * u[1] ** 2])
        return self.fft(up)


class GrayScott(BaseStepper):
    """exponax/stepper/reaction/_gray_scott.py:48-180."""

    def __init__(self, D, L, N, dt, *, diffusivity_1=2e-5, diffusivity_2=1e-5, feed_rate=0.04, kill_rate=0.06,
                 order=2, dealiasing_fraction=1 / 2, num_circle_points=16, circle_radius=1.0, dtype=np.float32):
        self.diffusivity_1, self.diffusivity_2 = diffusivity_1, diffusivity_2
        self.feed_rate, self.kill_rate = feed_rate, kill_rate
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=2, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        lap = build_laplace_operator(dop, order=2)
        return np.concatenate([t(self.diffusivity_1) * lap, t(self.diffusivity_2) * lap])

    def _build_nonlinear_fun(self, dop):
        return GrayScottNonlinearFun(self.num_spatial_dims, self.num_points, feed_rate=self.feed_rate,
                                     kill_rate=self.kill_rate, dealiasing_fraction=self.dealiasing_fraction,
                                     dtype=self.dtype)


class CahnHilliardNonlinearFun(BaseNonlinearFun):
    """exponax/stepper/reaction/_cahn_hilliard.py:12-37."""

    def __init__(self, D, N, *, derivative_operator, scale, dealiasing_fraction, dtype=np.float32):
        super().__init__(D, N, dealiasing_fraction=dealiasing_fraction, dtype=dtype)
        self.laplace_operator = build_laplace_operator(derivative_operator)
        self.scale = scale

    def __call__(self, u_hat):
        u = self.ifft(self.dealias(u_hat))
        up_hat = self.fft(u[0:1] ** 3)
        return self.laplace_operator * up_hat * self.dtype(self.scale)


class CahnHilliard(BaseStepper):
    """exponax/stepper/reaction/_cahn_hilliard.py:40-158."""

    def __init__(self, D, L, N, dt, *, diffusivity=1e-2, gamma=1e-3, first_order_coefficient=-1.0,
                 third_order_coefficient=1.0, order=2, dealiasing_fraction=1 / 2, num_circle_points=16,
                 circle_radius=1.0, dtype=np.float32):
        self.diffusivity, self.gamma = diffusivity, gamma
        self.first_order_coefficient, self.third_order_coefficient = first_order_coefficient, third_order_coefficient
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(D, L, N, dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius, dtype=dtype)

    def _build_linear_operator(self, dop):
        t = self.dtype
        lap = build_laplace_operator(dop, order=2)
        return t(self.diffusivity) * lap * (t(self.first_order_coefficient) - t(self.gamma) * lap)

    def _build_nonlinear_fun(self, dop):
        return CahnHilliardNonlinearFun(self.num_spatial_dims, self.num_points, derivative_operator=dop,
                                        dealiasing_fraction=self.dealiasing_fraction,
                                        scale=self.diffusivity * self.third_order_coefficient, dtype=self.dtype)


class Wave(BaseStepper):
    """exponax/stepper/_wave.py:14-197."""

    def __init__(self, D, L, N, dt, *, speed_of_sound=1.0, dtype=np.float32):
        self.speed_of_sound = speed_of_sound
        self.wavenumber_norm = np.linalg.norm(build_scaled_wavenumbers(D, L, N, dtype), axis=0,
                                              keepdims=True).astype(dtype)
        super().__init__(D, L, N, dt, num_channels=2, order=0, dtype=dtype)

    def _build_linear_operator(self, dop):
        val = 1j * self.dtype(self.speed_of_sound) * self.wavenumber_norm
        return np.concatenate((val, -val), axis=0)

    def _build_nonlinear_fun(self, dop):
        return ZeroNonlinearFun(self.num_spatial_dims, self.num_points, dtype=self.dtype)

    def step_fourier(self, u_hat):
        t = self.dtype
        c = t(self.speed_of_sound)
        kg = np.where(self.wavenumber_norm == 0, t(1.0), self.wavenumber_norm)
        s = t(1 / np.sqrt(2))
        w = 1j * c * kg * u_hat[0:1]
        waves = np.concatenate([s * (w + u_hat[1:2]), s * (w - u_hat[1:2])], axis=0)
        nxt = self._integrator.step_fourier(waves)
        w2 = s * (nxt[0:1] + nxt[1:2])
        v2 = s * (nxt[0:1] - nxt[1:2])
        out = np.concatenate([w2 / (1j * c * kg), v2], axis=0)
        dc = (0,) * self.num_spatial_dims
        out[(0,) + dc] += t(self.dt) * u_hat[(1,) + dc]
        return out


# --------------------------------------------------------------------------
# metrics/  (spatial family + correlation; consumers of saved snapshots)
# --------------------------------------------------------------------------
def spatial_aggregator(state_no_channel, *, num_spatial_dims=None, domain_extent=1.0, num_points=None,
                       inner_exponent=2.0, outer_exponent=None):
    """exponax/metrics/_spatial.py:8-83."""
    if num_spatial_dims is None:
        num_spatial_dims = state_no_channel.ndim
    if num_points is None:
        num_points = state_no_channel.shape[-1]
    if outer_exponent is None:
        outer_exponent = 1 / inner_exponent
    scale = (domain_extent / num_points) ** num_spatial_dims
    aggregated = np.sum(np.abs(state_no_channel) ** inner_exponent)
    return (scale * aggregated) ** outer_exponent


def spatial_norm(state, state_ref=None, *, mode="absolute", domain_extent=1.0, inner_exponent=2.0,
                 outer_exponent=None):
    """exponax/metrics/_spatial.py:86-196."""
    if state_ref is None:
        if mode == "normalized":
            raise ValueError("mode 'normalized' requires state_ref")
        if mode == "symmetric":
            raise ValueError("mode 'symmetric' requires state_ref")
        diff = state
    else:
        diff = state - state_ref
    agg = lambda x: np.array([spatial_aggregator(c, domain_extent=domain_extent, inner_exponent=inner_exponent,
                                                 outer_exponent=outer_exponent) for c in x])
    d = agg(diff)
    if mode == "normalized":
        d = d / agg(state_ref)
    elif mode == "symmetric":
        d = 2 * d / (agg(state) + agg(state_ref))
    return np.sum(d)


def _metric(mode, p, q):
    def fn(u_pred, u_ref=None, *, domain_extent=1.0):
        return spatial_norm(u_pred, u_ref, mode=mode, domain_extent=domain_extent, inner_exponent=p,
                            outer_exponent=q)
    return fn


MAE, nMAE, sMAE = _metric("absolute", 1.0, 1.0), _metric("normalized", 1.0, 1.0), _metric("symmetric", 1.0, 1.0)
MSE, nMSE, sMSE = _metric("absolute", 2.0, 1.0), _metric("normalized", 2.0, 1.0), _metric("symmetric", 2.0, 1.0)
RMSE, nRMSE, sRMSE = _metric("absolute", 2.0, 0.5), _metric("normalized", 2.0, 0.5), _metric("symmetric", 2.0, 0.5)


def correlation(u_pred, u_ref):
    """exponax/metrics/_correlation.py:6-60."""
    per_channel = [np.dot((a / np.linalg.norm(a)).ravel(), (b / np.linalg.norm(b)).ravel())
                   for a, b in zip(u_pred, u_ref)]
    return np.mean(per_channel)


def mean_metric(metric_fn, *args, **kwargs):
    """exponax/metrics/_utils.py:5-18."""
    return np.mean([metric_fn(*(a[i] for a in args), **kwargs) for i in range(len(args[0]))], axis=0)


# --------------------------------------------------------------------------
# ic/  (the spectral random generators, deterministic part: noise in, state out)
# --------------------------------------------------------------------------
def normalize_ic(ic, *, zero_mean=True, std_one=False, max_one=False):
    """exponax/ic/_base_ic.py:16-33."""
    if zero_mean:
        ic = ic - np.mean(ic)
    if std_one:
        ic = ic / np.std(ic)
    if max_one:
        ic = ic / np.max(np.abs(ic))
    return ic


def ic_truncated_fourier_series(noise, *, cutoff=5, offset=0.0, zero_mean=True, std_one=False, max_one=False):
    """exponax/ic/_truncated_fourier_series.py:65-100 from a given white-noise array (1, N, .., N)."""
    D, N = noise.ndim - 1, noise.shape[-1]
    nh = fft(noise, num_spatial_dims=D) * low_pass_filter_mask(D, N, cutoff=cutoff, axis_separate=True,
                                                                 dtype=noise.dtype.type)
    shp = nh.shape
    nh = nh.ravel()
    nh[0] = offset
    ic = ifft(nh.reshape(shp), num_spatial_dims=D, num_points=N).astype(noise.dtype)
    return normalize_ic(ic, zero_mean=zero_mean, std_one=std_one, max_one=max_one)


def ic_gaussian_random_field(noise, *, L=1.0, powerlaw_exponent=3.0, zero_mean=True, std_one=False, max_one=False):
    """exponax/ic/_gaussian_random_field.py:64-93 from a given white-noise array."""
    D, N = noise.ndim - 1, noise.shape[-1]
    dt = noise.dtype.type
    nh = fft(noise, num_spatial_dims=D)
    norm = np.linalg.norm(build_scaled_wavenumbers(D, L, N, dt), axis=0, keepdims=True)
    with np.errstate(divide="ignore"):
        amp = np.power(norm, dt(-powerlaw_exponent / 2.0))
    amp.ravel()[0] = 1.0
    ic = ifft(nh * amp, num_spatial_dims=D, num_points=N).astype(noise.dtype)
    return normalize_ic(ic, zero_mean=zero_mean, std_one=std_one, max_one=max_one)


def ic_diffused_noise(noise, *, L=1.0, intensity=0.001, zero_mean=True, std_one=False, max_one=False):
    """exponax/ic/_diffused_noise.py:58-77 from a given white-noise array."""
    D, N = noise.ndim - 1, noise.shape[-1]
    ic = Diffusion(D, L, N, 1.0, diffusivity=intensity, dtype=noise.dtype.type)(noise)
    return normalize_ic(ic, zero_mean=zero_mean, std_one=std_one, max_one=max_one)


def fourier_aggregator(state_no_channel, *, num_spatial_dims=None, domain_extent=1.0, num_points=None,
                       inner_exponent=2.0, outer_exponent=None, low=None, high=None, derivative_order=None):
    """exponax/metrics/_fourier.py:15-140."""
    D = state_no_channel.ndim if num_spatial_dims is None else num_spatial_dims
    N = state_no_channel.shape[-1] if num_points is None else num_points
    dt = state_no_channel.dtype.type
    if outer_exponent is None:
        outer_exponent = 1 / inner_exponent
    sh = fft(state_no_channel, num_spatial_dims=D)
    sh = np.where(np.abs(sh) < 1e-5, np.zeros_like(sh), sh)
    if low is not None or high is not None:
        low = 0 if low is None else low
        high = N // 2 + 1 if high is None else high
        low_mask = low_pass_filter_mask(D, N, cutoff=low - 1, dtype=dt)
        high_mask = low_pass_filter_mask(D, N, cutoff=high, dtype=dt)
        sh = sh * (np.invert(low_mask) & high_mask)[0]
    if derivative_order is not None:
        dop = build_derivative_operator(D, domain_extent, N, dt)
        with np.errstate(divide="ignore", invalid="ignore"):
            sd = sh[None] * np.abs(dop) ** derivative_order     # only |.| is used below: |(i k c)^o| = |k c|^o
    else:
        sd = sh[None]
    recon = build_scaling_array(D, N, mode="reconstruction", dtype=dt)[0]
    scale = (domain_extent / N) ** D
    return sum((scale * np.sum(np.abs(s) ** inner_exponent / recon)) ** outer_exponent for s in sd)


def fourier_norm(state, state_ref=None, *, mode="absolute", domain_extent=1.0, inner_exponent=2.0,
                 outer_exponent=None, low=None, high=None, derivative_order=None):
    """exponax/metrics/_fourier.py:143-235."""
    if state_ref is None:
        if mode == "normalized":
            raise ValueError("mode 'normalized' requires state_ref")
        diff = state
    else:
        diff = state - state_ref
    kw = dict(domain_extent=domain_extent, inner_exponent=inner_exponent, outer_exponent=outer_exponent, low=low,
              high=high, derivative_order=derivative_order)
    d = np.array([fourier_aggregator(c, **kw) for c in diff])
    if mode == "normalized":
        d = d / np.array([fourier_aggregator(c, **kw) for c in state_ref])
    return np.sum(d)


def _fmetric(mode, p, q):
    def fn(u_pred, u_ref=None, *, domain_extent=1.0, low=None, high=None, derivative_order=None):
        return fourier_norm(u_pred, u_ref, mode=mode, domain_extent=domain_extent, inner_exponent=p,
                            outer_exponent=q, low=low, high=high, derivative_order=derivative_order)
    return fn


fourier_MAE, fourier_nMAE = _fmetric("absolute", 1.0, 1.0), _fmetric("normalized", 1.0, 1.0)
fourier_MSE, fourier_nMSE = _fmetric("absolute", 2.0, 1.0), _fmetric("normalized", 2.0, 1.0)
fourier_RMSE, fourier_nRMSE = _fmetric("absolute", 2.0, 0.5), _fmetric("normalized", 2.0, 0.5)


def _h1(base):
    def fn(u_pred, u_ref=None, *, domain_extent=1.0, low=None, high=None):
        kw = dict(domain_extent=domain_extent, low=low, high=high)
        return base(u_pred, u_ref, derivative_order=None, **kw) + base(u_pred, u_ref, derivative_order=1, **kw)
    return fn


H1_MAE, H1_nMAE, H1_MSE = _h1(fourier_MAE), _h1(fourier_nMAE), _h1(fourier_MSE)
H1_nMSE, H1_RMSE, H1_nRMSE = _h1(fourier_nMSE), _h1(fourier_RMSE), _h1(fourier_nRMSE)
