"""
Spectral primitives.  The operator/mask builders are constructor-time host code (NumPy), as
in the reference where they run once in `BaseStepper.__init__`; `fft` / `ifft` are device ops
backed by libexb's shared-memory Stockham kernels.

Mirrors exponax/_spectral.py (function names, arguments, shapes, error messages).
"""
from __future__ import annotations

import os

from typing import Literal

import numpy as np

from . import _array as A
from . import _native as nat
from ._config import complex_dtype, real_dtype


# ---- slab context (config c5): constructors build only the LOCAL spectral slab -------------------
_SLAB = None  # (rank, nranks): 3-D spectral arrays are restricted to axis-1 indices of this rank
# Which axis-1 indices a rank owns: cyclic (r, r + P, r + 2P, ...; default) or block ([r N/P, (r+1) N/P)).  With a
# dealiasing mask the cyclic distribution gives every rank the same share of the wavenumbers inside the mask -- under
# the block distribution two of 8 ranks own nothing but dealiased modes and idle through the axis-0 passes and the
# transposes (EXB_SLAB_CYCLIC=0 restores it; all ranks must agree).
SLAB_CYCLIC = os.environ.get("EXB_SLAB_CYCLIC", "1") != "0"


def slab_indices(rank: int, nranks: int, num_points: int):
    """global axis-1 indices owned by `rank` (a slice object)."""
    n = num_points // nranks
    return slice(rank, None, nranks) if SLAB_CYCLIC else slice(rank * n, (rank + 1) * n)


class slab_context:
    """`with ex.spectral.slab_context(rank, nranks): stepper = ex.stepper.X(3, L, N, dt)` builds
    every spectral array of the stepper (operators, masks, ETDRK tables, injection) only on the
    local spectral slab (N, N/nranks, N/2+1) -- all of them are elementwise in the mode index, so
    restricting the wavenumber grid is enough.  The reference cannot construct 2048^3 operators on
    one device at all."""

    def __init__(self, rank: int, nranks: int):
        self.slab = (int(rank), int(nranks))

    def __enter__(self):
        global _SLAB
        self._prev, _SLAB = _SLAB, self.slab
        return self

    def __exit__(self, *exc):
        global _SLAB
        _SLAB = self._prev
        return False


def current_slab():
    return _SLAB


def _slab_slice(vec, num_spatial_dims, num_points):
    """restrict the axis-1 grid vector of a 3-D array to the local slab"""
    if _SLAB is None or num_spatial_dims != 3:
        return vec
    rank, nranks = _SLAB
    if num_points % nranks:
        raise ValueError(f"num_points={num_points} must be divisible by the number of slab ranks ({nranks})")
    return vec[slab_indices(rank, nranks, num_points)]


def build_wavenumbers(num_spatial_dims: int, num_points: int, *, indexing: str = "ij", dtype=None):
    """exponax/_spectral.py:13-49."""
    dtype = real_dtype() if dtype is None else dtype
    right = np.fft.rfftfreq(num_points, 1 / num_points).astype(dtype)
    other = np.fft.fftfreq(num_points, 1 / num_points).astype(dtype)
    grids = [other] * (num_spatial_dims - 1) + [right]
    if num_spatial_dims == 3:
        grids[1] = _slab_slice(other, 3, num_points)
    return np.stack(np.meshgrid(*grids, indexing=indexing))


def build_scaled_wavenumbers(num_spatial_dims, domain_extent, num_points, *, indexing="ij", dtype=None):
    """exponax/_spectral.py:52-83."""
    dtype = real_dtype() if dtype is None else dtype
    scale = dtype(2 * np.pi / domain_extent)
    return scale * build_wavenumbers(num_spatial_dims, num_points, indexing=indexing, dtype=dtype)


def build_derivative_operator(num_spatial_dims, domain_extent, num_points, *, indexing="ij", dtype=None):
    """exponax/_spectral.py:86-115: 1j * (2*pi/L) * k."""
    dtype = real_dtype() if dtype is None else dtype
    wn = build_scaled_wavenumbers(num_spatial_dims, domain_extent, num_points, indexing=indexing, dtype=dtype)
    return (1j * wn).astype(complex_dtype(dtype))


def build_laplace_operator(derivative_operator, *, order: int = 2):
    """exponax/_spectral.py:118-155."""
    if order % 2 != 0:
        raise ValueError("Order must be even.")
    if order == 0:
        return np.ones((1, *derivative_operator.shape[1:]), dtype=derivative_operator.dtype)
    return np.sum(derivative_operator**order, axis=0, keepdims=True)


def build_gradient_inner_product_operator(derivative_operator, velocity, *, order: int = 1):
    """exponax/_spectral.py:158-209."""
    if order % 2 != 1:
        raise ValueError("Order must be odd.")
    velocity = np.asarray(velocity, dtype=derivative_operator.real.dtype)
    if velocity.shape != (derivative_operator.shape[0],):
        raise ValueError(
            f"""Expected velocity shape to be {derivative_operator.shape[0]},
             got {velocity.shape}."""
        )
    operator = np.einsum("i,i...->...", velocity, derivative_operator**order)
    return operator[None, ...].astype(derivative_operator.dtype)


def space_indices(num_spatial_dims: int):
    """exponax/_spectral.py:212-227."""
    return tuple(range(-num_spatial_dims, 0))


def spatial_shape(num_spatial_dims: int, num_points: int):
    """exponax/_spectral.py:230-251."""
    return (num_points,) * num_spatial_dims


def wavenumber_shape(num_spatial_dims: int, num_points: int):
    """exponax/_spectral.py:254-275 (the local slab shape inside a `slab_context`)."""
    if _SLAB is not None and num_spatial_dims == 3:
        return (num_points, num_points // _SLAB[1], num_points // 2 + 1)
    return (num_points,) * (num_spatial_dims - 1) + (num_points // 2 + 1,)


def low_pass_filter_mask(num_spatial_dims, num_points, *, cutoff, axis_separate=True, indexing="ij", dtype=None):
    """exponax/_spectral.py:278-342 (cutoff inclusive, compared in the working precision)."""
    dtype = real_dtype() if dtype is None else dtype
    wn = build_wavenumbers(num_spatial_dims, num_points, indexing=indexing, dtype=dtype)
    cut = dtype(cutoff)
    if axis_separate:
        mask = True
        for g in wn:
            mask = mask & (np.abs(g) <= cut)
    else:
        mask = np.linalg.norm(wn, axis=0) <= cut
    return mask[np.newaxis, ...]


def dealias_kmax(num_points: int, cutoff: float, dtype=None) -> int:
    """Largest integer wavenumber kept by `|k| <= cutoff` in the working precision (what the
    kernels use instead of a mask array).  -1: nothing is kept."""
    dtype = real_dtype() if dtype is None else dtype
    k = np.arange(num_points // 2 + 1).astype(dtype)
    keep = np.nonzero(k <= dtype(cutoff))[0]
    return int(keep.max()) if keep.size else -1


def build_scaling_array(num_spatial_dims, num_points, *,
                        mode: Literal["norm_compensation", "reconstruction", "coef_extraction"],
                        indexing="ij", dtype=None):
    """exponax/_spectral.py:403-527."""
    dtype = real_dtype() if dtype is None else dtype
    den = {"norm_compensation": (1, 1), "reconstruction": (2, 1), "coef_extraction": (2, 2)}
    if mode not in den:
        raise ValueError("Invalid mode.")
    rden, oden = den[mode]
    right_wn = np.fft.rfftfreq(num_points, 1 / num_points)
    other_wn = np.fft.fftfreq(num_points, 1 / num_points)
    right = np.where(right_wn == 0, num_points, num_points / rden)
    other = np.where(other_wn == 0, num_points, num_points / oden)
    if num_points % 2 == 0:
        right = np.where(right_wn == num_points // 2, num_points, right)
        other = np.where(other_wn == -num_points // 2, num_points, other)
    grids = [other] * (num_spatial_dims - 1) + [right]
    if num_spatial_dims == 3:
        grids[1] = _slab_slice(other, 3, num_points)
    return np.prod(np.stack(np.meshgrid(*grids, indexing=indexing)), axis=0, keepdims=True).astype(dtype)


# ---------------------------------------------------------------------------- device transforms
_PLAIN_PLANS = {}


def _plain_plan(D: int, N: int, rd):
    """Order-0 plan without nonlinearity, used for the standalone transforms."""
    key = (D, N, np.dtype(rd).str, A.torch.cuda.current_device())
    p = _PLAIN_PLANS.get(key)
    if p is None:
        M = int(np.prod(wavenumber_shape(D, N)))
        p = nat.Plan(D=D, N=N, C_=1, E=1, order=0, dtype=rd, L=1.0, kmax=-1, nl={"kind": nat.NL_ZERO},
                     exp_term=np.ones(M, complex_dtype(rd)))
        _PLAIN_PLANS[key] = p
    return p


_WS = {}


def workspace(nbytes: int):
    """Grow-only scratch buffer handed to libexb as its workspace, one per (device, stream): calls
    enqueued on different streams may overlap on the GPU and must not share scratch memory."""
    if nbytes <= 0:
        return None
    key = (A.torch.cuda.current_device(), A.torch.cuda.current_stream().cuda_stream)
    w = _WS.get(key)
    if w is None or w.numel() < nbytes:
        _WS[key] = w = A.torch.empty(int(nbytes), dtype=A.torch.uint8, device="cuda")
    return w


def fft(field, *, num_spatial_dims: int | None = None):
    """Real-valued FFT of a field `(C, N, .., N)` (leading batch axes allowed),
    unnormalised == exponax.fft (exponax/_spectral.py:614-656)."""
    rd = real_dtype()
    t, kind = A.to_device(field, rd)
    if num_spatial_dims is None:
        num_spatial_dims = t.ndim - 1
    D = num_spatial_dims
    N = t.shape[-1]
    if any(s != N for s in t.shape[-D:]):
        raise ValueError("all spatial axes must have the same length")
    lead = t.shape[:-D]
    nf = int(np.prod(lead)) if lead else 1
    out = A.torch.empty(tuple(lead) + wavenumber_shape(D, N), dtype=A.cplx_t(rd), device="cuda")
    plan = _plain_plan(D, N, rd)
    ws = workspace(plan.workspace_bytes(nf))
    nat.check(nat.lib().exb_fft(plan.handle, A.stream_ptr(), nf, 1, A.ptr(t), A.ptr(out), A.ptr(ws)))
    return A.from_device(out, kind)


def ifft(field_hat, *, num_spatial_dims: int | None = None, num_points: int | None = None):
    """Inverse of `fft` == exponax.ifft (exponax/_spectral.py:659-721)."""
    rd = real_dtype()
    t, kind = A.to_device(field_hat, rd, complex_=True)
    if num_spatial_dims is None:
        num_spatial_dims = t.ndim - 1
    D = num_spatial_dims
    if num_points is None:
        if D >= 2:
            num_points = t.shape[-2]
        else:
            raise ValueError("num_points must be provided if num_spatial_dims == 1.")
    N = num_points
    if tuple(t.shape[-D:]) != wavenumber_shape(D, N):
        raise ValueError(f"expected trailing shape {wavenumber_shape(D, N)}, got {tuple(t.shape[-D:])}")
    lead = t.shape[:-D]
    nf = int(np.prod(lead)) if lead else 1
    out = A.torch.empty(tuple(lead) + spatial_shape(D, N), dtype=A.real_t(rd), device="cuda")
    plan = _plain_plan(D, N, rd)
    ws = workspace(plan.workspace_bytes(nf))
    nat.check(nat.lib().exb_ifft(plan.handle, A.stream_ptr(), nf, 1, A.ptr(t), A.ptr(out), A.ptr(ws)))
    return A.from_device(out, kind)


def derivative(field, domain_extent: float, *, order: int = 1, indexing: str = "ij"):
    """Spectral derivative of every channel along every axis == exponax.derivative
    (exponax/_spectral.py:724-792): `(C, N, .., N)` -> `(C, D, N, .., N)`, or `(D, N, .., N)` for C == 1.
    exb_fft -> exb_derivative (u_hat * (i k_d 2 pi / L)^order) -> exb_ifft."""
    if indexing != "ij":
        raise NotImplementedError("only indexing='ij' is supported")
    if int(order) != order or order < 0:
        raise ValueError("order must be a non-negative integer")
    rd = real_dtype()
    t, kind = A.to_device(field, rd)
    C_ = t.shape[0]
    D = t.ndim - 1
    N = t.shape[-1]
    if any(s != N for s in t.shape[1:]):
        raise ValueError("all spatial axes must have the same length")
    torch = A.torch
    uh = torch.empty((C_,) + wavenumber_shape(D, N), dtype=A.cplx_t(rd), device="cuda")
    dh = torch.empty((C_ * D,) + wavenumber_shape(D, N), dtype=A.cplx_t(rd), device="cuda")
    out = torch.empty((C_ * D,) + spatial_shape(D, N), dtype=A.real_t(rd), device="cuda")
    plan = _plain_plan(D, N, rd)
    ws = workspace(plan.workspace_bytes(C_ * D))
    lib, st = nat.lib(), A.stream_ptr()
    nat.check(lib.exb_fft(plan.handle, st, C_, 1, A.ptr(t), A.ptr(uh), A.ptr(ws)))
    nat.check(lib.exb_derivative(plan.handle, st, C_, A.ptr(uh), A.ptr(dh), int(order), float(domain_extent)))
    nat.check(lib.exb_ifft(plan.handle, st, C_ * D, 1, A.ptr(dh), A.ptr(out), A.ptr(ws)))
    out = out.view((C_, D) + spatial_shape(D, N))
    return A.from_device(out[0] if C_ == 1 else out, kind)


def get_spectrum(state, *, power: bool = True, radial_binning: Literal["average", "sum"] = "sum",
                 num_spatial_dims: int | None = None):
    """Power (default) or amplitude spectrum of a state `(C, N, .., N)` -> `(C, N//2 + 1)`, radially binned in
    2-D / 3-D == exponax.get_spectrum (exponax/_spectral.py:866-1030).  One `exb_fft` plus one binning kernel
    (`exb_spectrum`).  With `num_spatial_dims` given, leading batch axes are allowed: `(B, C, N, .., N)` ->
    `(B, C, N//2 + 1)` (what `jax.vmap(ex.get_spectrum)` returns in the reference)."""
    if radial_binning not in ("average", "sum"):
        raise ValueError("radial_binning must be 'average' or 'sum'")
    rd = real_dtype()
    t, kind = A.to_device(state, rd)
    D = t.ndim - 1 if num_spatial_dims is None else num_spatial_dims
    N = t.shape[-1]
    if any(s != N for s in t.shape[-D:]):
        raise ValueError("all spatial axes must have the same length")
    lead = tuple(t.shape[:-D])
    nf = int(np.prod(lead)) if lead else 1
    torch = A.torch
    uh = torch.empty(lead + wavenumber_shape(D, N), dtype=A.cplx_t(rd), device="cuda")
    out = torch.empty(lead + (N // 2 + 1,), dtype=A.real_t(rd), device="cuda")
    plan = _plain_plan(D, N, rd)
    ws = workspace(plan.workspace_bytes(nf))
    nat.check(nat.lib().exb_fft(plan.handle, A.stream_ptr(), nf, 1, A.ptr(t), A.ptr(uh), A.ptr(ws)))
    average = radial_binning == "average" and D > 1   # 1-D: one mode per bin, nothing to average
    counts = torch.empty(N // 2 + 1, dtype=torch.int32, device="cuda") if average else None
    M = int(np.prod(wavenumber_shape(D, N)))
    for f0 in range(0, nf, 65535):           # one launch covers at most 65535 fields (grid.y)
        n = min(65535, nf - f0)
        nat.check(nat.lib().exb_spectrum(plan.handle, A.stream_ptr(), n, A.ptr(uh) + f0 * M * uh.element_size(),
                                         A.ptr(out) + f0 * (N // 2 + 1) * out.element_size(), int(bool(power)),
                                         int(average), A.ptr(counts)))
    return A.from_device(out, kind)


def _get_spectrum_batched(states, **kw):
    return get_spectrum(states, num_spatial_dims=states.ndim - 2, **kw)


get_spectrum._batched = _get_spectrum_batched
