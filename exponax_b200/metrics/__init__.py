"""`exponax.metrics`, spatial family + correlation, on the device.

Every function makes ONE pass over its inputs (`exb_metric_sums`: per channel sum|a-b|^p, sum|b|^p,
sum|a|^p, sum a*b in a single fused reduction kernel) and combines the handful of sums on the device; the
result is a 0-d array of the caller's framework.  Signatures, argument meaning and error messages follow
exponax/metrics/_spatial.py, _correlation.py and _utils.py.  The Fourier / Sobolev (H1) families
(metrics/_fourier.py, _derivative.py) are `exb_fft` + one pass of `exb_fourier_sums` per operand."""
from __future__ import annotations

from typing import Literal

import numpy as np

from .. import _array as A
from .. import _native as nat
from .._config import real_dtype

__all__ = ["spatial_aggregator", "spatial_norm", "MAE", "nMAE", "sMAE", "MSE", "nMSE", "sMSE", "RMSE", "nRMSE",
           "sRMSE", "correlation", "mean_metric", "fourier_aggregator", "fourier_norm", "fourier_MAE",
           "fourier_nMAE", "fourier_MSE", "fourier_nMSE", "fourier_RMSE", "fourier_nRMSE", "H1_MAE", "H1_nMAE",
           "H1_MSE", "H1_nMSE", "H1_RMSE", "H1_nRMSE"]


def _sums(a, b, p: float, nfields: int):
    """(nfields, 4) float64 device tensor of the four sums of every field."""
    rd = real_dtype()
    ta, kind = A.to_device(a, rd)
    tb = None
    if b is not None:
        tb, _ = A.to_device(b, rd)
        if tuple(tb.shape) != tuple(ta.shape):
            raise ValueError(f"shape mismatch: {tuple(ta.shape)} vs {tuple(tb.shape)}")
    npoints = ta.numel() // nfields
    out = A.torch.empty((nfields, 4), dtype=A.torch.float64, device="cuda")
    es = ta.element_size()
    for f0 in range(0, nfields, 65535):      # one launch covers at most 65535 fields (grid.y)
        n = min(65535, nfields - f0)
        nat.check(nat.lib().exb_metric_sums(
            A.stream_ptr(), nat.EXB_F32 if rd == np.float32 else nat.EXB_F64, n, npoints,
            A.ptr(ta) + f0 * npoints * es, (A.ptr(tb) + f0 * npoints * es) if tb is not None else 0, float(p),
            A.ptr(out) + f0 * 4 * 8))
    return out, kind, rd, ta


def _finish(x, kind, rd):
    return A.from_device(x.to(A.real_t(rd)), kind)


def spatial_aggregator(state_no_channel, *, num_spatial_dims: int | None = None, domain_extent: float = 1.0,
                       num_points: int | None = None, inner_exponent: float = 2.0,
                       outer_exponent: float | None = None):
    """((L/N)^D sum_i |u_i|^p)^q of a channel-less state (exponax/metrics/_spatial.py:8-83)."""
    s, kind, rd, t = _sums(state_no_channel, None, inner_exponent, 1)
    if num_spatial_dims is None:
        num_spatial_dims = t.ndim
    if num_points is None:
        num_points = t.shape[-1]
    if outer_exponent is None:
        outer_exponent = 1 / inner_exponent
    scale = (domain_extent / num_points) ** num_spatial_dims
    return _finish((scale * s[0, 2]) ** outer_exponent, kind, rd)


def _norm_from_sums(s, t, lead, has_ref, mode, domain_extent, inner_exponent, outer_exponent):
    """s: (prod(lead) * C, 4) sums; t: the device tensor (lead..., C, N, .., N) -> metric of shape lead."""
    D = t.ndim - 1 - len(lead)
    N = t.shape[-1]
    q = 1 / inner_exponent if outer_exponent is None else outer_exponent
    scale = (domain_extent / N) ** D
    diff = (scale * (s[:, 0] if has_ref else s[:, 2])) ** q
    if mode == "normalized":
        per_channel = diff / (scale * s[:, 1]) ** q
    elif mode == "symmetric":
        per_channel = 2 * diff / ((scale * s[:, 2]) ** q + (scale * s[:, 1]) ** q)
    else:
        per_channel = diff
    return per_channel.reshape(tuple(lead) + (-1,)).sum(dim=-1)


def _spatial_norm(state, state_ref, lead_ndim, *, mode, domain_extent, inner_exponent, outer_exponent):
    if state_ref is None:
        if mode == "normalized":
            raise ValueError("mode 'normalized' requires state_ref")
        if mode == "symmetric":
            raise ValueError("mode 'symmetric' requires state_ref")
    shape = np.shape(state)
    lead = shape[:lead_ndim]
    nfields = int(np.prod(shape[:lead_ndim + 1]))
    s, kind, rd, t = _sums(state, state_ref, inner_exponent, nfields)
    return _finish(_norm_from_sums(s, t, lead, state_ref is not None, mode, domain_extent, inner_exponent,
                                   outer_exponent), kind, rd)


def spatial_norm(state, state_ref=None, *, mode: Literal["absolute", "normalized", "symmetric"] = "absolute",
                 domain_extent: float = 1.0, inner_exponent: float = 2.0, outer_exponent: float | None = None):
    """Consistent counterpart of the L^p norm of `state` (or of `state - state_ref`), absolute / normalised /
    symmetric, summed over channels AFTER the aggregation (exponax/metrics/_spatial.py:86-196)."""
    return _spatial_norm(state, state_ref, 0, mode=mode, domain_extent=domain_extent,
                         inner_exponent=inner_exponent, outer_exponent=outer_exponent)


def _make(mode, p, q, need_ref, name, ref_lines):
    if need_ref:
        def fn(u_pred, u_ref, *, domain_extent: float = 1.0):
            return spatial_norm(u_pred, u_ref, mode=mode, domain_extent=domain_extent, inner_exponent=p,
                                outer_exponent=q)
    else:
        def fn(u_pred, u_ref=None, *, domain_extent: float = 1.0):
            return spatial_norm(u_pred, u_ref, mode=mode, domain_extent=domain_extent, inner_exponent=p,
                                outer_exponent=q)

    def batched(u_pred, u_ref=None, *, domain_extent: float = 1.0):
        # leading batch axis, one fused reduction for the whole batch (what jax.vmap(metric) does)
        return _spatial_norm(u_pred, u_ref, 1, mode=mode, domain_extent=domain_extent, inner_exponent=p,
                             outer_exponent=q)
    fn._batched = batched
    fn.__name__ = fn.__qualname__ = name
    fn.__doc__ = f"exponax.metrics.{name} (exponax/metrics/_spatial.py:{ref_lines}): spatial_norm(mode={mode!r}, " \
                 f"inner_exponent={p}, outer_exponent={q})."
    return fn


MAE = _make("absolute", 1.0, 1.0, False, "MAE", "198-242")
nMAE = _make("normalized", 1.0, 1.0, True, "nMAE", "245-287")
sMAE = _make("symmetric", 1.0, 1.0, True, "sMAE", "290-340")
MSE = _make("absolute", 2.0, 1.0, False, "MSE", "343-387")
nMSE = _make("normalized", 2.0, 1.0, True, "nMSE", "390-432")
sMSE = _make("symmetric", 2.0, 1.0, True, "sMSE", "435-485")
RMSE = _make("absolute", 2.0, 0.5, False, "RMSE", "488-534")
nRMSE = _make("normalized", 2.0, 0.5, True, "nRMSE", "537-584")
sRMSE = _make("symmetric", 2.0, 0.5, True, "sRMSE", "587-637")


def _correlation(u_pred, u_ref, lead_ndim):
    shape = np.shape(u_pred)
    s, kind, rd, _ = _sums(u_pred, u_ref, 2.0, int(np.prod(shape[:lead_ndim + 1])))
    per_channel = s[:, 3] / (s[:, 2].sqrt() * s[:, 1].sqrt())
    return _finish(per_channel.reshape(tuple(shape[:lead_ndim]) + (-1,)).mean(dim=-1), kind, rd)


def correlation(u_pred, u_ref):
    """Mean over channels of <u, v> / (|u| |v|) (exponax/metrics/_correlation.py:6-60)."""
    return _correlation(u_pred, u_ref, 0)


correlation._batched = lambda u_pred, u_ref: _correlation(u_pred, u_ref, 1)


def mean_metric(metric_fn, *args, **kwargs):
    """'Meanifies' a metric over a leading batch axis (exponax/metrics/_utils.py:5-18)."""
    if hasattr(metric_fn, "_batched"):
        vals = metric_fn._batched(*args, **kwargs)
        return vals.mean(axis=0) if isinstance(vals, np.ndarray) else vals.mean(dim=0)
    n = len(args[0])
    vals = [metric_fn(*(a[i] for a in args), **kwargs) for i in range(n)]
    if isinstance(vals[0], np.ndarray) or np.isscalar(vals[0]):
        return np.mean(np.stack([np.asarray(v) for v in vals]), axis=0)
    return A.torch.stack(vals).mean(dim=0)


# ---------------------------------------------------------------------------- Fourier / Sobolev families
def _fourier_agg(t, lead_ndim, D, *, domain_extent, inner_exponent, outer_exponent, low, high, derivative_order):
    """fourier_aggregator of every field of the device tensor `t` (lead..., N, .., N) -> (prod(lead),) float64."""
    from .. import _spectral as sp
    rd = real_dtype()
    N = t.shape[-1]
    nfields = int(np.prod(t.shape[:lead_ndim])) if lead_ndim else 1
    q = 1 / inner_exponent if outer_exponent is None else outer_exponent
    th = sp.fft(t.reshape((nfields,) + tuple(t.shape[lead_ndim:])), num_spatial_dims=D)
    ncomp = D if derivative_order is not None else 1
    out = A.torch.empty((nfields, ncomp), dtype=A.torch.float64, device="cuda")
    plan = sp._plain_plan(D, N, rd)
    for f0 in range(0, nfields, 65535):
        n = min(65535, nfields - f0)
        nat.check(nat.lib().exb_fourier_sums(
            plan.handle, A.stream_ptr(), n, th[f0:f0 + n].data_ptr(), float(inner_exponent),
            -1 if low is None else int(low), -1 if high is None else int(high),
            -1.0 if derivative_order is None else float(derivative_order), float(domain_extent),
            out[f0:f0 + n].data_ptr()))
    scale = (domain_extent / N) ** D
    return ((scale * out) ** q).sum(dim=1)


def fourier_aggregator(state_no_channel, *, num_spatial_dims: int | None = None, domain_extent: float = 1.0,
                       num_points: int | None = None, inner_exponent: float = 2.0,
                       outer_exponent: float | None = None, low: int | None = None, high: int | None = None,
                       derivative_order: float | None = None):
    """Fourier-space counterpart of `spatial_aggregator` with optional band-pass [low, high] and derivative
    weighting (exponax/metrics/_fourier.py:15-140)."""
    rd = real_dtype()
    t, kind = A.to_device(state_no_channel, rd)
    D = t.ndim if num_spatial_dims is None else num_spatial_dims
    v = _fourier_agg(t, 0, D, domain_extent=domain_extent, inner_exponent=inner_exponent,
                     outer_exponent=outer_exponent, low=low, high=high, derivative_order=derivative_order)
    return _finish(v[0], kind, rd)


def _fourier_norm(state, state_ref, lead_ndim, *, mode, domain_extent, inner_exponent, outer_exponent, low, high,
                  derivative_order):
    if state_ref is None and mode == "normalized":
        raise ValueError("mode 'normalized' requires state_ref")
    rd = real_dtype()
    ta, kind = A.to_device(state, rd)
    tb = None
    if state_ref is not None:
        tb, _ = A.to_device(state_ref, rd)
        if tuple(tb.shape) != tuple(ta.shape):
            raise ValueError(f"shape mismatch: {tuple(ta.shape)} vs {tuple(tb.shape)}")
    D = ta.ndim - 1 - lead_ndim
    kw = dict(domain_extent=domain_extent, inner_exponent=inner_exponent, outer_exponent=outer_exponent, low=low,
              high=high, derivative_order=derivative_order)
    diff = ta if tb is None else ta - tb                     # the reference transforms the physical difference
    per_channel = _fourier_agg(diff, lead_ndim + 1, D, **kw)
    if mode == "normalized":
        per_channel = per_channel / _fourier_agg(tb, lead_ndim + 1, D, **kw)
    lead = tuple(ta.shape[:lead_ndim])
    return _finish(per_channel.reshape(lead + (-1,)).sum(dim=-1), kind, rd)


def fourier_norm(state, state_ref=None, *, mode: Literal["absolute", "normalized"] = "absolute",
                 domain_extent: float = 1.0, inner_exponent: float = 2.0, outer_exponent: float | None = None,
                 low: int | None = None, high: int | None = None, derivative_order: float | None = None):
    """exponax/metrics/_fourier.py:143-235."""
    return _fourier_norm(state, state_ref, 0, mode=mode, domain_extent=domain_extent, inner_exponent=inner_exponent,
                         outer_exponent=outer_exponent, low=low, high=high, derivative_order=derivative_order)


def _make_fourier(mode, p, q, need_ref, name, ref_lines):
    def impl(lead, u_pred, u_ref, domain_extent, low, high, derivative_order):
        return _fourier_norm(u_pred, u_ref, lead, mode=mode, domain_extent=domain_extent, inner_exponent=p,
                             outer_exponent=q, low=low, high=high, derivative_order=derivative_order)
    if need_ref:
        def fn(u_pred, u_ref, *, domain_extent: float = 1.0, low: int | None = None, high: int | None = None,
               derivative_order: float | None = None):
            return impl(0, u_pred, u_ref, domain_extent, low, high, derivative_order)
    else:
        def fn(u_pred, u_ref=None, *, domain_extent: float = 1.0, low: int | None = None, high: int | None = None,
               derivative_order: float | None = None):
            return impl(0, u_pred, u_ref, domain_extent, low, high, derivative_order)

    def batched(u_pred, u_ref=None, *, domain_extent: float = 1.0, low=None, high=None, derivative_order=None):
        return impl(1, u_pred, u_ref, domain_extent, low, high, derivative_order)
    fn._batched = batched
    fn.__name__ = fn.__qualname__ = name
    fn.__doc__ = f"exponax.metrics.{name} (exponax/metrics/_fourier.py:{ref_lines}): fourier_norm(mode={mode!r}, " \
                 f"inner_exponent={p}, outer_exponent={q})."
    return fn


fourier_MAE = _make_fourier("absolute", 1.0, 1.0, False, "fourier_MAE", "238-294")
fourier_nMAE = _make_fourier("normalized", 1.0, 1.0, True, "fourier_nMAE", "297-353")
fourier_MSE = _make_fourier("absolute", 2.0, 1.0, False, "fourier_MSE", "356-413")
fourier_nMSE = _make_fourier("normalized", 2.0, 1.0, True, "fourier_nMSE", "416-471")
fourier_RMSE = _make_fourier("absolute", 2.0, 0.5, False, "fourier_RMSE", "474-530")
fourier_nRMSE = _make_fourier("normalized", 2.0, 0.5, True, "fourier_nRMSE", "533-588")


def _make_h1(base, need_ref, name, ref_lines):
    if need_ref:
        def fn(u_pred, u_ref, *, domain_extent: float = 1.0, low: int | None = None, high: int | None = None):
            kw = dict(domain_extent=domain_extent, low=low, high=high)
            return base(u_pred, u_ref, derivative_order=None, **kw) + base(u_pred, u_ref, derivative_order=1, **kw)
    else:
        def fn(u_pred, u_ref=None, *, domain_extent: float = 1.0, low: int | None = None, high: int | None = None):
            kw = dict(domain_extent=domain_extent, low=low, high=high)
            return base(u_pred, u_ref, derivative_order=None, **kw) + base(u_pred, u_ref, derivative_order=1, **kw)

    def batched(u_pred, u_ref=None, *, domain_extent: float = 1.0, low=None, high=None):
        kw = dict(domain_extent=domain_extent, low=low, high=high)
        return base._batched(u_pred, u_ref, derivative_order=None, **kw) + \
            base._batched(u_pred, u_ref, derivative_order=1, **kw)
    fn._batched = batched
    fn.__name__ = fn.__qualname__ = name
    fn.__doc__ = f"exponax.metrics.{name} (exponax/metrics/_derivative.py:{ref_lines}): fourier metric + the same " \
                 "metric of the first derivatives."
    return fn


H1_MAE = _make_h1(fourier_MAE, False, "H1_MAE", "13-70")
H1_nMAE = _make_h1(fourier_nMAE, True, "H1_nMAE", "73-128")
H1_MSE = _make_h1(fourier_MSE, False, "H1_MSE", "131-188")
H1_nMSE = _make_h1(fourier_nMSE, True, "H1_nMSE", "191-247")
H1_RMSE = _make_h1(fourier_RMSE, False, "H1_RMSE", "250-307")
H1_nRMSE = _make_h1(fourier_nRMSE, True, "H1_nRMSE", "310-366")
