"""`exponax.metrics`, spatial family + correlation, on the device.

Every function makes ONE pass over its inputs (`exb_metric_sums`: per channel sum|a-b|^p, sum|b|^p,
sum|a|^p, sum a*b in a single fused reduction kernel) and combines the handful of sums on the device; the
result is a 0-d array of the caller's framework.  Signatures, argument meaning and error messages follow
exponax/metrics/_spatial.py, _correlation.py and _utils.py.  The Fourier / Sobolev (H1) families are not
mirrored yet (they stay with the reference)."""
from __future__ import annotations

from typing import Literal

import numpy as np

from .. import _array as A
from .. import _native as nat
from .._config import real_dtype

__all__ = ["spatial_aggregator", "spatial_norm", "MAE", "nMAE", "sMAE", "MSE", "nMSE", "sMSE", "RMSE", "nRMSE",
           "sRMSE", "correlation", "mean_metric"]


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
    nat.check(nat.lib().exb_metric_sums(A.stream_ptr(), nat.EXB_F32 if rd == np.float32 else nat.EXB_F64, nfields,
                                        npoints, A.ptr(ta), A.ptr(tb), float(p), A.ptr(out)))
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
