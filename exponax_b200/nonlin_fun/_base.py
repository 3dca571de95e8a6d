"""
Base class of the nonlinear functions (mirrors exponax/nonlin_fun/_base.py:9-159).

A built-in nonlinear function is a *descriptor*: constructor arguments are stored exactly as in
the reference, and `_native_desc()` translates them into the parameters the fused sm_100a
kernels need (kind, flags, scales, dealiasing cutoff).  Calling the object evaluates N(u_hat) on
the device through `exb_nonlinear_fun`.  User subclasses that only implement `__call__` with
`self.fft` / `self.ifft` (the reference's extension API) keep working through the standalone
device transforms -- unfused, but still on the GPU.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .. import _array as A
from .. import _native as nat
from .. import _spectral as sp
from .._config import complex_dtype, real_dtype


class BaseNonlinearFun(ABC):
    num_spatial_dims: int
    num_points: int

    def __init__(self, num_spatial_dims: int, num_points: int, *, dealiasing_fraction: float | None = None):
        self.num_spatial_dims = num_spatial_dims
        self.num_points = num_points
        self._dtype = real_dtype()
        self._plan_cache = {}
        if dealiasing_fraction is None:
            self.dealiasing_mask = None
            self._kmax = -1
        else:
            # exponax/nonlin_fun/_base.py:59-71
            nyquist_mode = (num_points // 2) + 1
            highest_resolved_mode = nyquist_mode - 1
            start_of_aliased_modes = dealiasing_fraction * highest_resolved_mode
            cutoff = start_of_aliased_modes - 1
            self.dealiasing_mask = sp.low_pass_filter_mask(num_spatial_dims, num_points, cutoff=cutoff,
                                                           dtype=self._dtype)
            self._kmax = sp.dealias_kmax(num_points, cutoff, self._dtype)
            if self._kmax < 0:
                raise NotImplementedError("dealiasing_fraction removes every mode")

    # ---- reference API -------------------------------------------------------------------
    def dealias(self, u_hat):
        if self.dealiasing_mask is None:
            raise ValueError("Nonlinear function was set up without dealiasing")
        t, kind = A.to_device(u_hat, self._dtype, complex_=True)
        mask = A.torch.as_tensor(self.dealiasing_mask, device="cuda")
        return A.from_device(t * mask, kind)

    def fft(self, u):
        u_hat = sp.fft(u, num_spatial_dims=self.num_spatial_dims)
        if self.dealiasing_mask is not None:
            u_hat = self.dealias(u_hat)
        return u_hat

    def ifft(self, u_hat):
        if self.dealiasing_mask is not None:
            u_hat = self.dealias(u_hat)
        return sp.ifft(u_hat, num_spatial_dims=self.num_spatial_dims, num_points=self.num_points)

    # ---- native dispatch -----------------------------------------------------------------
    def _native_desc(self, num_channels: int) -> dict | None:
        """Parameters for the fused kernels, or None for a user-defined function."""
        return None

    def _domain_extent(self) -> float:
        dop = getattr(self, "derivative_operator", None)
        if dop is None:
            return 1.0
        # derivative_operator[D-1][0, .., 0, 1] = 1j * 2*pi/L   (k = 1 on the last axis)
        idx = (dop.shape[0] - 1,) + (0,) * (dop.ndim - 2) + (1,)
        return float(2 * np.pi / dop[idx].imag)

    def _eval_plan(self, num_channels: int):
        key = (num_channels, A.torch.cuda.current_device())
        p = self._plan_cache.get(key)
        if p is None:
            desc = self._native_desc(num_channels)
            if desc is None:
                raise NotImplementedError("no fused kernel for this nonlinear function")
            D, N = self.num_spatial_dims, self.num_points
            M = int(np.prod(sp.wavenumber_shape(D, N)))
            p = nat.Plan(D=D, N=N, C_=num_channels, E=1, order=0, dtype=self._dtype, L=self._domain_extent(),
                         kmax=self._kmax, nl=desc, exp_term=np.ones(M, complex_dtype(self._dtype)))
            self._plan_cache[key] = p
        return p

    def _native_call(self, u_hat):
        D, N = self.num_spatial_dims, self.num_points
        t, kind = A.to_device(u_hat, self._dtype, complex_=True)
        wshape = sp.wavenumber_shape(D, N)
        if t.ndim < D + 1 or tuple(t.shape[-D:]) != wshape:
            raise ValueError(f"Expected trailing shape (C,)+{wshape}, got {tuple(t.shape)}")
        C = t.shape[-D - 1]
        lead = t.shape[: -D - 1]
        batch = int(np.prod(lead)) if lead else 1
        plan = self._eval_plan(C)
        if not plan.fused_ok():
            # 1-D grid too large for the fused shared-memory kernel: the reference's own formulation of this
            # function on device arrays around the native transforms (no performance claim, README "eager paths")
            return A.from_device(self._array_call(t), kind)
        out = A.torch.empty_like(t)
        ws = sp.workspace(plan.workspace_bytes(batch))
        nat.check(nat.lib().exb_nonlinear_fun(plan.handle, A.stream_ptr(), batch, A.ptr(t), A.ptr(out), A.ptr(ws)))
        return A.from_device(out, kind)

    def _array_call(self, u_hat):
        """Array-level evaluation (torch tensors, channel axis at -D-1) with `self.fft` / `self.ifft`."""
        raise NotImplementedError("no array-level formulation of this nonlinear function")

    def _dop(self):
        """derivative operator as a device tensor, shape (D, ..., N//2+1)."""
        return A.torch.as_tensor(np.asarray(self.derivative_operator), device="cuda")

    @abstractmethod
    def __call__(self, u_hat):
        """Evaluate the nonlinear function on a state in Fourier space."""
