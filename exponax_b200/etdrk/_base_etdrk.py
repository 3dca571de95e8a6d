"""
ETDRK integrators.  Constructor = host-side precompute of exp(dt L) and the real
phi-coefficients by the Kassam-Trefethen contour mean (as in the reference; uploaded to the
device once when the plan is created).  `step_fourier` = one fused device call when the
nonlinear function has a native kernel, otherwise the generic stage formulas evaluated with
device array arithmetic around the user's `nonlinear_fun` (the reference's extension API).

Mirrors exponax/etdrk/_base_etdrk.py:8-80.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .. import _array as A
from .. import _native as nat
from .. import _spectral as sp
from ._utils import roots_of_unity

GPU_TABLE_THRESHOLD = 1 << 26  # modes; above this the ETDRK tables are built on the GPU


class BaseETDRK(ABC):
    dt: float
    order: int = -1

    def __init__(self, dt: float, linear_operator):
        self.dt = dt
        self._nonlinear_fun = None
        self._plans = {}
        self._dev_coefs = {}
        if isinstance(linear_operator, A.torch.Tensor):
            # operator assembled directly on the GPU (lean slab constructors, config c5)
            self._cd = np.dtype(np.complex64 if linear_operator.dtype == A.torch.complex64 else np.complex128)
            self._rd = np.float32 if linear_operator.dtype == A.torch.complex64 else np.float64
            self._linear_operator = linear_operator
            self._on_gpu = True
            self._L_dev = linear_operator
            self._exp_term = A.torch.exp(self._rd(dt) * self._L_dev)
            return
        linear_operator = np.asarray(linear_operator)
        if not np.iscomplexobj(linear_operator):
            linear_operator = linear_operator.astype(np.complex128 if linear_operator.dtype == np.float64
                                                     else np.complex64)
        self._cd = linear_operator.dtype
        self._rd = linear_operator.real.dtype.type
        self._linear_operator = linear_operator
        # Grids whose tables take minutes in NumPy (config c5: > 1e8 modes per rank) build them with
        # the same formulas on the GPU; the tables then stay on the device (tables_on_device).
        self._on_gpu = linear_operator.size >= GPU_TABLE_THRESHOLD and A.torch.cuda.is_available()
        if self._on_gpu:
            self._L_dev = A.torch.as_tensor(linear_operator, device="cuda")
            self._exp_term = A.torch.exp(self._rd(dt) * self._L_dev)
        else:
            self._exp_term = np.exp(self._rd(dt) * linear_operator).astype(self._cd)
        self._nonlinear_fun = None
        self._plans = {}
        self._dev_coefs = {}

    # ---- host-side coefficient precompute ------------------------------------------------
    def _contour_means(self, fns, num_circle_points: int, circle_radius: float):
        """dt * mean_j Re[f(r_j + dt L)] for each f; accumulation order = the reference's scan
        over the roots (exponax/etdrk/_etdrk_2.py:74-89)."""
        rd, cd = self._rd, self._cd
        roots = roots_of_unity(num_circle_points, rd)
        if self._on_gpu:
            t = A.torch
            L_dt = self._L_dev * rd(self.dt)
            accs = [t.zeros_like(L_dt.real) for _ in fns]
            need_half = getattr(self, "_needs_half_exp", True)
            for root in roots:
                lr = L_dt + complex(rd(circle_radius) * root)
                exp_lr = t.exp(lr)
                exp_lr_half = t.exp(lr / 2) if need_half else None
                for i, f in enumerate(fns):
                    accs[i] += f(lr, exp_lr, exp_lr_half).real
                del lr, exp_lr, exp_lr_half
            return [(a / rd(num_circle_points)) * rd(self.dt) for a in accs]
        L_dt = (self._linear_operator * rd(self.dt)).astype(cd)
        accs = [np.zeros_like(L_dt.real) for _ in fns]
        for root in roots:
            lr = (rd(circle_radius) * root + L_dt).astype(cd)
            exp_lr = np.exp(lr)
            exp_lr_half = np.exp(lr / rd(2))
            for i, f in enumerate(fns):
                accs[i] = accs[i] + f(lr, exp_lr, exp_lr_half).real.astype(rd)
        return [(rd(self.dt) * (a / rd(num_circle_points))).astype(rd) for a in accs]

    def _coef_list(self):
        return []

    def _half_exp(self):
        return None

    # ---- native plan ---------------------------------------------------------------------
    def _plan(self, num_channels: int, num_points: int, domain_extent: float):
        key = (num_channels, A.torch.cuda.current_device())
        p = self._plans.get(key)
        if p is None:
            nl = self._nonlinear_fun
            D = self._linear_operator.ndim - 1
            if self.order == 0 or nl is None:
                desc, kmax = {"kind": nat.NL_ZERO}, -1
            else:
                desc = nl._native_desc(num_channels)
                kmax = nl._kmax
                if desc is None:
                    return None
            p = nat.Plan(D=D, N=num_points, C_=num_channels, E=self._linear_operator.shape[0],
                         order=self.order, dtype=self._rd, L=domain_extent, kmax=kmax, nl=desc,
                         exp_term=self._exp_term, half_exp_term=self._half_exp(), coefs=self._coef_list())
            self._plans[key] = p
        return p

    def _dev(self, name):
        """device copy of a coefficient array (generic path only)."""
        key = (name, A.torch.cuda.current_device())
        t = self._dev_coefs.get(key)
        if t is None:
            t = A.torch.as_tensor(getattr(self, name), device="cuda")
            self._dev_coefs[key] = t
        return t

    @abstractmethod
    def step_fourier(self, u_hat):
        """Advance the state in Fourier space (generic, array-level formulation)."""
