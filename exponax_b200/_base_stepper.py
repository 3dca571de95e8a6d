"""
BaseStepper -- same constructor protocol, attributes, methods and error messages as
exponax/_base_stepper.py:16-271; `step`, `step_fourier` and the rollout loop dispatch to the
fused sm_100a kernels of libexb through its C ABI (include/exb.h).

The constructor is host code (NumPy), like the reference's: derivative operator, linear
operator (`_build_linear_operator`), nonlinear function (`_build_nonlinear_fun`) and the ETDRK
coefficient tables are built once; the device plan is created lazily on the first call so that
constructing a stepper needs no GPU.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from . import _array as A
from . import _native as nat
from . import _spectral as sp
from ._config import real_dtype
from .etdrk import ETDRK0, ETDRK1, ETDRK2, ETDRK3, ETDRK4, BaseETDRK
from .nonlin_fun import BaseNonlinearFun


class BaseStepper(ABC):
    num_spatial_dims: int
    domain_extent: float
    num_points: int
    num_channels: int
    dt: float
    dx: float
    _integrator: BaseETDRK

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 num_channels: int, order: int, num_circle_points: int = 16, circle_radius: float = 1.0):
        self.num_spatial_dims = num_spatial_dims
        self.domain_extent = domain_extent
        self.num_points = num_points
        self.dt = dt
        self.num_channels = num_channels
        self.dx = domain_extent / num_points
        self._dtype = real_dtype()
        self._slab = sp.current_slab()  # (rank, nranks) when built inside ex.spectral.slab_context

        derivative_operator = sp.build_derivative_operator(num_spatial_dims, domain_extent, num_points,
                                                           dtype=self._dtype)
        linear_operator = np.asarray(self._build_linear_operator(derivative_operator))
        single_channel_shape = (1,) + sp.wavenumber_shape(self.num_spatial_dims, self.num_points)
        multi_channel_shape = (self.num_channels,) + sp.wavenumber_shape(self.num_spatial_dims, self.num_points)
        if linear_operator.shape not in (single_channel_shape, multi_channel_shape):
            raise ValueError(
                f"""Expected linear operator to have shape
                 {single_channel_shape} or {multi_channel_shape}, got
                 {linear_operator.shape}."""
            )
        linear_operator = linear_operator.astype(sp.complex_dtype(self._dtype))
        nonlinear_fun = self._build_nonlinear_fun(derivative_operator)

        kw = dict(num_circle_points=num_circle_points, circle_radius=circle_radius)
        if order == 0:
            self._integrator = ETDRK0(dt, linear_operator)
        elif order == 1:
            self._integrator = ETDRK1(dt, linear_operator, nonlinear_fun, **kw)
        elif order == 2:
            self._integrator = ETDRK2(dt, linear_operator, nonlinear_fun, **kw)
        elif order == 3:
            self._integrator = ETDRK3(dt, linear_operator, nonlinear_fun, **kw)
        elif order == 4:
            self._integrator = ETDRK4(dt, linear_operator, nonlinear_fun, **kw)
        else:
            raise NotImplementedError(f"Order {order} not implemented.")
        self._order = order
        self._nonlinear_fun = nonlinear_fun
        self._native = None  # decided lazily

    @abstractmethod
    def _build_linear_operator(self, derivative_operator):
        """Assemble the L operator in Fourier space, shape (1|C, ..., N//2+1)."""

    @abstractmethod
    def _build_nonlinear_fun(self, derivative_operator) -> BaseNonlinearFun:
        """Build the function that evaluates the nonlinear part of the PDE in Fourier space."""

    # ---- native dispatch -----------------------------------------------------------------
    def _plan(self):
        """The libexb plan of this stepper, or None when the nonlinear function is user code."""
        if self._native is False:
            return None
        if self._slab is not None:
            raise RuntimeError("this stepper was built inside a slab_context: use it through ex.SlabStepper")
        p = self._integrator._plan(self.num_channels, self.num_points, self.domain_extent)
        if p is not None and not p.fused_ok():
            # 1-D grid too large for the persistent shared-memory kernel: the stage formulas run on device arrays
            # around the native transforms (`_step_fourier_generic`), as for user-defined nonlinear functions
            p = None
        self._native = p is not None
        return p

    def _plan_available(self) -> bool:
        """True when this stepper runs on the fused kernels (built-in nonlinear function)."""
        if self._native is None:
            nl = self._nonlinear_fun
            self._native = self._order == 0 or nl._native_desc(self.num_channels) is not None
        return bool(self._native)

    def _step_fourier_generic(self, u_hat):
        """Array-level step in Fourier space used when no fused plan applies (user-defined nonlinear
        functions, steppers overriding `step_fourier` such as `Wave`)."""
        return self._integrator.step_fourier(u_hat)

    def _state_shape(self):
        return (self.num_channels,) + sp.spatial_shape(self.num_spatial_dims, self.num_points)

    def _fourier_shape(self):
        return (self.num_channels,) + sp.wavenumber_shape(self.num_spatial_dims, self.num_points)

    def _split_batch(self, t, shape):
        nd = len(shape)
        if t.ndim < nd or tuple(t.shape[-nd:]) != tuple(shape):
            raise ValueError(
                f"""Expected shape {shape}, got {tuple(t.shape)}. For batched
                 operation use `jax.vmap` on this function."""
            )
        lead = tuple(t.shape[:-nd])
        return lead, int(np.prod(lead)) if lead else 1

    def _step_batched(self, u, *, substeps: int = 1):
        """u: (..., C, N, .., N) on any supported array type; leading axes are batch axes."""
        t, kind = A.to_device(u, self._dtype)
        lead, batch = self._split_batch(t, self._state_shape())
        plan = self._plan()
        if plan is None:
            u_hat = sp.fft(t, num_spatial_dims=self.num_spatial_dims)
            for _ in range(substeps):
                u_hat = self._step_fourier_generic(u_hat)
            out = sp.ifft(u_hat, num_spatial_dims=self.num_spatial_dims, num_points=self.num_points)
            return A.from_device(out, kind)
        out = A.torch.empty_like(t)
        ws = sp.workspace(plan.workspace_bytes(batch))
        nat.check(nat.lib().exb_rollout(plan.handle, A.stream_ptr(), batch, 1, substeps, nat.ROLLOUT_FINAL_ONLY,
                                        A.ptr(t), A.ptr(out), A.ptr(ws)))
        return A.from_device(out, kind)

    def _step_fourier_batched(self, u_hat, *, substeps: int = 1):
        t, kind = A.to_device(u_hat, self._dtype, complex_=True)
        lead, batch = self._split_batch(t, self._fourier_shape())
        plan = self._plan()
        if plan is None:
            for _ in range(substeps):
                t = self._step_fourier_generic(t)
            return A.from_device(t, kind)
        out = A.torch.empty_like(t)
        ws = sp.workspace(plan.workspace_bytes(batch))
        src = t
        for _ in range(substeps):
            nat.check(nat.lib().exb_step_fourier(plan.handle, A.stream_ptr(), batch, A.ptr(src), A.ptr(out), A.ptr(ws)))
            src = out
        return A.from_device(out, kind)

    def _forcing_hat(self, f, lead, n, constant, time_first):
        """Fourier transform of a forcing + its (step, trajectory) strides in complex elements.
        constant: f is (C, N..) [shared] or lead + (C, N..); otherwise (n, C, N..) [shared], lead + (n, C, N..)
        (`vmap(rollout(..))`) or (n,) + lead + (C, N..) (`rollout(vmap(..))`, time_first)."""
        tf, _ = A.to_device(f, self._dtype)
        st = self._state_shape()
        per = int(np.prod(self._fourier_shape()))
        fl = tuple(tf.shape[:-len(st)])
        if tuple(tf.shape[-len(st):]) != tuple(st):
            raise ValueError(f"Expected a forcing with trailing shape {st}, got {tuple(tf.shape)}")
        batch = int(np.prod(lead)) if lead else 1
        if constant:
            if fl == ():
                strides = (0, 0)
            elif fl == tuple(lead):
                strides = (0, per)
            else:
                raise ValueError(f"constant forcing of shape {tuple(tf.shape)} does not match the batch shape {lead}")
        else:
            if fl == (n,):
                strides = (per, 0)
            elif fl == tuple(lead) + (n,):
                strides = (per, n * per)
            elif time_first and fl == (n,) + tuple(lead):
                strides = (batch * per, per)
            else:
                raise ValueError(f"forcing of shape {tuple(tf.shape)}: expected a leading time axis of length {n}")
        fh = sp.fft(tf, num_spatial_dims=self.num_spatial_dims)
        return fh.contiguous(), strides

    def _rollout_batched(self, u0, n: int, *, include_init: bool, layout_tb: bool, final_only: bool,
                         substeps: int = 1, spectral_carry: bool = False, cuda_graph: bool = False,
                         forcing=None, forcing_constant: bool = True):
        """Fused `rollout`/`repeat` over a batch: u0 (B.., C, N..) -> (B.., T, C, N..) /
        (T, B.., C, N..) / (B.., C, N..).  `forcing`: ForcedStepper semantics, `step(u + dt f)` every step."""
        t, kind = A.to_device(u0, self._dtype)
        lead, batch = self._split_batch(t, self._state_shape())
        plan = self._plan()
        if plan is None:
            raise NotImplementedError("fused rollout needs a native nonlinear function")
        if forcing is not None:
            fh, (fstep, fbatch) = self._forcing_hat(forcing, lead, n, forcing_constant, layout_tb)
            T = 1 if final_only else n + (1 if include_init else 0)
            st = self._state_shape()
            shape = lead + st if final_only else ((T,) + lead + st if layout_tb else lead + (T,) + st)
            out = A.torch.empty(shape, dtype=t.dtype, device="cuda")
            flags = ((nat.ROLLOUT_INCLUDE_INIT if include_init else 0) | (nat.ROLLOUT_LAYOUT_TB if layout_tb else 0)
                     | (nat.ROLLOUT_FINAL_ONLY if final_only else 0))
            ws = sp.workspace(plan.workspace_bytes(batch))
            nat.check(nat.lib().exb_rollout_forced(plan.handle, A.stream_ptr(), batch, n, 1, flags, A.ptr(t), A.ptr(out),
                                                   A.ptr(ws), A.ptr(fh), fstep, fbatch, float(self.dt)))
            return A.from_device(out, kind)
        T = 1 if final_only else n + (1 if include_init else 0)
        st = self._state_shape()
        if final_only:
            shape = lead + st
        elif layout_tb:
            shape = (T,) + lead + st
        else:
            shape = lead + (T,) + st
        out = A.torch.empty(shape, dtype=t.dtype, device="cuda")
        flags = 0
        flags |= nat.ROLLOUT_INCLUDE_INIT if include_init else 0
        flags |= nat.ROLLOUT_LAYOUT_TB if layout_tb else 0
        flags |= nat.ROLLOUT_FINAL_ONLY if final_only else 0
        flags |= nat.ROLLOUT_SPECTRAL_CARRY if spectral_carry else 0
        if cuda_graph:
            return A.from_device(self._rollout_graph(plan, t, shape, batch, n, substeps, flags), kind)
        ws = sp.workspace(plan.workspace_bytes(batch))
        nat.check(nat.lib().exb_rollout(plan.handle, A.stream_ptr(), batch, n, substeps, flags, A.ptr(t),
                                        A.ptr(out), A.ptr(ws)))
        return A.from_device(out, kind)

    _GRAPH_CACHE_ENTRIES = 4

    def release_graphs(self):
        """Drop every captured CUDA graph of this stepper together with the buffers it owns."""
        self.__dict__.pop("_graphs", None)

    def _rollout_graph(self, plan, t, shape, batch, n, substeps, flags):
        """Replay the launch sequence of one fused rollout from a captured CUDA graph (the library only
        enqueues on the stream it is given, so the whole call is capturable).  Pays off for N-D problems
        whose hundreds of small pass launches are launch-bound; buffers are owned by the graph entry."""
        key = (tuple(t.shape), t.dtype, n, substeps, flags, A.torch.cuda.current_device())
        cache = self.__dict__.setdefault("_graphs", {})
        ent = cache.pop(key, None)
        if ent is not None:
            cache[key] = ent          # most recently used last
        if ent is None:
            # every entry pins a graph + its static input, full-trajectory output and workspace: keep the cache
            # small (LRU) so that sweeping batch sizes / rollout lengths cannot exhaust device memory
            while len(cache) >= self._GRAPH_CACHE_ENTRIES:
                cache.pop(next(iter(cache)))
            torch = A.torch
            static_in = torch.empty_like(t)
            static_out = torch.empty(shape, dtype=t.dtype, device="cuda")
            nbytes = plan.workspace_bytes(batch)
            ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device="cuda")
            static_in.copy_(t)

            def enqueue():
                nat.check(nat.lib().exb_rollout(plan.handle, A.stream_ptr(), batch, n, substeps, flags,
                                                A.ptr(static_in), A.ptr(static_out), A.ptr(ws)))

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                enqueue()  # warm-up outside capture (function attributes, lazy module load)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                enqueue()
            ent = cache[key] = (graph, static_in, static_out, ws)
        graph, static_in, static_out, _ = ent
        static_in.copy_(t)
        graph.replay()
        return static_out.clone()

    # ---- reference API -------------------------------------------------------------------
    def step(self, u):
        """Perform one step of the time integration (exponax/_base_stepper.py:201-220)."""
        return self._step_batched(u)

    def step_fourier(self, u_hat):
        """One step entirely in Fourier space (exponax/_base_stepper.py:222-239)."""
        return self._step_fourier_batched(u_hat)

    def __call__(self, u):
        """One step on a single state `(C, N, .., N)`; validates the shape
        (exponax/_base_stepper.py:241-271)."""
        expected_shape = self._state_shape()
        if tuple(np.shape(u)) != expected_shape:
            raise ValueError(
                f"""Expected shape {expected_shape}, got {tuple(np.shape(u))}. For batched
                 operation use `jax.vmap` on this function."""
            )
        return self.step(u)
