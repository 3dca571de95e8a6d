"""
Loop transforms: `rollout`, `repeat` (exponax/_utils.py:92-254), `make_grid` (:11-66) and a
`vmap` stand-in for `jax.vmap` over the leading batch axis.

`rollout` / `repeat` dispatch on the type of `stepper_fn`: a native stepper (BaseStepper,
RepeatedStepper, or a vmapped one) becomes ONE fused `exb_rollout` call -- for 1-D grids one
persistent kernel that keeps every trajectory resident in shared memory for the whole loop;
any other callable (neural networks, custom steppers, aux-taking steppers) runs the reference's
plain autoregressive loop.
"""
from __future__ import annotations

import numpy as np

from . import _array as A
from ._base_stepper import BaseStepper
from ._config import real_dtype
from ._ensemble import StepperEnsemble
from ._forced_stepper import ForcedStepper
from ._repeated_stepper import RepeatedStepper


def make_grid(num_spatial_dims: int, domain_extent: float, num_points: int, *, full: bool = False,
              zero_centered: bool = False, indexing: str = "ij"):
    """exponax/_utils.py:11-66 (host array)."""
    dt = real_dtype()
    if full:
        grid_1d = np.linspace(0, domain_extent, num_points + 1, endpoint=True).astype(dt)
    else:
        grid_1d = np.linspace(0, domain_extent, num_points, endpoint=False).astype(dt)
    if zero_centered:
        grid_1d = grid_1d - dt(domain_extent / 2)
    return np.stack(np.meshgrid(*([grid_1d] * num_spatial_dims), indexing=indexing))


class vmap:
    """Batch a stepper / rollout function over the leading axis (stand-in for `jax.vmap`)."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, x, *aux):
        fn = self.fn
        if not aux:
            if isinstance(fn, (BaseStepper, RepeatedStepper)):
                return fn._step_batched(x)
            if isinstance(fn, (_Rollout, _Repeat)):
                return fn._call(x, batched=True)
        elif isinstance(fn, (_Rollout, _Repeat)) and fn.takes_aux and _forced_target(fn.stepper_fn) is not None:
            return fn._call(x, *aux, batched=True)      # forced rollouts: one fused call for the whole batch
        elif isinstance(fn, ForcedStepper) and _forced_target(fn) is not None and len(aux) == 1:
            return fn.step(x, aux[0])
        if hasattr(fn, "_batched"):   # functions with a native batched form (ex.get_spectrum, ex.metrics.*)
            return fn._batched(x, *aux)
        outs = [fn(x[i], *(a[i] for a in aux)) for i in range(len(x))]
        if isinstance(outs[0], np.ndarray):
            return np.stack(outs)
        return A.torch.stack(outs)


def stack_sub_trajectories(trj, sub_len: int):
    """Slice a trajectory `(T, ...)` into its `T - sub_len + 1` overlapping windows `(n_windows, sub_len, ...)`
    == exponax.stack_sub_trajectories (exponax/_utils.py:257-313).  `trj` is an array or a (nested) tuple /
    list / dict of arrays with the same number of time steps.  Torch tensors come back as a strided VIEW of
    `trj` (no copy: the window axis reuses the time stride); NumPy arrays as a read-only sliding-window view."""
    if isinstance(trj, (tuple, list, dict)):
        leaves = list(trj.values()) if isinstance(trj, dict) else list(trj)
        flat = []
        def collect(x):
            if isinstance(x, (tuple, list)):
                for y in x:
                    collect(y)
            elif isinstance(x, dict):
                for y in x.values():
                    collect(y)
            else:
                flat.append(x.shape[0])
        collect(leaves)
        if len(set(flat)) != 1:
            raise ValueError("All arrays in trj must have the same number of time steps in the leading axis")
        if isinstance(trj, dict):
            return {k: stack_sub_trajectories(v, sub_len) for k, v in trj.items()}
        return type(trj)(stack_sub_trajectories(v, sub_len) for v in trj)
    n_time = trj.shape[0]
    if sub_len > n_time:
        raise ValueError("n must be smaller than or equal to the number of time steps in trj")
    n_win = n_time - sub_len + 1
    if isinstance(trj, np.ndarray):
        v = np.lib.stride_tricks.sliding_window_view(trj, sub_len, axis=0)   # (n_win, ..., sub_len)
        return np.moveaxis(v, -1, 1)
    st = trj.stride()
    return trj.as_strided((n_win, sub_len) + tuple(trj.shape[1:]), (st[0], st[0]) + tuple(st[1:]))


def wrap_bc(u):
    """Append the periodic image: `(C, N, .., N)` -> `(C, N+1, .., N+1)` == exponax.wrap_bc
    (exponax/_utils.py:69-89; for plotting the full domain)."""
    if isinstance(u, np.ndarray):
        return np.pad(u, ((0, 0),) + ((0, 1),) * (u.ndim - 1), mode="wrap")
    for axis in range(1, u.ndim):
        u = A.torch.cat([u, u.narrow(axis, 0, 1)], dim=axis)
    return u


def build_ic_set(ic_generator, *, num_points: int, num_samples: int, key=0):
    """`num_samples` initial conditions `(S, C, N, .., N)` == exponax.build_ic_set (exponax/_utils.py:316-348).
    Generators of `exponax_b200.ic` produce the whole set in one batched pipeline; any other callable
    `ic_generator(num_points, key=...)` is called sample by sample with the seeds key, key + 1, ..."""
    if hasattr(ic_generator, "batch"):
        return ic_generator.batch(num_points, num_samples, key=key)
    outs = [ic_generator(num_points, key=key + i) for i in range(num_samples)]
    if isinstance(outs[0], np.ndarray):
        return np.stack(outs)
    return A.torch.stack(outs)


def _native_target(stepper_fn):
    """(base stepper, substeps, vmapped?) if `stepper_fn` can run as one fused rollout."""
    vm = False
    if isinstance(stepper_fn, vmap):
        vm, stepper_fn = True, stepper_fn.fn
    sub = 1
    if isinstance(stepper_fn, RepeatedStepper):
        sub, stepper_fn = stepper_fn.num_sub_steps, stepper_fn.stepper
    if isinstance(stepper_fn, StepperEnsemble):
        # an ensemble maps (S, C, N..) -> (S, C, N..): batched by construction, trajectories come out time-major
        return (stepper_fn, sub, True) if stepper_fn._plan() is not None else None
    if isinstance(stepper_fn, BaseStepper) and stepper_fn._plan_available():
        # a subclass that overrides the stepping methods must be honoured: the reference's scan calls
        # `stepper_fn(u)` every iteration (exponax/_utils.py:167-172), the fused kernel would bypass the override
        t = type(stepper_fn)
        if (t.step is BaseStepper.step and t.step_fourier is BaseStepper.step_fourier
                and t.__call__ is BaseStepper.__call__):
            return stepper_fn, sub, vm
    return None


def _forced_target(stepper_fn):
    """(base stepper, vmapped?) if `stepper_fn` is a ForcedStepper (or vmap of one) around a natively planned stepper:
    the forcing then rides inside the fused rollout (exb_rollout_forced)."""
    vm = False
    if isinstance(stepper_fn, vmap):
        vm, stepper_fn = True, stepper_fn.fn
    if isinstance(stepper_fn, ForcedStepper) and type(stepper_fn).step is ForcedStepper.step \
            and type(stepper_fn).__call__ is ForcedStepper.__call__:
        tgt = _native_target(stepper_fn.stepper)
        if tgt is not None and tgt[1] == 1 and not tgt[2] and tgt[0]._plan() is not None:
            return tgt[0], vm
    return None


def _stack(xs):
    if isinstance(xs[0], np.ndarray):
        return np.stack(xs)
    return A.torch.stack(xs)


class _Rollout:
    def __init__(self, stepper_fn, n, include_init, takes_aux, constant_aux, spectral_carry, cuda_graph=False):
        self.stepper_fn, self.n = stepper_fn, n
        self.include_init, self.takes_aux, self.constant_aux = include_init, takes_aux, constant_aux
        self.spectral_carry, self.cuda_graph = spectral_carry, cuda_graph

    def _call(self, u_0, *aux, batched=False):
        if self.takes_aux and len(aux) == 1 and A.torch.cuda.is_available():
            ft = _forced_target(self.stepper_fn)
            if ft is not None:   # ForcedStepper: the forcing is added inside the fused rollout
                st, vm = ft
                return st._rollout_batched(u_0, self.n, include_init=self.include_init, layout_tb=vm and not batched,
                                           final_only=False, forcing=aux[0], forcing_constant=self.constant_aux)
        tgt = None if self.takes_aux else _native_target(self.stepper_fn)
        if tgt is not None and A.torch.cuda.is_available():
            st, sub, vm = tgt
            if not (vm or batched) and tuple(np.shape(u_0)) != st._state_shape():
                raise ValueError(
                    f"""Expected shape {st._state_shape()}, got {tuple(np.shape(u_0))}. For batched
                 operation use `jax.vmap` on this function."""
                )
            # rollout(vmap(stepper)) stacks time first: (T, B, ...); vmap(rollout(stepper)): (B, T, ...)
            return st._rollout_batched(u_0, self.n, include_init=self.include_init, layout_tb=vm and not batched,
                                       final_only=False, substeps=sub, spectral_carry=self.spectral_carry,
                                       cuda_graph=self.cuda_graph)
        # reference loop (exponax/_utils.py:137-186) for arbitrary callables
        if batched:
            return _stack([self._call(u_0[i], *(a[i] for a in aux)) for i in range(len(u_0))])
        u, trj = u_0, []
        for i in range(self.n):
            if self.takes_aux:
                a = aux[0] if self.constant_aux else _tree_index(aux[0], i)
                u = self.stepper_fn(u, a)
            else:
                u = self.stepper_fn(u)
            trj.append(u)
        if self.include_init:
            trj = [u_0 if type(u_0) is type(trj[0]) else _like(u_0, trj[0])] + trj
        return _stack(trj)

    def __call__(self, u_0, *aux):
        return self._call(u_0, *aux)


class _Repeat:
    def __init__(self, stepper_fn, n, takes_aux, constant_aux, spectral_carry, cuda_graph=False):
        self.stepper_fn, self.n = stepper_fn, n
        self.takes_aux, self.constant_aux = takes_aux, constant_aux
        self.spectral_carry, self.cuda_graph = spectral_carry, cuda_graph

    def _call(self, u_0, *aux, batched=False):
        if self.takes_aux and len(aux) == 1 and self.n >= 1 and A.torch.cuda.is_available():
            ft = _forced_target(self.stepper_fn)
            if ft is not None:
                st, vm = ft
                return st._rollout_batched(u_0, self.n, include_init=False, layout_tb=False, final_only=True,
                                           forcing=aux[0], forcing_constant=self.constant_aux)
        tgt = None if self.takes_aux else _native_target(self.stepper_fn)
        if tgt is not None and self.n >= 1 and A.torch.cuda.is_available():
            st, sub, vm = tgt
            if not (vm or batched) and tuple(np.shape(u_0)) != st._state_shape():
                raise ValueError(
                    f"""Expected shape {st._state_shape()}, got {tuple(np.shape(u_0))}. For batched
                 operation use `jax.vmap` on this function."""
                )
            return st._rollout_batched(u_0, self.n, include_init=False, layout_tb=False, final_only=True,
                                       substeps=sub, spectral_carry=self.spectral_carry, cuda_graph=self.cuda_graph)
        if batched:
            return _stack([self._call(u_0[i], *(a[i] for a in aux)) for i in range(len(u_0))])
        u = u_0
        for i in range(self.n):
            if self.takes_aux:
                a = aux[0] if self.constant_aux else _tree_index(aux[0], i)
                u = self.stepper_fn(u, a)
            else:
                u = self.stepper_fn(u)
        return u

    def __call__(self, u_0, *aux):
        return self._call(u_0, *aux)


def _tree_index(aux, i):
    if isinstance(aux, dict):
        return {k: _tree_index(v, i) for k, v in aux.items()}
    if isinstance(aux, (tuple, list)):
        return type(aux)(_tree_index(v, i) for v in aux)
    return aux[i]


def _like(x, ref):
    if isinstance(ref, np.ndarray):
        return np.asarray(x.cpu() if hasattr(x, "cpu") else x)
    return A.torch.as_tensor(x, device=ref.device, dtype=ref.dtype)


def rollout(stepper_fn, n: int, *, include_init: bool = False, takes_aux: bool = False,
            constant_aux: bool = True, spectral_carry: bool = False, cuda_graph: bool = False):
    """Autoregressive rollout returning the stacked trajectory (exponax/_utils.py:92-186).

    `spectral_carry=True` (extension) keeps the carry in Fourier space between saved steps
    instead of the reference's ifft -> fft round trip (identical up to rounding).
    `cuda_graph=True` (extension) captures the launch sequence of the fused call once per input shape
    and replays it (worth it for N-D grids whose many small pass launches are launch-bound)."""
    return _Rollout(stepper_fn, n, include_init, takes_aux, constant_aux, spectral_carry, cuda_graph)


def repeat(stepper_fn, n: int, *, takes_aux: bool = False, constant_aux: bool = True,
           spectral_carry: bool = False, cuda_graph: bool = False):
    """Apply the stepper n times, return only the final state (exponax/_utils.py:189-254)."""
    return _Repeat(stepper_fn, n, takes_aux, constant_aux, spectral_carry, cuda_graph)
