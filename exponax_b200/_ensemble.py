"""
Stepper ensembles: many steppers of the same kind that differ in their constructor arguments (viscosity, dt, domain
size, ...) advanced together, one trajectory (or a group of trajectories) per member.

In the reference this is `eqx.filter_vmap` over the constructor (docs/examples/performance_hints.ipynb, "batch over
the stepper": the coefficient arrays `_exp_term`, `_coef_1`, ... of the resulting stepper carry a leading ensemble axis)
followed by `eqx.filter_vmap(lambda stepper, u: stepper(u))(steppers, u0s)`.  Here the members' ETDRK tables are
stacked into ONE plan (`exb_desc.table_sets`): trajectory `b` of a batch reads table set `b / (batch / n_members)`, so
the whole ensemble -- and whole rollouts of it -- run as single fused calls.

What may differ between members: everything that only enters the linear operator / the time step (the tables).  The
nonlinear function's own parameters (kind, scales, dealiasing cutoff, injection) and the grid are shared by
construction and validated.  
"""
from __future__ import annotations

import numpy as np

from . import _array as A
from . import _native as nat
from ._base_stepper import BaseStepper


class StepperEnsemble(BaseStepper):
    """`StepperEnsemble([stepper_0, ..., stepper_{S-1}])`; call it on `(S, C, N, .., N)` -- or `(S * r, C, N, ..)`, then
    member `s` advances the trajectories `[s * r, (s + 1) * r)`.  Works with `ex.rollout` / `ex.repeat` (time-major
    trajectories `(T, S, C, N, ..)`, as for `ex.rollout(ex.vmap(stepper), n)`) and `ex.RepeatedStepper`."""

    def __init__(self, steppers):
        steppers = list(steppers)
        if not steppers:
            raise ValueError("an ensemble needs at least one stepper")
        ref = steppers[0]
        for st in steppers:
            if not isinstance(st, BaseStepper) or type(st) is StepperEnsemble:
                raise TypeError("ensemble members must be (non-ensemble) BaseStepper instances")
            same = (st.num_spatial_dims == ref.num_spatial_dims and st.num_points == ref.num_points
                    and st.num_channels == ref.num_channels and st._order == ref._order and st._dtype == ref._dtype
                    and st._integrator._linear_operator.shape == ref._integrator._linear_operator.shape)
            if not same:
                raise ValueError("ensemble members must share grid, channels, ETDRK order and precision")
            if not st._plan_available():
                raise NotImplementedError("ensemble members need a native nonlinear function")
        self.members = steppers
        self.num_spatial_dims, self.num_points, self.num_channels = ref.num_spatial_dims, ref.num_points, ref.num_channels
        self.domain_extent, self.dx = ref.domain_extent, ref.dx
        self.dt = ref.dt            # (members may differ; the tables carry each member's own dt)
        self._dtype, self._order = ref._dtype, ref._order
        self._integrator, self._nonlinear_fun = ref._integrator, ref._nonlinear_fun
        self._native, self._slab = True, None
        self._ens_plans = {}
        descs = [self._desc_of(st) for st in steppers]
        if any(d != descs[0] for d in descs):
            raise ValueError("ensemble members must share the nonlinear function's parameters (kind, scales, "
                             "dealiasing cutoff, domain extent); only the linear operator / dt may differ")

    @staticmethod
    def _desc_of(st):
        if st._order == 0:
            return ({"kind": nat.NL_ZERO}, -1, float(st.domain_extent))
        nl = st._nonlinear_fun
        return (nl._native_desc(st.num_channels), nl._kmax, float(st.domain_extent))

    def _build_linear_operator(self, derivative_operator):  # pragma: no cover - members are built already
        raise NotImplementedError

    def _build_nonlinear_fun(self, derivative_operator):  # pragma: no cover
        raise NotImplementedError

    def _plan_available(self) -> bool:
        return True

    def _plan(self):
        dev = A.torch.cuda.current_device()
        p = self._ens_plans.get(dev)
        if p is None:
            its = [st._integrator for st in self.members]
            desc, kmax, L = self._desc_of(self.members[0])

            def stack(get):
                return np.ascontiguousarray(np.stack([np.asarray(get(it)) for it in its]))

            half = None if its[0]._half_exp() is None else stack(lambda it: it._half_exp())
            ncoef = len(its[0]._coef_list())
            p = self._ens_plans[dev] = nat.Plan(
                D=self.num_spatial_dims, N=self.num_points, C_=self.num_channels, E=its[0]._linear_operator.shape[0],
                order=its[0].order, dtype=self._dtype, L=L, kmax=kmax, nl=desc, exp_term=stack(lambda it: it._exp_term),
                half_exp_term=half, coefs=[stack(lambda it, i=i: it._coef_list()[i]) for i in range(ncoef)],
                table_sets=len(its))
        return p if p.fused_ok() else None

    def _split_batch(self, t, shape):
        lead, batch = super()._split_batch(t, shape)
        if batch % len(self.members):
            raise ValueError(f"an ensemble of {len(self.members)} steppers needs a leading axis that is a multiple of "
                             f"{len(self.members)}, got {tuple(t.shape)}")
        return lead, batch

    def __call__(self, u):
        """One step of every member on its own state(s): `(S [* r], C, N, .., N)` -> same shape."""
        return self._step_batched(u)
