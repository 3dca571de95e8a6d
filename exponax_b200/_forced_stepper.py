"""ForcedStepper (exponax/_forced_stepper.py:7-108): `stepper.step(u + dt * f)`.

The forcing is folded into the fused kernels: `exb_rollout_forced` adds `dt * fft(f)` to the spectral state in front of
every ETDRK step (the transform is linear), so `forced(u, f)`, `ex.rollout(forced, n, takes_aux=True, ...)` and
`ex.repeat(...)` are single native calls like their unforced counterparts.  Steppers without a native plan (user
nonlinear functions) keep the array-level statement `step(u + dt f)`."""
from __future__ import annotations

from . import _array as A
from ._base_stepper import BaseStepper


class ForcedStepper:
    def __init__(self, stepper: BaseStepper):
        self.stepper = stepper

    def _native(self):
        st = self.stepper
        return isinstance(st, BaseStepper) and st._plan_available() and st._plan() is not None

    def step(self, u, f):
        st = self.stepper
        if self._native():
            return st._rollout_batched(u, 1, include_init=False, layout_tb=False, final_only=True, forcing=f,
                                       forcing_constant=True)
        rd = st._dtype
        tu, kind = A.to_device(u, rd)
        tf, _ = A.to_device(f, rd)
        return A.from_device(st._step_batched(tu + st.dt * tf), kind)

    def step_fourier(self, u_hat, f_hat):
        rd = self.stepper._dtype
        tu, kind = A.to_device(u_hat, rd, complex_=True)
        tf, _ = A.to_device(f_hat, rd, complex_=True)
        return A.from_device(self.stepper._step_fourier_batched(tu + self.stepper.dt * tf), kind)

    def __call__(self, u, f):
        return self.step(u, f)
