"""ForcedStepper (exponax/_forced_stepper.py:7-108): `stepper.step(u + dt * f)`."""
from __future__ import annotations

from . import _array as A
from ._base_stepper import BaseStepper


class ForcedStepper:
    def __init__(self, stepper: BaseStepper):
        self.stepper = stepper

    def step(self, u, f):
        rd = self.stepper._dtype
        tu, kind = A.to_device(u, rd)
        tf, _ = A.to_device(f, rd)
        return A.from_device(self.stepper._step_batched(tu + self.stepper.dt * tf), kind)

    def step_fourier(self, u_hat, f_hat):
        rd = self.stepper._dtype
        tu, kind = A.to_device(u_hat, rd, complex_=True)
        tf, _ = A.to_device(f_hat, rd, complex_=True)
        return A.from_device(self.stepper._step_fourier_batched(tu + self.stepper.dt * tf), kind)

    def __call__(self, u, f):
        return self.step(u, f)
