"""RepeatedStepper: `num_sub_steps` ETDRK steps with a spectral carry per call
(exponax/_repeated_stepper.py:9-139); maps onto the `substeps` argument of exb_rollout."""
from __future__ import annotations

import numpy as np

from ._base_stepper import BaseStepper
from ._spectral import spatial_shape


class RepeatedStepper:
    def __init__(self, stepper: BaseStepper, num_sub_steps: int):
        self.stepper = stepper
        self.num_sub_steps = num_sub_steps
        self.dt = stepper.dt * num_sub_steps
        self.num_spatial_dims = stepper.num_spatial_dims
        self.domain_extent = stepper.domain_extent
        self.num_points = stepper.num_points
        self.num_channels = stepper.num_channels
        self.dx = stepper.dx

    def _step_batched(self, u):
        return self.stepper._step_batched(u, substeps=self.num_sub_steps)

    def step(self, u):
        return self._step_batched(u)

    def step_fourier(self, u_hat):
        return self.stepper._step_fourier_batched(u_hat, substeps=self.num_sub_steps)

    def __call__(self, u):
        expected_shape = (self.num_channels,) + spatial_shape(self.num_spatial_dims, self.num_points)
        if tuple(np.shape(u)) != expected_shape:
            raise ValueError(
                f"""Expected shape {expected_shape}, got {tuple(np.shape(u))}. For batched
                 operation use `jax.vmap` on this function."""
            )
        return self.step(u)
