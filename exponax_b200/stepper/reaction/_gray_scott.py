import numpy as np

from ... import _native as nat
from ..._base_stepper import BaseStepper
from ..._spectral import build_laplace_operator
from ...nonlin_fun import BaseNonlinearFun


class GrayScottNonlinearFun(BaseNonlinearFun):
    """Two-species reaction terms `f (1 - u0) - u0 u1^2`, `-(f + k) u1 + u0 u1^2`
    (exponax/stepper/reaction/_gray_scott.py:12-45); fused between the c2r and r2c row passes."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, dealiasing_fraction: float, feed_rate: float,
                 kill_rate: float):
        super().__init__(num_spatial_dims, num_points, dealiasing_fraction=dealiasing_fraction)
        self.feed_rate = feed_rate
        self.kill_rate = kill_rate

    def _native_desc(self, num_channels):
        if num_channels != 2:
            raise ValueError("num_channels must be 2")
        return {"kind": nat.NL_GRAY_SCOTT, "general_scales": (self.feed_rate, self.kill_rate, 0.0)}

    def __call__(self, u_hat):
        return self._native_call(u_hat)


class GrayScott(BaseStepper):
    """Gray-Scott reaction-diffusion (per-channel linear operator, E = C = 2);
    exponax/stepper/reaction/_gray_scott.py:48-180."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity_1: float = 2e-5, diffusivity_2: float = 1e-5, feed_rate: float = 0.04,
                 kill_rate: float = 0.06, order: int = 2, dealiasing_fraction: float = 1 / 2,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.diffusivity_1 = diffusivity_1
        self.diffusivity_2 = diffusivity_2
        self.feed_rate = feed_rate
        self.kill_rate = kill_rate
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=2, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        laplace = build_laplace_operator(derivative_operator, order=2)
        return np.concatenate([t(self.diffusivity_1) * laplace, t(self.diffusivity_2) * laplace])

    def _build_nonlinear_fun(self, derivative_operator):
        return GrayScottNonlinearFun(self.num_spatial_dims, self.num_points, feed_rate=self.feed_rate,
                                     kill_rate=self.kill_rate, dealiasing_fraction=self.dealiasing_fraction)
