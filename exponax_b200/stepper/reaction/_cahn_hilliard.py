from ... import _native as nat
from ..._base_stepper import BaseStepper
from ..._spectral import build_laplace_operator
from ...nonlin_fun import BaseNonlinearFun


class CahnHilliardNonlinearFun(BaseNonlinearFun):
    """`scale * laplace(u^3)` (exponax/stepper/reaction/_cahn_hilliard.py:12-37); the Laplacian is
    applied from the mode indices in the epilogue of the forward transform."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, derivative_operator, scale: float,
                 dealiasing_fraction: float):
        super().__init__(num_spatial_dims, num_points, dealiasing_fraction=dealiasing_fraction)
        self.derivative_operator = derivative_operator
        self.laplace_operator = build_laplace_operator(derivative_operator)
        self.scale = scale

    def _native_desc(self, num_channels):
        return {"kind": nat.NL_CAHN_HILLIARD, "scale": self.scale}

    def __call__(self, u_hat):
        return self._native_call(u_hat)


class CahnHilliard(BaseStepper):
    """exponax/stepper/reaction/_cahn_hilliard.py:40-158."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity: float = 1e-2, gamma: float = 1e-3, first_order_coefficient: float = -1.0,
                 third_order_coefficient: float = 1.0, order: int = 2, dealiasing_fraction: float = 1 / 2,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.diffusivity = diffusivity
        self.gamma = gamma
        self.first_order_coefficient = first_order_coefficient
        self.third_order_coefficient = third_order_coefficient
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        laplace = build_laplace_operator(derivative_operator, order=2)
        return t(self.diffusivity) * laplace * (t(self.first_order_coefficient) - t(self.gamma) * laplace)

    def _build_nonlinear_fun(self, derivative_operator):
        return CahnHilliardNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.diffusivity * self.third_order_coefficient)
