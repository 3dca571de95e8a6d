"""Reaction-diffusion steppers with polynomial nonlinearity.
exponax/stepper/reaction/_fisher_kpp.py, _allen_cahn.py, _swift_hohenberg.py."""
from ..._base_stepper import BaseStepper
from ..._spectral import build_laplace_operator
from ...nonlin_fun import PolynomialNonlinearFun


class FisherKPP(BaseStepper):
    """exponax/stepper/reaction/_fisher_kpp.py:8-129."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity: float = 0.01, reactivity=1.0, order: int = 2, dealiasing_fraction: float = 2 / 3,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.dealiasing_fraction = dealiasing_fraction
        self.diffusivity = diffusivity
        self.reactivity = reactivity
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        laplace = build_laplace_operator(derivative_operator, order=2)
        return t(self.diffusivity) * laplace + t(self.reactivity)

    def _build_nonlinear_fun(self, derivative_operator):
        return PolynomialNonlinearFun(self.num_spatial_dims, self.num_points,
                                      dealiasing_fraction=self.dealiasing_fraction,
                                      coefficients=[0.0, 0.0, -self.reactivity])


class AllenCahn(BaseStepper):
    """exponax/stepper/reaction/_allen_cahn.py:8-128 (default dealiasing 1/2: cubic term)."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity: float = 5e-3, first_order_coefficient: float = 1.0,
                 third_order_coefficient: float = -1.0, order: int = 2, dealiasing_fraction: float = 1 / 2,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.diffusivity = diffusivity
        self.first_order_coefficient = first_order_coefficient
        self.third_order_coefficient = third_order_coefficient
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        laplace = build_laplace_operator(derivative_operator, order=2)
        return t(self.diffusivity) * laplace + t(self.first_order_coefficient)

    def _build_nonlinear_fun(self, derivative_operator):
        return PolynomialNonlinearFun(self.num_spatial_dims, self.num_points,
                                      dealiasing_fraction=self.dealiasing_fraction,
                                      coefficients=[0.0, 0.0, 0.0, self.third_order_coefficient])


class SwiftHohenberg(BaseStepper):
    """exponax/stepper/reaction/_swift_hohenberg.py:8-128."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 reactivity: float = 0.7, critical_number: float = 1.0,
                 polynomial_coefficients: tuple[float, ...] = (0.0, 0.0, 1.0, -1.0), order: int = 2,
                 dealiasing_fraction: float = 1 / 2, num_circle_points: int = 16, circle_radius: float = 1.0):
        self.reactivity = reactivity
        self.critical_number = critical_number
        self.polynomial_coefficients = polynomial_coefficients
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        laplace = build_laplace_operator(derivative_operator, order=2)
        return t(self.reactivity) - (t(self.critical_number) + laplace) ** 2

    def _build_nonlinear_fun(self, derivative_operator):
        return PolynomialNonlinearFun(self.num_spatial_dims, self.num_points,
                                      dealiasing_fraction=self.dealiasing_fraction,
                                      coefficients=self.polynomial_coefficients)
