from .._base_stepper import BaseStepper
from .._spectral import build_laplace_operator
from ..nonlin_fun import ConvectionNonlinearFun


class Burgers(BaseStepper):
    """Burgers equation `u_t + b 1/2 (u.grad) u = nu lap u`; constructor arguments, defaults and
    operator assembly follow exponax/stepper/_burgers.py:8-155."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity: float = 0.1, convection_scale: float = 1.0, single_channel: bool = False,
                 conservative: bool = False, order=2, dealiasing_fraction: float = 2 / 3,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.diffusivity = diffusivity
        self.convection_scale = convection_scale
        self.single_channel = single_channel
        self.conservative = conservative
        self.dealiasing_fraction = dealiasing_fraction
        num_channels = 1 if single_channel else num_spatial_dims
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=num_channels, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        return self._dtype(self.diffusivity) * build_laplace_operator(derivative_operator)

    def _build_nonlinear_fun(self, derivative_operator):
        return ConvectionNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.convection_scale,
            single_channel=self.single_channel, conservative=self.conservative)
