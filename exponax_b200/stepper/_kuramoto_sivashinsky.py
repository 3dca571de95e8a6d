from .._base_stepper import BaseStepper
from .._spectral import build_laplace_operator
from ..nonlin_fun import ConvectionNonlinearFun, GradientNormNonlinearFun


class KuramotoSivashinsky(BaseStepper):
    """KS equation in combustion format (gradient-norm nonlinearity with zero-mode fix);
    exponax/stepper/_kuramoto_sivashinsky.py:8-168."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 gradient_norm_scale: float = 1.0, second_order_scale: float = 1.0,
                 fourth_order_scale: float = 1.0, dealiasing_fraction: float = 2 / 3, order: int = 2,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.gradient_norm_scale = gradient_norm_scale
        self.second_order_scale = second_order_scale
        self.fourth_order_scale = fourth_order_scale
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        return (-t(self.second_order_scale) * build_laplace_operator(derivative_operator, order=2)
                - t(self.fourth_order_scale) * build_laplace_operator(derivative_operator, order=4))

    def _build_nonlinear_fun(self, derivative_operator):
        return GradientNormNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction, zero_mode_fix=True, scale=self.gradient_norm_scale)


class KuramotoSivashinskyConservative(BaseStepper):
    """KS equation in conservative (convection) format; exponax/stepper/_kuramoto_sivashinsky.py
    :171-315 (defaults: conservative=True, single_channel=False)."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 convection_scale: float = 1.0, second_order_scale: float = 1.0,
                 fourth_order_scale: float = 1.0, single_channel: bool = False, conservative: bool = True,
                 dealiasing_fraction: float = 2 / 3, order: int = 2, num_circle_points: int = 16,
                 circle_radius: float = 1.0):
        self.convection_scale = convection_scale
        self.second_order_scale = second_order_scale
        self.fourth_order_scale = fourth_order_scale
        self.single_channel = single_channel
        self.conservative = conservative
        self.dealiasing_fraction = dealiasing_fraction
        if num_spatial_dims > 1:
            print("Warning: The KS equation in conservative format does not generalize well to higher dimensions.")
            print("Consider using the combustion format (`exponax.stepper.KuramotoSivashinsky`) instead.")
        num_channels = 1 if single_channel else num_spatial_dims
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=num_channels, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        return (-t(self.second_order_scale) * build_laplace_operator(derivative_operator, order=2)
                - t(self.fourth_order_scale) * build_laplace_operator(derivative_operator, order=4))

    def _build_nonlinear_fun(self, derivative_operator):
        return ConvectionNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.convection_scale,
            single_channel=self.single_channel, conservative=self.conservative)
