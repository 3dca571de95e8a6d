import numpy as np

from .._base_stepper import BaseStepper
from .._spectral import build_gradient_inner_product_operator, build_laplace_operator
from ..nonlin_fun import ConvectionNonlinearFun


class KortewegDeVries(BaseStepper):
    """KdV equation (complex linear operator: dispersion + hyper-diffusion);
    exponax/stepper/_korteweg_de_vries.py:16-216."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 convection_scale: float = -6.0, diffusivity: float = 0.0, dispersivity: float = 1.0,
                 hyper_diffusivity: float = 0.01, advect_over_diffuse: bool = False,
                 diffuse_over_diffuse: bool = False, single_channel: bool = False, conservative: bool = False,
                 order: int = 2, dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16,
                 circle_radius: float = 1.0):
        self.convection_scale = convection_scale
        self.diffusivity = diffusivity
        self.dispersivity = dispersivity
        self.hyper_diffusivity = hyper_diffusivity
        self.advect_over_diffuse = advect_over_diffuse
        self.diffuse_over_diffuse = diffuse_over_diffuse
        self.single_channel = single_channel
        self.conservative = conservative
        self.dealiasing_fraction = dealiasing_fraction
        num_channels = 1 if single_channel else num_spatial_dims
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=num_channels, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        dispersion_velocity = t(self.dispersivity) * np.ones(self.num_spatial_dims, dtype=t)
        laplace_operator = build_laplace_operator(derivative_operator, order=2)
        diffusion_operator = t(self.diffusivity) * laplace_operator
        if self.advect_over_diffuse:
            dispersion_operator = (
                -build_gradient_inner_product_operator(derivative_operator, dispersion_velocity, order=1)
                * laplace_operator)
        else:
            dispersion_operator = -build_gradient_inner_product_operator(
                derivative_operator, dispersion_velocity, order=3)
        if self.diffuse_over_diffuse:
            hyper_diffusion_operator = -t(self.hyper_diffusivity) * laplace_operator * laplace_operator
        else:
            hyper_diffusion_operator = -t(self.hyper_diffusivity) * build_laplace_operator(
                derivative_operator, order=4)
        return diffusion_operator + dispersion_operator + hyper_diffusion_operator

    def _build_nonlinear_fun(self, derivative_operator):
        return ConvectionNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.convection_scale,
            single_channel=self.single_channel, conservative=self.conservative)
