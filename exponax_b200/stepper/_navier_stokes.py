from .._base_stepper import BaseStepper
from .._spectral import build_laplace_operator
from ..nonlin_fun import (
    ProjectedConvection3d,
    ProjectedConvection3dKolmogorov,
    VorticityConvection2d,
    VorticityConvection2dKolmogorov,
)


class _DiffusionDragStepper(BaseStepper):
    # L = nu * laplace + lambda   (exponax/stepper/_navier_stokes.py:135-141, 307-313, 452-456, 579-585)
    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        return (t(self.diffusivity) * build_laplace_operator(derivative_operator, order=2)
                + t(self.drag) * build_laplace_operator(derivative_operator, order=0))


class NavierStokesVorticity(_DiffusionDragStepper):
    """2-D incompressible Navier-Stokes in streamfunction-vorticity form;
    exponax/stepper/_navier_stokes.py:13-153."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity: float = 0.01, vorticity_convection_scale: float = 1.0, drag: float = 0.0,
                 order: int = 2, dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16,
                 circle_radius: float = 1.0):
        if num_spatial_dims != 2:
            raise ValueError(
                f"Expected num_spatial_dims = 2, got {num_spatial_dims}. For 3D, use NavierStokesVelocity instead."
            )
        self.diffusivity = diffusivity
        self.vorticity_convection_scale = vorticity_convection_scale
        self.drag = drag
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        return VorticityConvection2d(
            self.num_spatial_dims, self.num_points, convection_scale=self.vorticity_convection_scale,
            derivative_operator=derivative_operator, dealiasing_fraction=self.dealiasing_fraction)


class KolmogorovFlowVorticity(_DiffusionDragStepper):
    """2-D Kolmogorov flow (forced NS, vorticity form); exponax/stepper/_navier_stokes.py:156-327."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity: float = 0.001, convection_scale: float = 1.0, drag: float = -0.1,
                 injection_mode: int = 4, injection_scale: float = 1.0, order: int = 2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        if num_spatial_dims != 2:
            raise ValueError(
                f"Expected num_spatial_dims = 2, got {num_spatial_dims}. For 3D, use KolmogorovFlowVelocity instead."
            )
        self.diffusivity = diffusivity
        self.convection_scale = convection_scale
        self.drag = drag
        self.injection_mode = injection_mode
        self.injection_scale = injection_scale
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        return VorticityConvection2dKolmogorov(
            self.num_spatial_dims, self.num_points, convection_scale=self.convection_scale,
            injection_mode=self.injection_mode, injection_scale=self.injection_scale,
            derivative_operator=derivative_operator, dealiasing_fraction=self.dealiasing_fraction)


class NavierStokesVelocity(_DiffusionDragStepper):
    """3-D incompressible Navier-Stokes in velocity form (rotational convection + Leray);
    exponax/stepper/_navier_stokes.py:330-463."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity: float = 0.01, drag: float = 0.0, order: int = 2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        if num_spatial_dims != 3:
            raise ValueError(
                f"Expected num_spatial_dims = 3, got {num_spatial_dims}. For 2D, use NavierStokesVorticity instead."
            )
        self.diffusivity = diffusivity
        self.drag = drag
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=3, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        return ProjectedConvection3d(
            num_spatial_dims=self.num_spatial_dims, num_points=self.num_points,
            derivative_operator=derivative_operator, dealiasing_fraction=self.dealiasing_fraction)


class KolmogorovFlowVelocity(_DiffusionDragStepper):
    """3-D Kolmogorov flow in velocity form; exponax/stepper/_navier_stokes.py:466-598."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 diffusivity: float = 0.01, drag: float = 0.0, injection_mode: int = 4,
                 injection_scale: float = 1.0, order: int = 2, dealiasing_fraction: float = 2 / 3,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        if num_spatial_dims != 3:
            raise ValueError(
                f"Expected num_spatial_dims = 3, got {num_spatial_dims}. For 2D, use KolmogorovFlowVorticity instead."
            )
        self.diffusivity = diffusivity
        self.drag = drag
        self.injection_mode = injection_mode
        self.injection_scale = injection_scale
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=3, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        return ProjectedConvection3dKolmogorov(
            self.num_spatial_dims, self.num_points, injection_mode=self.injection_mode,
            injection_scale=self.injection_scale, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction)
