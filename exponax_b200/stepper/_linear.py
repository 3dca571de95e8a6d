"""Linear steppers (always ETDRK0: one complex multiply between two transforms).
exponax/stepper/_advection.py, _diffusion.py, _advection_diffusion.py, _dispersion.py,
_hyper_diffusion.py."""
import numpy as np

from .._base_stepper import BaseStepper
from .._spectral import build_gradient_inner_product_operator, build_laplace_operator
from ..nonlin_fun import ZeroNonlinearFun


def _as_vector(x, D, dtype):
    if isinstance(x, (int, float)):
        return np.ones(D, dtype=dtype) * dtype(x)
    return np.asarray(x, dtype=dtype)


def _as_matrix(x, D, dtype):
    if isinstance(x, (int, float)):
        return np.diag(np.ones(D, dtype=dtype)) * dtype(x)
    x = np.asarray(x, dtype=dtype)
    return np.diag(x) if x.ndim == 1 else x


class _LinearStepper(BaseStepper):
    def __init__(self, num_spatial_dims, domain_extent, num_points, dt):
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=0)

    def _build_nonlinear_fun(self, derivative_operator):
        return ZeroNonlinearFun(self.num_spatial_dims, self.num_points)


class Advection(_LinearStepper):
    """exponax/stepper/_advection.py:13-104."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *, velocity=1.0):
        from .._config import real_dtype
        self.velocity = _as_vector(velocity, num_spatial_dims, real_dtype())
        super().__init__(num_spatial_dims, domain_extent, num_points, dt)

    def _build_linear_operator(self, derivative_operator):
        return -build_gradient_inner_product_operator(derivative_operator, self.velocity, order=1)


class Diffusion(_LinearStepper):
    """exponax/stepper/_diffusion.py:12-122 (scalar, diagonal or full anisotropic diffusivity)."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *, diffusivity=0.01):
        from .._config import real_dtype
        self.diffusivity = _as_matrix(diffusivity, num_spatial_dims, real_dtype())
        super().__init__(num_spatial_dims, domain_extent, num_points, dt)

    def _build_linear_operator(self, derivative_operator):
        laplace_outer_product = derivative_operator[:, None] * derivative_operator[None, :]
        return np.einsum("ij,ij...->...", self.diffusivity, laplace_outer_product)[None, ...]


class AdvectionDiffusion(_LinearStepper):
    """exponax/stepper/_advection_diffusion.py:13-137."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 velocity=1.0, diffusivity=0.01):
        from .._config import real_dtype
        self.velocity = _as_vector(velocity, num_spatial_dims, real_dtype())
        self.diffusivity = _as_matrix(diffusivity, num_spatial_dims, real_dtype())
        super().__init__(num_spatial_dims, domain_extent, num_points, dt)

    def _build_linear_operator(self, derivative_operator):
        laplace_outer_product = derivative_operator[:, None] * derivative_operator[None, :]
        diffusion_operator = np.einsum("ij,ij...->...", self.diffusivity, laplace_outer_product)[None, ...]
        advection_operator = -build_gradient_inner_product_operator(derivative_operator, self.velocity, order=1)
        return advection_operator + diffusion_operator


class Dispersion(_LinearStepper):
    """exponax/stepper/_dispersion.py:13-126."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 dispersivity=1.0, advect_on_diffusion: bool = False):
        from .._config import real_dtype
        self.dispersivity = _as_vector(dispersivity, num_spatial_dims, real_dtype())
        self.advect_on_diffusion = advect_on_diffusion
        super().__init__(num_spatial_dims, domain_extent, num_points, dt)

    def _build_linear_operator(self, derivative_operator):
        if self.advect_on_diffusion:
            laplace_operator = build_laplace_operator(derivative_operator)
            advection_operator = build_gradient_inner_product_operator(
                derivative_operator, self.dispersivity, order=1)
            return advection_operator * laplace_operator
        return build_gradient_inner_product_operator(derivative_operator, self.dispersivity, order=3)


class HyperDiffusion(_LinearStepper):
    """exponax/stepper/_hyper_diffusion.py:8-119."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 hyper_diffusivity: float = 0.0001, diffuse_on_diffuse: bool = False):
        self.hyper_diffusivity = hyper_diffusivity
        self.diffuse_on_diffuse = diffuse_on_diffuse
        super().__init__(num_spatial_dims, domain_extent, num_points, dt)

    def _build_linear_operator(self, derivative_operator):
        t = self._dtype
        if self.diffuse_on_diffuse:
            laplace_operator = build_laplace_operator(derivative_operator)
            return -t(self.hyper_diffusivity) * laplace_operator * laplace_operator
        return -t(self.hyper_diffusivity) * build_laplace_operator(derivative_operator, order=4)
