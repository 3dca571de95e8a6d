"""
Generic steppers: the linear operator is `sum_i a_i * sum_d (d/dx_d)^i`, the nonlinear part one of
the built-in nonlinear functions.  Pure constructor arithmetic on top of the same kernels; class
names, arguments and defaults follow exponax/stepper/generic/_linear.py, _convection.py,
_gradient_norm.py, _polynomial.py, _nonlinear.py and _vorticity_convection.py.  Each physical
("General"), unit-less ("Normalized": L = 1, dt = 1) and resolution-aware ("Difficulty") variant is
generated from one table instead of being written out three times.
"""
from __future__ import annotations

import warnings

import numpy as np

from ..._base_stepper import BaseStepper
from ...nonlin_fun import (
    ConvectionNonlinearFun,
    GeneralNonlinearFun,
    GradientNormNonlinearFun,
    PolynomialNonlinearFun,
    VorticityConvection2d,
    VorticityConvection2dKolmogorov,
    ZeroNonlinearFun,
)
from . import _utils as U


class _GenericLinearPart(BaseStepper):
    """L_hat = sum_i a_i * sum_d (i k_d)^i  (for i = 0 the sum over d contributes a factor D, as in
    the reference: `jnp.sum(c * derivative_operator**i, axis=0)`)."""

    linear_coefficients: tuple

    def _build_linear_operator(self, derivative_operator):
        cd = derivative_operator.dtype
        rd = derivative_operator.real.dtype.type
        op = np.zeros((1,) + derivative_operator.shape[1:], dtype=cd)
        for i, c in enumerate(self.linear_coefficients):
            op = op + np.sum(rd(c) * derivative_operator**i, axis=0, keepdims=True)
        return op.astype(cd)


_ETDRK_KW = ("order", "dealiasing_fraction", "num_circle_points", "circle_radius")


# ------------------------------------------------------------------------------------- linear
class GeneralLinearStepper(_GenericLinearPart):
    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 linear_coefficients=(0.0, -0.1, 0.01)):
        self.linear_coefficients = linear_coefficients
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=0)

    def _build_nonlinear_fun(self, derivative_operator):
        return ZeroNonlinearFun(self.num_spatial_dims, self.num_points)


class NormalizedLinearStepper(GeneralLinearStepper):
    def __init__(self, num_spatial_dims: int, num_points: int, *, normalized_linear_coefficients=(0.0, -0.5, 0.01)):
        self.normalized_linear_coefficients = normalized_linear_coefficients
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=1.0, num_points=num_points, dt=1.0,
                         linear_coefficients=normalized_linear_coefficients)


class DifficultyLinearStepper(NormalizedLinearStepper):
    def __init__(self, num_spatial_dims: int = 1, num_points: int = 48, *, linear_difficulties=(0.0, -2.0)):
        self.linear_difficulties = linear_difficulties
        super().__init__(num_spatial_dims=num_spatial_dims, num_points=num_points,
                         normalized_linear_coefficients=U.extract_normalized_coefficients_from_difficulty(
                             linear_difficulties, num_spatial_dims=num_spatial_dims, num_points=num_points))


class DifficultyLinearStepperSimple(DifficultyLinearStepper):
    def __init__(self, num_spatial_dims: int = 1, num_points: int = 48, *, difficulty: float = -2.0, order: int = 1):
        super().__init__(linear_difficulties=(0.0,) * order + (difficulty,), num_spatial_dims=num_spatial_dims,
                         num_points=num_points)


def DiffultyLinearStepperSimple(*args, **kwargs):
    warnings.warn("`DiffultyLinearStepperSimple` is deprecated due to a typo. Use "
                  "`DifficultyLinearStepperSimple` instead.", DeprecationWarning, stacklevel=2)
    return DifficultyLinearStepperSimple(*args, **kwargs)


# --------------------------------------------------------------------------------- convection
class GeneralConvectionStepper(_GenericLinearPart):
    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 linear_coefficients=(0.0, 0.0, 0.01), convection_scale: float = 1.0, single_channel: bool = False,
                 conservative: bool = False, order=2, dealiasing_fraction: float = 2 / 3,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.linear_coefficients = linear_coefficients
        self.convection_scale = convection_scale
        self.single_channel = single_channel
        self.dealiasing_fraction = dealiasing_fraction
        self.conservative = conservative
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1 if single_channel else num_spatial_dims, order=order,
                         num_circle_points=num_circle_points, circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        return ConvectionNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.convection_scale,
            single_channel=self.single_channel, conservative=self.conservative)


class NormalizedConvectionStepper(GeneralConvectionStepper):
    def __init__(self, num_spatial_dims: int, num_points: int, *,
                 normalized_linear_coefficients=(0.0, 0.0, 0.01 * 0.1), normalized_convection_scale: float = 1.0 * 0.1,
                 single_channel: bool = False, conservative: bool = False, order: int = 2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        self.normalized_linear_coefficients = normalized_linear_coefficients
        self.normalized_convection_scale = normalized_convection_scale
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=1.0, num_points=num_points, dt=1.0,
                         linear_coefficients=normalized_linear_coefficients,
                         convection_scale=normalized_convection_scale, order=order,
                         dealiasing_fraction=dealiasing_fraction, num_circle_points=num_circle_points,
                         circle_radius=circle_radius, single_channel=single_channel, conservative=conservative)


class DifficultyConvectionStepper(NormalizedConvectionStepper):
    def __init__(self, num_spatial_dims: int = 1, num_points: int = 48, *, linear_difficulties=(0.0, 0.0, 4.5),
                 convection_difficulty: float = 5.0, single_channel: bool = False, conservative: bool = False,
                 maximum_absolute: float = 1.0, order: int = 2, dealiasing_fraction: float = 2 / 3,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.linear_difficulties = linear_difficulties
        self.convection_difficulty = convection_difficulty
        kw = dict(num_spatial_dims=num_spatial_dims, num_points=num_points)
        super().__init__(
            normalized_linear_coefficients=U.extract_normalized_coefficients_from_difficulty(linear_difficulties, **kw),
            normalized_convection_scale=U.extract_normalized_convection_scale_from_difficulty(
                convection_difficulty, maximum_absolute=maximum_absolute, **kw),
            single_channel=single_channel, order=order, dealiasing_fraction=dealiasing_fraction,
            num_circle_points=num_circle_points, circle_radius=circle_radius, conservative=conservative, **kw)


# ------------------------------------------------------------------------------ gradient norm
class GeneralGradientNormStepper(_GenericLinearPart):
    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 linear_coefficients=(0.0, 0.0, -1.0, 0.0, -1.0), gradient_norm_scale: float = 1.0, order=2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        self.linear_coefficients = linear_coefficients
        self.gradient_norm_scale = gradient_norm_scale
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        return GradientNormNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction, scale=self.gradient_norm_scale, zero_mode_fix=True)


class NormalizedGradientNormStepper(GeneralGradientNormStepper):
    def __init__(self, num_spatial_dims: int, num_points: int, *,
                 normalized_linear_coefficients=(0.0, 0.0, -1.0 * 0.1 / 60.0**2, 0.0, -1.0 * 0.1 / 60.0**4),
                 normalized_gradient_norm_scale: float = 1.0 * 0.1 / 60.0**2, order: int = 2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        self.normalized_linear_coefficients = normalized_linear_coefficients
        self.normalized_gradient_norm_scale = normalized_gradient_norm_scale
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=1.0, num_points=num_points, dt=1.0,
                         linear_coefficients=normalized_linear_coefficients,
                         gradient_norm_scale=normalized_gradient_norm_scale, order=order,
                         dealiasing_fraction=dealiasing_fraction, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)


class DifficultyGradientNormStepper(NormalizedGradientNormStepper):
    def __init__(self, num_spatial_dims: int = 1, num_points: int = 48, *,
                 linear_difficulties=(0.0, 0.0, -0.128, 0.0, -0.32768), gradient_norm_difficulty: float = 0.064,
                 maximum_absolute: float = 1.0, order: int = 2, dealiasing_fraction: float = 2 / 3,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.linear_difficulties = linear_difficulties
        self.gradient_norm_difficulty = gradient_norm_difficulty
        kw = dict(num_spatial_dims=num_spatial_dims, num_points=num_points)
        super().__init__(
            normalized_linear_coefficients=U.extract_normalized_coefficients_from_difficulty(linear_difficulties, **kw),
            normalized_gradient_norm_scale=U.extract_normalized_gradient_norm_scale_from_difficulty(
                gradient_norm_difficulty, maximum_absolute=maximum_absolute, **kw),
            order=order, dealiasing_fraction=dealiasing_fraction, num_circle_points=num_circle_points,
            circle_radius=circle_radius, **kw)


# --------------------------------------------------------------------------------- polynomial
class GeneralPolynomialStepper(_GenericLinearPart):
    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 linear_coefficients=(10.0, 0.0, 1.0), polynomial_coefficients=(0.0, 0.0, -10.0), order=2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        self.linear_coefficients = linear_coefficients
        self.polynomial_coefficients = polynomial_coefficients
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        return PolynomialNonlinearFun(self.num_spatial_dims, self.num_points,
                                      dealiasing_fraction=self.dealiasing_fraction,
                                      coefficients=self.polynomial_coefficients)


class NormalizedPolynomialStepper(GeneralPolynomialStepper):
    def __init__(self, num_spatial_dims: int, num_points: int, *,
                 normalized_linear_coefficients=(10.0 * 0.001 / (10.0**0), 0.0, 1.0 * 0.001 / (10.0**2)),
                 normalized_polynomial_coefficients=(0.0, 0.0, -10.0 * 0.001), order: int = 2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        self.normalized_linear_coefficients = normalized_linear_coefficients
        self.normalized_polynomial_coefficients = normalized_polynomial_coefficients
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=1.0, num_points=num_points, dt=1.0,
                         linear_coefficients=normalized_linear_coefficients,
                         polynomial_coefficients=normalized_polynomial_coefficients, order=order,
                         dealiasing_fraction=dealiasing_fraction, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)


class DifficultyPolynomialStepper(NormalizedPolynomialStepper):
    def __init__(self, num_spatial_dims: int = 1, num_points: int = 48, *,
                 linear_difficulties=(10.0 * 0.001 / (10.0**0) * 48**0, 0.0, 1.0 * 0.001 / (10.0**2) * 48**2 * 2**1),
                 polynomial_difficulties=(0.0, 0.0, -10.0 * 0.001), order: int = 2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        self.linear_difficulties = linear_difficulties
        self.polynomial_difficulties = polynomial_difficulties
        super().__init__(
            num_spatial_dims=num_spatial_dims, num_points=num_points,
            normalized_linear_coefficients=U.extract_normalized_coefficients_from_difficulty(
                linear_difficulties, num_spatial_dims=num_spatial_dims, num_points=num_points),
            normalized_polynomial_coefficients=polynomial_difficulties, order=order,
            dealiasing_fraction=dealiasing_fraction, num_circle_points=num_circle_points,
            circle_radius=circle_radius)


# ---------------------------------------------------------------------------- general nonlinear
class GeneralNonlinearStepper(_GenericLinearPart):
    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 linear_coefficients=(0.0, 0.0, 0.01), nonlinear_coefficients=(0.0, -1.0, 0.0), order=2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        if len(nonlinear_coefficients) != 3:
            raise ValueError("The nonlinear coefficients list must have exactly 3 elements")
        self.linear_coefficients = linear_coefficients
        self.nonlinear_coefficients = nonlinear_coefficients
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        return GeneralNonlinearFun(
            self.num_spatial_dims, self.num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=self.dealiasing_fraction, scale_list=self.nonlinear_coefficients,
            zero_mode_fix=True)


class NormalizedNonlinearStepper(GeneralNonlinearStepper):
    def __init__(self, num_spatial_dims: int, num_points: int, *, normalized_linear_coefficients=(0.0, 0.0, 0.1 * 0.1),
                 normalized_nonlinear_coefficients=(0.0, -1.0 * 0.1, 0.0), order=2, dealiasing_fraction: float = 2 / 3,
                 num_circle_points: int = 16, circle_radius: float = 1.0):
        self.normalized_linear_coefficients = normalized_linear_coefficients
        self.normalized_nonlinear_coefficients = normalized_nonlinear_coefficients
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=1.0, num_points=num_points, dt=1.0,
                         linear_coefficients=normalized_linear_coefficients,
                         nonlinear_coefficients=normalized_nonlinear_coefficients, order=order,
                         dealiasing_fraction=dealiasing_fraction, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)


class DifficultyNonlinearStepper(NormalizedNonlinearStepper):
    def __init__(self, num_spatial_dims: int = 1, num_points: int = 48, *,
                 linear_difficulties=(0.0, 0.0, 0.1 * 0.1 / 1.0 * 48**2 * 2),
                 nonlinear_difficulties=(0.0, -1.0 * 0.1 / 1.0 * 48, 0.0), maximum_absolute: float = 1.0,
                 order: int = 2, dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16,
                 circle_radius: float = 1.0):
        self.linear_difficulties = linear_difficulties
        self.nonlinear_difficulties = nonlinear_difficulties
        kw = dict(num_spatial_dims=num_spatial_dims, num_points=num_points)
        super().__init__(
            normalized_linear_coefficients=U.extract_normalized_coefficients_from_difficulty(linear_difficulties, **kw),
            normalized_nonlinear_coefficients=U.extract_normalized_nonlinear_scales_from_difficulty(
                nonlinear_difficulties, maximum_absolute=maximum_absolute, **kw),
            order=order, dealiasing_fraction=dealiasing_fraction, num_circle_points=num_circle_points,
            circle_radius=circle_radius, **kw)


# ------------------------------------------------------------------------- vorticity convection
class GeneralVorticityConvectionStepper(_GenericLinearPart):
    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 vorticity_convection_scale: float = 1.0, linear_coefficients=(0.0, 0.0, 0.001),
                 injection_mode: int = 4, injection_scale: float = 0.0, order: int = 2,
                 dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16, circle_radius: float = 1.0):
        if num_spatial_dims != 2:
            raise ValueError(f"Expected num_spatial_dims = 2, got {num_spatial_dims}.")
        self.vorticity_convection_scale = vorticity_convection_scale
        self.linear_coefficients = linear_coefficients
        self.injection_mode = injection_mode
        self.injection_scale = injection_scale
        self.dealiasing_fraction = dealiasing_fraction
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=1, order=order, num_circle_points=num_circle_points,
                         circle_radius=circle_radius)

    def _build_nonlinear_fun(self, derivative_operator):
        common = dict(convection_scale=self.vorticity_convection_scale, derivative_operator=derivative_operator,
                      dealiasing_fraction=self.dealiasing_fraction)
        if self.injection_scale == 0.0:
            return VorticityConvection2d(self.num_spatial_dims, self.num_points, **common)
        return VorticityConvection2dKolmogorov(self.num_spatial_dims, self.num_points,
                                               injection_mode=self.injection_mode,
                                               injection_scale=self.injection_scale, **common)
