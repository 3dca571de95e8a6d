"""Coefficient reparametrisations of the generic steppers: physical <-> normalized
(`alpha_i = a_i * dt / L^i`) <-> difficulty (`gamma_i = alpha_i * N^i * 2^(i-1) * D`).
Formulas: exponax/stepper/generic/_utils.py:1-546."""
from __future__ import annotations


def normalize_coefficients(coefficients, *, domain_extent: float, dt: float):
    return tuple(c * dt / domain_extent**i for i, c in enumerate(coefficients))


def denormalize_coefficients(normalized_coefficients, *, domain_extent: float, dt: float):
    return tuple(c / dt * domain_extent**i for i, c in enumerate(normalized_coefficients))


def normalize_convection_scale(convection_scale: float, *, domain_extent: float, dt: float) -> float:
    return convection_scale * dt / domain_extent


def denormalize_convection_scale(normalized_convection_scale: float, *, domain_extent: float, dt: float) -> float:
    return normalized_convection_scale / dt * domain_extent


def normalize_gradient_norm_scale(gradient_norm_scale: float, *, domain_extent: float, dt: float) -> float:
    return gradient_norm_scale * dt / domain_extent**2


def denormalize_gradient_norm_scale(normalized_gradient_norm_scale: float, *, domain_extent: float, dt: float) -> float:
    return normalized_gradient_norm_scale / dt * domain_extent**2


def normalize_polynomial_scales(polynomial_scales, *, domain_extent: float = None, dt: float):
    return tuple(c * dt for c in polynomial_scales)


def denormalize_polynomial_scales(normalized_polynomial_scales, *, domain_extent: float = None, dt: float):
    return tuple(c / dt for c in normalized_polynomial_scales)


def _difficulty_factor(j: int, num_spatial_dims: int, num_points: int) -> float:
    return num_points**j * 2 ** (j - 1) * num_spatial_dims


def reduce_normalized_coefficients_to_difficulty(normalized_coefficients, *, num_spatial_dims: int, num_points: int):
    out = [a * _difficulty_factor(j, num_spatial_dims, num_points) for j, a in enumerate(normalized_coefficients)]
    out[0] = normalized_coefficients[0]  # the zeroth order is not rescaled
    return tuple(out)


def extract_normalized_coefficients_from_difficulty(difficulty_coefficients, *, num_spatial_dims: int, num_points: int):
    out = [g / _difficulty_factor(j, num_spatial_dims, num_points) for j, g in enumerate(difficulty_coefficients)]
    out[0] = difficulty_coefficients[0]
    return tuple(out)


def reduce_normalized_convection_scale_to_difficulty(normalized_convection_scale: float, *, num_spatial_dims: int,
                                                     num_points: int, maximum_absolute: float) -> float:
    return normalized_convection_scale * maximum_absolute * num_points * num_spatial_dims


def extract_normalized_convection_scale_from_difficulty(difficulty_convection_scale: float, *, num_spatial_dims: int,
                                                        num_points: int, maximum_absolute: float) -> float:
    return difficulty_convection_scale / (maximum_absolute * num_points * num_spatial_dims)


def reduce_normalized_gradient_norm_scale_to_difficulty(normalized_gradient_norm_scale: float, *,
                                                        num_spatial_dims: int, num_points: int,
                                                        maximum_absolute: float) -> float:
    return normalized_gradient_norm_scale * maximum_absolute * num_points**2 * num_spatial_dims


def extract_normalized_gradient_norm_scale_from_difficulty(difficulty_gradient_norm_scale: float, *,
                                                           num_spatial_dims: int, num_points: int,
                                                           maximum_absolute: float) -> float:
    return difficulty_gradient_norm_scale / (maximum_absolute * num_points**2 * num_spatial_dims)


def reduce_normalized_nonlinear_scales_to_difficulty(normalized_nonlinear_scales, *, num_spatial_dims: int,
                                                     num_points: int, maximum_absolute: float):
    kw = dict(num_spatial_dims=num_spatial_dims, num_points=num_points, maximum_absolute=maximum_absolute)
    return (normalized_nonlinear_scales[0],
            reduce_normalized_convection_scale_to_difficulty(normalized_nonlinear_scales[1], **kw),
            reduce_normalized_gradient_norm_scale_to_difficulty(normalized_nonlinear_scales[2], **kw))


def extract_normalized_nonlinear_scales_from_difficulty(nonlinear_difficulties, *, num_spatial_dims: int,
                                                        num_points: int, maximum_absolute: float):
    kw = dict(num_spatial_dims=num_spatial_dims, num_points=num_points, maximum_absolute=maximum_absolute)
    return (nonlinear_difficulties[0],
            extract_normalized_convection_scale_from_difficulty(nonlinear_difficulties[1], **kw),
            extract_normalized_gradient_norm_scale_from_difficulty(nonlinear_difficulties[2], **kw))
