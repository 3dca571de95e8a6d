"""Generic steppers (mirrors exponax/stepper/generic/__init__.py)."""
from ._steppers import (
    DifficultyConvectionStepper,
    DifficultyGradientNormStepper,
    DifficultyLinearStepper,
    DifficultyLinearStepperSimple,
    DifficultyNonlinearStepper,
    DifficultyPolynomialStepper,
    DiffultyLinearStepperSimple,
    GeneralConvectionStepper,
    GeneralGradientNormStepper,
    GeneralLinearStepper,
    GeneralNonlinearStepper,
    GeneralPolynomialStepper,
    GeneralVorticityConvectionStepper,
    NormalizedConvectionStepper,
    NormalizedGradientNormStepper,
    NormalizedLinearStepper,
    NormalizedNonlinearStepper,
    NormalizedPolynomialStepper,
)
from ._utils import (
    denormalize_coefficients,
    denormalize_convection_scale,
    denormalize_gradient_norm_scale,
    denormalize_polynomial_scales,
    extract_normalized_coefficients_from_difficulty,
    extract_normalized_convection_scale_from_difficulty,
    extract_normalized_gradient_norm_scale_from_difficulty,
    normalize_coefficients,
    normalize_convection_scale,
    normalize_gradient_norm_scale,
    normalize_polynomial_scales,
    reduce_normalized_coefficients_to_difficulty,
    reduce_normalized_convection_scale_to_difficulty,
    reduce_normalized_gradient_norm_scale_to_difficulty,
)

__all__ = [n for n in dir() if not n.startswith("_")]
