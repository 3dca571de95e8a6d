from .. import _native as nat
from ._base import BaseNonlinearFun
from ._convection import ConvectionNonlinearFun
from ._gradient_norm import GradientNormNonlinearFun
from ._polynomial import PolynomialNonlinearFun


class GeneralNonlinearFun(BaseNonlinearFun):
    """b0*u^2 + b1*1/2 (1.grad)u^2 + b2*1/2 |grad u|^2, exponax/nonlin_fun/_general_nonlinear.py
    :8-119.  The reference sums three separate sub-functions (4 + D transforms); here they share
    one set of 1 + D inverse and 2 forward transforms."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, derivative_operator,
                 dealiasing_fraction: float, scale_list: tuple[float, float, float] = (0.0, -1.0, 0.0),
                 zero_mode_fix: bool = True):
        if len(scale_list) != 3:
            raise ValueError("The scale list must have exactly 3 elements")
        self.derivative_operator = derivative_operator
        self.scale_list = tuple(scale_list)
        self.zero_mode_fix = zero_mode_fix
        self.square_nonlinear_fun = PolynomialNonlinearFun(
            num_spatial_dims, num_points, dealiasing_fraction=dealiasing_fraction,
            coefficients=[0.0, 0.0, scale_list[0]])
        self.convection_nonlinear_fun = ConvectionNonlinearFun(
            num_spatial_dims, num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=dealiasing_fraction, scale=-scale_list[1], single_channel=True,
            conservative=True)
        self.gradient_norm_nonlinear_fun = GradientNormNonlinearFun(
            num_spatial_dims, num_points, derivative_operator=derivative_operator,
            dealiasing_fraction=dealiasing_fraction, scale=-scale_list[2], zero_mode_fix=zero_mode_fix)
        super().__init__(num_spatial_dims, num_points, dealiasing_fraction=dealiasing_fraction)

    def _native_desc(self, num_channels):
        return {"kind": nat.NL_GENERAL, "general_scales": self.scale_list, "zero_mode_fix": self.zero_mode_fix}

    def __call__(self, u_hat):
        return self._native_call(u_hat)

    def _array_call(self, u_hat):
        """exponax/nonlin_fun/_general_nonlinear.py:111-119: the sum of the three sub-functions."""
        return (self.square_nonlinear_fun._array_call(u_hat) + self.convection_nonlinear_fun._array_call(u_hat)
                + self.gradient_norm_nonlinear_fun._array_call(u_hat))
