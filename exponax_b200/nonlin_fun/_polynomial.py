from .. import _native as nat
from ._base import BaseNonlinearFun


class PolynomialNonlinearFun(BaseNonlinearFun):
    """exponax/nonlin_fun/_polynomial.py:6-76."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, dealiasing_fraction: float,
                 coefficients: tuple[float, ...]):
        super().__init__(num_spatial_dims, num_points, dealiasing_fraction=dealiasing_fraction)
        self.coefficients = coefficients

    def _native_desc(self, num_channels):
        return {"kind": nat.NL_POLYNOMIAL, "poly": tuple(self.coefficients)}

    def __call__(self, u_hat):
        return self._native_call(u_hat)
