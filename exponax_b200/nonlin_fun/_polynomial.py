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

    def _array_call(self, u_hat):
        """exponax/nonlin_fun/_polynomial.py:64-76."""
        u = self.ifft(u_hat)
        acc = 0.0 * u
        for k, c in enumerate(self.coefficients):
            acc = acc + c * u**k
        return self.fft(acc)
