from .. import _native as nat
from ._base import BaseNonlinearFun


class GradientNormNonlinearFun(BaseNonlinearFun):
    """exponax/nonlin_fun/_gradient_norm.py:8-101.  The zero-mode fix (subtracting the spatial
    mean of |grad u|^2) is applied as "DC mode of the forward transform := 0"."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, derivative_operator,
                 dealiasing_fraction: float, zero_mode_fix: bool = True, scale: float = 1.0):
        super().__init__(num_spatial_dims, num_points, dealiasing_fraction=dealiasing_fraction)
        self.derivative_operator = derivative_operator
        self.zero_mode_fix = zero_mode_fix
        self.scale = scale

    def _native_desc(self, num_channels):
        return {"kind": nat.NL_GRADIENT_NORM, "scale": self.scale, "zero_mode_fix": self.zero_mode_fix}

    def __call__(self, u_hat):
        return self._native_call(u_hat)

    def _array_call(self, u_hat):
        """exponax/nonlin_fun/_gradient_norm.py:84-101."""
        D = self.num_spatial_dims
        cax = -D - 1
        grad = self.ifft(self._dop() * u_hat.unsqueeze(cax))     # (.., C, D, N..)
        sq = (grad * grad).sum(dim=cax)
        if self.zero_mode_fix:
            sq = sq - sq.mean(dim=tuple(range(-D, 0)), keepdim=True)
        return -self.scale * 0.5 * self.fft(sq)
