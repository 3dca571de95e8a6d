from .. import _native as nat
from ._base import BaseNonlinearFun


class ZeroNonlinearFun(BaseNonlinearFun):
    """exponax/nonlin_fun/_zero.py:8-38."""

    def __init__(self, num_spatial_dims: int, num_points: int):
        super().__init__(num_spatial_dims, num_points)

    def _native_desc(self, num_channels):
        return {"kind": nat.NL_ZERO}

    def __call__(self, u_hat):
        return self._native_call(u_hat)
