from .. import _native as nat
from ._base import BaseNonlinearFun


class ConvectionNonlinearFun(BaseNonlinearFun):
    """Convection nonlinearity, four variants; exponax/nonlin_fun/_convection.py:7-245.

    Fused evaluation: the dealiasing mask and the `i k_d` multiplies are the prologue of the
    inverse transforms, the products `u_d * d_d u_c` (or `u_c u_d`) are formed between the
    last-axis c2r and r2c, and `-scale * mask` (`* 0.5 * sum_d i k_d` when conservative) is the
    epilogue of the forward transform."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, derivative_operator,
                 dealiasing_fraction: float = 2 / 3, scale: float = 1.0, single_channel: bool = False,
                 conservative: bool = False):
        self.derivative_operator = derivative_operator
        self.scale = scale
        self.single_channel = single_channel
        self.conservative = conservative
        super().__init__(num_spatial_dims, num_points, dealiasing_fraction=dealiasing_fraction)

    def _native_desc(self, num_channels):
        if not self.single_channel and num_channels != self.num_spatial_dims:
            raise ValueError("Number of channels in u_hat should match number of spatial dimensions")
        return {"kind": nat.NL_CONVECTION, "scale": self.scale, "single_channel": self.single_channel,
                "conservative": self.conservative}

    def __call__(self, u_hat):
        return self._native_call(u_hat)
