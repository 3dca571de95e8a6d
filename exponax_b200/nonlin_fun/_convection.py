from .. import _array as A
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

    def _array_call(self, u_hat):
        """exponax/nonlin_fun/_convection.py:140-245, the four variants, on (.., C, N.., N//2+1) tensors."""
        t, D = A.torch, self.num_spatial_dims
        dop = self._dop()
        cax = -D - 1
        if self.single_channel and self.conservative:        # -s/2 * sum_d d_d F[u^2]
            u = self.ifft(u_hat)
            return -self.scale * 0.5 * dop.sum(dim=0, keepdim=True) * self.fft(u * u)
        if self.single_channel:                              # -s * F[u * sum_d d_d u]
            u = self.ifft(u_hat)
            grad = self.ifft(dop * u_hat)                    # (.., D, N..): u_hat has one channel
            return -self.scale * self.fft(u * grad.sum(dim=cax, keepdim=True))
        if self.conservative:                                # -s/2 * sum_d d_d F[u_c u_d]
            u = self.ifft(u_hat)
            outer = u.unsqueeze(cax) * u.unsqueeze(cax - 1)  # (.., C, D, N..)
            return -self.scale * 0.5 * (dop * self.fft(outer)).sum(dim=cax)
        u = self.ifft(u_hat)                                 # -s * F[sum_d u_d d_d u_c]
        grad = self.ifft(dop * u_hat.unsqueeze(cax))         # (.., C, D, N..)
        return -self.scale * self.fft((u.unsqueeze(cax - 1) * grad).sum(dim=cax))
