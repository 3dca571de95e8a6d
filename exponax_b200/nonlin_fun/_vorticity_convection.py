import numpy as np

from .. import _native as nat
from .. import _spectral as sp
from ._base import BaseNonlinearFun


class VorticityConvection2d(BaseNonlinearFun):
    """exponax/nonlin_fun/_vorticity_convection.py:12-99.  The inverse Laplacian (set to 1 at
    k = 0) and the four derivative multiplies are computed from the mode indices in the
    prologue of the inverse column pass -- no operator arrays are read from HBM."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, convection_scale: float = 1.0,
                 derivative_operator, dealiasing_fraction: float):
        if num_spatial_dims != 2:
            raise ValueError(f"Expected num_spatial_dims = 2, got {num_spatial_dims}.")
        super().__init__(num_spatial_dims, num_points, dealiasing_fraction=dealiasing_fraction)
        self.convection_scale = convection_scale
        self.derivative_operator = derivative_operator
        laplacian = sp.build_laplace_operator(derivative_operator, order=2)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.inv_laplacian = np.where(laplacian == 0, 1.0, 1 / laplacian).astype(laplacian.dtype)

    def _injection(self):
        return None

    def _native_desc(self, num_channels):
        return {"kind": nat.NL_VORTICITY_2D, "scale": self.convection_scale, "injection": self._injection()}

    def __call__(self, u_hat):
        return self._native_call(u_hat)


class VorticityConvection2dKolmogorov(VorticityConvection2d):
    """exponax/nonlin_fun/_vorticity_convection.py:102-182."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, convection_scale: float = 1.0,
                 injection_mode: int = 4, injection_scale: float = 1.0, derivative_operator,
                 dealiasing_fraction: float):
        super().__init__(num_spatial_dims, num_points, convection_scale=convection_scale,
                         derivative_operator=derivative_operator, dealiasing_fraction=dealiasing_fraction)
        self.injection_mode = injection_mode
        self.injection_scale = injection_scale
        wavenumbers = sp.build_wavenumbers(num_spatial_dims, num_points, dtype=self._dtype)
        injection_mask = (wavenumbers[0] == 0) & (wavenumbers[1] == injection_mode)
        self.injection = np.where(
            injection_mask,
            self._dtype(-injection_mode * injection_scale)
            * sp.build_scaling_array(num_spatial_dims, num_points, mode="coef_extraction", dtype=self._dtype),
            self._dtype(0.0),
        ).astype(self._dtype)

    def _injection(self):
        nz = np.argwhere(self.injection[0] != 0)
        if len(nz) == 0:
            return None
        assert len(nz) == 1
        idx = tuple(int(i) for i in nz[0])
        return idx, float(self.injection[0][idx])
