import numpy as np

from .. import _native as nat
from .. import _spectral as sp
from ._base import BaseNonlinearFun
from ._leray import Leray


class ProjectedConvection3d(BaseNonlinearFun):
    """Rotational-form convection with Leray projection,
    exponax/nonlin_fun/_projected_convection.py:19-136.  curl (prologue), u x omega (row pass)
    and the projection (epilogue) are fused around 6 inverse + 3 forward 3-D transforms."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, derivative_operator,
                 dealiasing_fraction: float = 2 / 3):
        if num_spatial_dims != 3:
            raise ValueError("ProjectedConvection3d only supports 3 spatial dimensions.")
        super().__init__(num_spatial_dims=num_spatial_dims, num_points=num_points,
                         dealiasing_fraction=dealiasing_fraction)
        self.derivative_operator = derivative_operator
        self.leray_projection = Leray(num_spatial_dims=num_spatial_dims, num_points=num_points,
                                      derivative_operator=derivative_operator)

    def _injection(self):
        return None

    def _native_desc(self, num_channels):
        return {"kind": nat.NL_PROJECTED_3D, "injection": self._injection()}

    def __call__(self, u_hat):
        return self._native_call(u_hat)


class ProjectedConvection3dKolmogorov(ProjectedConvection3d):
    """exponax/nonlin_fun/_projected_convection.py:139-226 (injection reproduced as coded:
    one real entry on channel 0 at (0, +k_f, 0), no Hermitian partner)."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, injection_mode: int = 4,
                 injection_scale: float = 1.0, derivative_operator, dealiasing_fraction: float):
        super().__init__(num_spatial_dims, num_points, derivative_operator=derivative_operator,
                         dealiasing_fraction=dealiasing_fraction)
        self.injection_mode = injection_mode
        self.injection_scale = injection_scale
        self._slab = sp.current_slab()
        wavenumbers = sp.build_wavenumbers(num_spatial_dims, num_points, dtype=self._dtype)
        injection_mask = (wavenumbers[0] == 0) & (wavenumbers[1] == injection_mode) & (wavenumbers[2] == 0)
        injection_single = np.where(
            injection_mask[None],
            self._dtype(injection_scale)
            * sp.build_scaling_array(num_spatial_dims, num_points, mode="coef_extraction", dtype=self._dtype),
            self._dtype(0.0),
        ).astype(self._dtype)
        zeros = np.zeros_like(injection_single)
        self.injection = np.concatenate([injection_single, zeros, zeros], axis=0)

    def _injection(self):
        nz = np.argwhere(self.injection[0] != 0)
        if len(nz) == 0:
            return None  # (inside a slab_context: the forced mode lives on another rank)
        assert len(nz) == 1
        idx = tuple(int(i) for i in nz[0])
        val = float(self.injection[0][idx])
        slab = sp.current_slab() if self._slab is None else self._slab
        if slab is not None:  # the kernels compare GLOBAL mode indices
            local = idx[1]
            glob = local * slab[1] + slab[0] if sp.SLAB_CYCLIC else local + slab[0] * (self.num_points // slab[1])
            idx = (idx[0], glob, idx[2])
        return idx, val
