import numpy as np

from .. import _array as A
from .. import _spectral as sp
from ._base import BaseNonlinearFun


class Leray(BaseNonlinearFun):
    """Leray projection, exponax/nonlin_fun/_leray.py:8-136.  Inside
    `ProjectedConvection3d` the projection is fused into the epilogue of the last forward
    pass; called on its own it is one per-mode kernel (`exb_leray`; a Laplacian order other than 2 keeps the
    array-level formulation)."""

    def __init__(self, num_spatial_dims: int, num_points: int, *, derivative_operator, order: int = 2):
        super().__init__(num_spatial_dims=num_spatial_dims, num_points=num_points)
        laplace_operator = sp.build_laplace_operator(derivative_operator, order=order)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.inv_laplacian = np.where(laplace_operator != 0, 1.0 / laplace_operator, 0.0).astype(
                laplace_operator.dtype)
        self.derivative_operator = derivative_operator
        self.order = order

    def __call__(self, u_hat):
        t, kind = A.to_device(u_hat, self._dtype, complex_=True)
        D, N = self.num_spatial_dims, self.num_points
        if self.order == 2 and t.ndim >= D + 1 and t.shape[-D - 1] == D and tuple(t.shape[-D:]) == sp.wavenumber_shape(D, N):
            # native: one per-mode kernel (exb_leray); wavenumbers and the inverse Laplacian come from the mode indices
            from .. import _native as nat
            lead = t.shape[:-D - 1]
            nf = int(np.prod(lead)) if lead else 1
            out = A.torch.empty_like(t)
            plan = sp._plain_plan(D, N, self._dtype)
            nat.check(nat.lib().exb_leray(plan.handle, A.stream_ptr(), nf, A.ptr(t), A.ptr(out), self._domain_extent()))
            return A.from_device(out, kind)
        dop = A.torch.as_tensor(self.derivative_operator, device="cuda")
        ilap = A.torch.as_tensor(self.inv_laplacian, device="cuda")
        D = self.num_spatial_dims
        div = (dop * t).sum(dim=-D - 1, keepdim=True)
        p = -ilap * div
        return A.from_device(t + dop * p, kind)
