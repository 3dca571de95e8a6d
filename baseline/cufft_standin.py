"""
"Library FFT on the same GPU" comparator: the reference's ETDRK2 step written
the way XLA executes it on a GPU -- one cuFFT call per `rfftn` / `irfftn` (here through `torch.fft`) and unfused
elementwise kernels in between, every one a full HBM round trip.  JAX itself is not installable in this image, so
this is the stand-in for "exponax on JAX-GPU (XLA + cuFFT)"; nothing of libexb.so runs on this path.

The operator / coefficient arrays are the ones the (host-side, reference-style) constructors of `exponax_b200` build
(`exp_term`, `coef_1`, `coef_2`, derivative operator, dealiasing mask, inverse Laplacian): dense per-mode arrays read
from HBM on every use, exactly as the reference holds them (exponax/etdrk/_etdrk_2.py:91-102,
nonlin_fun/_convection.py:140-171, _vorticity_convection.py:78-99, _projected_convection.py:114-136,
_leray.py:114-136, _base.py:99-137, _spectral.py:656,717-721).
"""
from __future__ import annotations

import torch


class CufftEtdrk2:
    """ETDRK2 `step` / `repeat` / `rollout` of one built-in stepper with torch.fft transforms (batched like vmap)."""

    def __init__(self, stepper, device="cuda"):
        def _dev(a):
            return torch.as_tensor(a, device=device)

        self.D = stepper.num_spatial_dims
        self.N = stepper.num_points
        integ = stepper._integrator
        assert integ.order == 2, "the comparator implements the benchmarked order (ETDRK2)"
        nl = integ._nonlinear_fun
        self.kind = type(nl).__name__
        self.E = _dev(integ._exp_term)
        self.c1 = _dev(integ._coef_1)
        self.c2 = _dev(integ._coef_2)
        self.mask = _dev(nl.dealiasing_mask)
        self.dop = _dev(nl.derivative_operator)
        self.dims = tuple(range(-self.D, 0))
        self.shape = (self.N,) * self.D
        if self.kind.startswith("VorticityConvection2d"):
            self.inv_lap = _dev(nl.inv_laplacian)
            self.scale = float(nl.convection_scale)
            self.inj = _dev(nl.injection) if hasattr(nl, "injection") else None
        elif self.kind.startswith("ProjectedConvection3d"):
            self.leray_inv_lap = _dev(nl.leray_projection.inv_laplacian)
            self.inj = _dev(nl.injection) if hasattr(nl, "injection") else None
        elif self.kind == "ConvectionNonlinearFun":
            assert not nl.single_channel and not nl.conservative
            self.scale = float(nl.scale)
        else:
            raise NotImplementedError(self.kind)

    # BaseNonlinearFun.fft / ifft: post- / pre-dealiasing around the library transforms
    def _ifft(self, x_hat):
        return torch.fft.irfftn(self.mask * x_hat, s=self.shape, dim=self.dims)

    def _fft(self, x):
        return self.mask * torch.fft.rfftn(x, dim=self.dims)

    def nonlin(self, u_hat):
        if self.kind.startswith("VorticityConvection2d"):
            psi = self.inv_lap * u_hat
            u = self._ifft(self.dop[1:2] * psi)
            v = self._ifft(-self.dop[0:1] * psi)
            wx = self._ifft(self.dop[0:1] * u_hat)
            wy = self._ifft(self.dop[1:2] * u_hat)
            out = -self.scale * self._fft(u * wx + v * wy)
            return out + self.inj if self.inj is not None else out
        if self.kind.startswith("ProjectedConvection3d"):
            d, v = self.dop, u_hat
            curl_hat = torch.stack([d[1] * v[..., 2, :, :, :] - d[2] * v[..., 1, :, :, :],
                                    d[2] * v[..., 0, :, :, :] - d[0] * v[..., 2, :, :, :],
                                    d[0] * v[..., 1, :, :, :] - d[1] * v[..., 0, :, :, :]], dim=-4)
            w = self._ifft(curl_hat)
            u = self._ifft(v)
            conv = torch.stack([u[..., 1, :, :, :] * w[..., 2, :, :, :] - u[..., 2, :, :, :] * w[..., 1, :, :, :],
                                u[..., 2, :, :, :] * w[..., 0, :, :, :] - u[..., 0, :, :, :] * w[..., 2, :, :, :],
                                u[..., 0, :, :, :] * w[..., 1, :, :, :] - u[..., 1, :, :, :] * w[..., 0, :, :, :]], dim=-4)
            c_hat = self._fft(conv)
            div = (d * c_hat).sum(dim=-4, keepdim=True)
            out = c_hat + d * (-self.leray_inv_lap * div)
            return out + self.inj if self.inj is not None else out
        # multi-channel non-conservative convection (Burgers): (u . grad) u
        C = u_hat.shape[-self.D - 1]
        u = self._ifft(u_hat)
        grad = self._ifft(self.dop.unsqueeze(0) * u_hat.unsqueeze(-self.D - 1))      # (.., C, D, N..)
        conv = (u.unsqueeze(-self.D - 2) * grad).sum(dim=-self.D - 1) if C > 1 else u * grad.squeeze(-self.D - 1)
        return -self.scale * self._fft(conv)

    def step_fourier(self, u_hat):
        n0 = self.nonlin(u_hat)
        a = self.E * u_hat + self.c1 * n0
        n1 = self.nonlin(a)
        return a + self.c2 * (n1 - n0)

    def step(self, u):
        u_hat = torch.fft.rfftn(u, dim=self.dims)
        return torch.fft.irfftn(self.step_fourier(u_hat), s=self.shape, dim=self.dims)

    def repeat(self, u, n, substeps=1):
        """`ex.repeat(ex.RepeatedStepper(stepper, substeps), n)`: substeps with a spectral carry."""
        for _ in range(n):
            u_hat = torch.fft.rfftn(u, dim=self.dims)
            for _ in range(substeps):
                u_hat = self.step_fourier(u_hat)
            u = torch.fft.irfftn(u_hat, s=self.shape, dim=self.dims)
        return u

    def rollout(self, u, n, out=None):
        """`ex.rollout(stepper, n)` batched: (B, C, N..) -> (B, n, C, N..)."""
        if out is None:
            out = torch.empty((u.shape[0], n) + tuple(u.shape[1:]), dtype=u.dtype, device=u.device)
        for t in range(n):
            u = self.step(u)
            out[:, t] = u
        return out
