import numpy as np

from .. import _array as A
from .._base_stepper import BaseStepper
from .._spectral import build_scaled_wavenumbers
from ..nonlin_fun import ZeroNonlinearFun


class Wave(BaseStepper):
    """Second-order wave equation as a first-order system `u = (h, v)`: rotate into the two
    travelling-wave modes, advance them exactly with `exp(+-i c |k| dt)`, rotate back, add
    `dt * v_0` to the mean of `h` (exponax/stepper/_wave.py:14-197).

    The three per-mode 2x2 maps (rotation, diagonal exponential, inverse rotation) plus the
    DC correction are composed ONCE at construction time into a single per-mode 2x2 matrix
    (obtained by pushing the two basis states through the reference's own sequence of operations,
    so the rounding is the reference's), and one step in Fourier space is that matrix applied to
    `(h_hat, v_hat)` -- natively: the plan carries the matrix as its order-0 "linear operator"
    (`exb_desc.lin_matrix`), so the step runs inside the fused kernels like every other stepper
    ("next" row).  `_step_fourier_generic` is the array-level statement of the same map."""

    def __init__(self, num_spatial_dims: int, domain_extent: float, num_points: int, dt: float, *,
                 speed_of_sound: float = 1.0):
        self.speed_of_sound = speed_of_sound
        from .._config import real_dtype
        rd = real_dtype()
        self.wavenumber_norm = np.linalg.norm(
            build_scaled_wavenumbers(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent,
                                     num_points=num_points, dtype=rd), axis=0, keepdims=True).astype(rd)
        super().__init__(num_spatial_dims=num_spatial_dims, domain_extent=domain_extent, num_points=num_points,
                         dt=dt, num_channels=2, order=0)
        cd = self._integrator._cd
        shape = self.wavenumber_norm.shape[1:]
        cols = []
        for basis in ((1.0, 0.0), (0.0, 1.0)):
            e = np.stack([np.full(shape, basis[0], cd), np.full(shape, basis[1], cd)])
            cols.append(self._step_fourier_host(e))
        # _matrix[i][j]: contribution of input channel j to output channel i
        self._matrix = np.stack([np.stack([cols[0][i], cols[1][i]]) for i in range(2)]).astype(cd)
        self._matrix_dev = {}
        self._wave_plans = {}

    def _plan(self):
        """Native plan: an order-0 step whose per-mode factor is the composed 2 x 2 matrix (`exb_desc.lin_matrix`),
        so `stepper(u)`, `ex.rollout` / `ex.repeat` and `RepeatedStepper` run inside the fused `exb_rollout`."""
        from .. import _native as nat
        dev = A.torch.cuda.current_device()
        p = self._wave_plans.get(dev)
        if p is None:
            M = int(np.prod(self._matrix.shape[2:]))
            p = self._wave_plans[dev] = nat.Plan(
                D=self.num_spatial_dims, N=self.num_points, C_=2, E=4, order=0, dtype=self._dtype,
                L=self.domain_extent, kmax=-1, nl={"kind": nat.NL_ZERO}, exp_term=self._matrix.reshape(4, M),
                lin_matrix=True)
        self._native = True
        return p if p.fused_ok() else None

    # ---- the reference's sequence of operations, on host arrays (constructor only) ----------------
    def _forward_transform(self, u_hat):
        t = self._dtype
        h_hat, v_hat = u_hat[0:1], u_hat[1:2]
        k_guard = np.where(self.wavenumber_norm == 0, t(1.0), self.wavenumber_norm)
        w_hat = 1j * t(self.speed_of_sound) * k_guard * h_hat
        s = t(1 / np.sqrt(2))
        return np.concatenate([s * (w_hat + v_hat), s * (w_hat - v_hat)], axis=0)

    def _inverse_transform(self, waves_hat):
        t = self._dtype
        pos, neg = waves_hat[0:1], waves_hat[1:2]
        s = t(1 / np.sqrt(2))
        w_hat = s * (pos + neg)
        v_hat = s * (pos - neg)
        k_guard = np.where(self.wavenumber_norm == 0, t(1.0), self.wavenumber_norm)
        h_hat = w_hat / (1j * t(self.speed_of_sound) * k_guard)
        return np.concatenate([h_hat, v_hat], axis=0)

    def _step_fourier_host(self, u_hat):
        cd = self._integrator._cd
        waves = self._forward_transform(u_hat).astype(cd)
        waves_next = (self._integrator._exp_term * waves).astype(cd)
        nxt = self._inverse_transform(waves_next).astype(cd)
        dc = (0,) * self.num_spatial_dims
        nxt[(0,) + dc] += self._dtype(self.dt) * u_hat[(1,) + dc]
        return nxt

    def _build_linear_operator(self, derivative_operator):
        val = 1j * self._dtype(self.speed_of_sound) * self.wavenumber_norm
        return np.concatenate((val, -val), axis=0)

    def _build_nonlinear_fun(self, derivative_operator):
        return ZeroNonlinearFun(self.num_spatial_dims, self.num_points)

    # ---- device step --------------------------------------------------------------------------------
    def _step_fourier_generic(self, u_hat):
        dev = A.torch.cuda.current_device()
        M = self._matrix_dev.get(dev)
        if M is None:
            M = self._matrix_dev[dev] = A.torch.as_tensor(self._matrix, device="cuda")
        D = self.num_spatial_dims
        h = u_hat.select(-D - 1, 0)
        v = u_hat.select(-D - 1, 1)
        return A.torch.stack([M[0, 0] * h + M[0, 1] * v, M[1, 0] * h + M[1, 1] * v], dim=-D - 1)
