import numpy as np

from .. import _array as A
from ._base_etdrk import BaseETDRK


class ETDRK4(BaseETDRK):
    """exponax/etdrk/_etdrk_4.py:9-224 (the code, not its docstring, is the specification)."""

    order = 4

    def __init__(self, dt, linear_operator, nonlinear_fun, *, num_circle_points: int = 16,
                 circle_radius: float = 1.0):
        super().__init__(dt, linear_operator)
        self._nonlinear_fun = nonlinear_fun
        if self._on_gpu:
            self._half_exp_term = A.torch.exp(self._rd(0.5) * self._rd(dt) * self._L_dev)
        else:
            self._half_exp_term = np.exp(self._rd(0.5) * self._rd(dt) * self._linear_operator).astype(self._cd)
        (self._coef_1, self._coef_4, self._coef_5, self._coef_6) = self._contour_means(
            [
                lambda lr, e, eh: (eh - 1) / lr,
                lambda lr, e, eh: (-4 - lr + e * (4 - 3 * lr + lr**2)) / lr**3,
                lambda lr, e, eh: (2 + lr + e * (-2 + lr)) / lr**3,
                lambda lr, e, eh: (-4 - 3 * lr - lr**2 + e * (4 - lr)) / lr**3,
            ],
            num_circle_points, circle_radius)
        self._coef_2 = self._coef_1
        self._coef_3 = self._coef_1

    def _coef_list(self):
        return [self._coef_1, self._coef_2, self._coef_3, self._coef_4, self._coef_5, self._coef_6]

    def _half_exp(self):
        return self._half_exp_term

    def step_fourier(self, u_hat):
        E, Eh = self._dev("_exp_term"), self._dev("_half_exp_term")
        n0 = self._nonlinear_fun(u_hat)
        a = Eh * u_hat + self._dev("_coef_1") * n0
        n1 = self._nonlinear_fun(a)
        b = Eh * u_hat + self._dev("_coef_2") * n1
        n2 = self._nonlinear_fun(b)
        c = Eh * a + self._dev("_coef_3") * (2 * n2 - n0)
        n3 = self._nonlinear_fun(c)
        return (E * u_hat + self._dev("_coef_4") * n0 + self._dev("_coef_5") * 2 * (n1 + n2)
                + self._dev("_coef_6") * n3)
