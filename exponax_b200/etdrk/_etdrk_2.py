from ._base_etdrk import BaseETDRK


class ETDRK2(BaseETDRK):
    """exponax/etdrk/_etdrk_2.py:9-102."""

    order = 2

    def __init__(self, dt, linear_operator, nonlinear_fun, *, num_circle_points: int = 16,
                 circle_radius: float = 1.0):
        super().__init__(dt, linear_operator)
        self._nonlinear_fun = nonlinear_fun
        self._needs_half_exp = False
        self._coef_1, self._coef_2 = self._contour_means(
            [lambda lr, e, eh: (e - 1) / lr, lambda lr, e, eh: (e - 1 - lr) / lr**2],
            num_circle_points, circle_radius)

    def _coef_list(self):
        return [self._coef_1, self._coef_2]

    def step_fourier(self, u_hat):
        n0 = self._nonlinear_fun(u_hat)
        a = self._dev("_exp_term") * u_hat + self._dev("_coef_1") * n0
        n1 = self._nonlinear_fun(a)
        return a + self._dev("_coef_2") * (n1 - n0)
