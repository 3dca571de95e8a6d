from ._base_etdrk import BaseETDRK


class ETDRK1(BaseETDRK):
    """exponax/etdrk/_etdrk_1.py:10-82."""

    order = 1

    def __init__(self, dt, linear_operator, nonlinear_fun, *, num_circle_points: int = 16,
                 circle_radius: float = 1.0):
        super().__init__(dt, linear_operator)
        self._nonlinear_fun = nonlinear_fun
        self._needs_half_exp = False
        (self._coef_1,) = self._contour_means([lambda lr, e, eh: (e - 1) / lr], num_circle_points, circle_radius)

    def _coef_list(self):
        return [self._coef_1]

    def step_fourier(self, u_hat):
        return self._dev("_exp_term") * u_hat + self._dev("_coef_1") * self._nonlinear_fun(u_hat)
