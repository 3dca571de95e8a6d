from ._base_etdrk import BaseETDRK


class ETDRK0(BaseETDRK):
    """Exactly solve a linear PDE in Fourier space, exponax/etdrk/_etdrk_0.py:6-34."""

    order = 0

    def __init__(self, dt: float, linear_operator):
        super().__init__(dt, linear_operator)

    def step_fourier(self, u_hat):
        return self._dev("_exp_term") * u_hat
