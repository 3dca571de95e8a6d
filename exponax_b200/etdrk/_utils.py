import numpy as np

from .._config import complex_dtype, real_dtype


def roots_of_unity(M: int, dtype=None):
    """M points on the unit circle shifted by half a segment, exponax/etdrk/_utils.py:9-23."""
    dtype = real_dtype() if dtype is None else dtype
    return np.exp(2j * np.pi * (np.arange(1, M + 1) - 0.5) / M).astype(complex_dtype(dtype))
