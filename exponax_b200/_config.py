"""Global precision switch, the analogue of `jax.config.update("jax_enable_x64", True)`
(the reference selects f64/c128 that way, exponax/_base_stepper.py:80-82)."""
import numpy as np

_state = {"enable_x64": False}


class _Config:
    def update(self, key: str, value):
        if key in ("enable_x64", "jax_enable_x64"):
            _state["enable_x64"] = bool(value)
        else:
            raise KeyError(key)

    @property
    def enable_x64(self) -> bool:
        return _state["enable_x64"]


config = _Config()


def real_dtype():
    return np.float64 if _state["enable_x64"] else np.float32


def complex_dtype(rd=None):
    rd = real_dtype() if rd is None else rd
    return np.complex128 if np.dtype(rd) == np.float64 else np.complex64
