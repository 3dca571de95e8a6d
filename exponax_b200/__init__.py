"""
exponax_b200 -- B200-native (sm_100a) implementation of exponax's forward ETDRK spectral
time-stepping hot path behind exponax's own stepper API.

    import exponax_b200 as ex
    stepper = ex.stepper.KuramotoSivashinskyConservative(1, 100.0, 200, 0.1)
    trj = ex.rollout(stepper, 500, include_init=True)(u0)          # (501, 1, 200)
    trjs = ex.vmap(ex.rollout(stepper, 500))(u0_batch)              # (B, 500, 1, 200)

Arrays may be torch CUDA tensors, anything with `__cuda_array_interface__`/DLPack, or host
NumPy arrays (copied to the device and back).  There is no CPU fallback.
"""
from . import _spectral as spectral
from . import _distributed as distributed
from . import etdrk, ic, metrics, nonlin_fun, stepper
from ._base_stepper import BaseStepper
from ._config import config
from ._ensemble import StepperEnsemble
from ._forced_stepper import ForcedStepper
from ._repeated_stepper import RepeatedStepper
from ._slab import SlabStepper
from ._spectral import derivative, fft, get_spectrum, ifft
from ._utils import build_ic_set, make_grid, repeat, rollout, stack_sub_trajectories, vmap, wrap_bc

__version__ = "0.1.0"

__all__ = [
    "BaseStepper",
    "ForcedStepper",
    "StepperEnsemble",
    "RepeatedStepper",
    "SlabStepper",
    "build_ic_set",
    "config",
    "derivative",
    "distributed",
    "etdrk",
    "fft",
    "get_spectrum",
    "ic",
    "ifft",
    "make_grid",
    "metrics",
    "nonlin_fun",
    "repeat",
    "rollout",
    "spectral",
    "stack_sub_trajectories",
    "stepper",
    "vmap",
    "wrap_bc",
]
