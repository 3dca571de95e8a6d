"""
Array bridge between the caller's framework and raw device pointers.

Device arrays: torch CUDA tensors, or anything exposing `__cuda_array_interface__` / DLPack
(jax arrays, cupy) -- converted zero-copy with torch.as_tensor / torch.from_dlpack.
Host arrays (numpy): staged through pinned memory, host->device before and device->host after
the native call; results come back as numpy.  torch is plumbing only (memory + streams).
"""
from __future__ import annotations

import numpy as np
import torch

_TORCH_REAL = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}
_TORCH_CPLX = {np.dtype(np.float32): torch.complex64, np.dtype(np.float64): torch.complex128}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "exponax_b200 needs a CUDA device (sm_100a); there is no CPU fallback for the ETDRK hot path."
        )


def real_t(rd):
    return _TORCH_REAL[np.dtype(rd)]


def cplx_t(rd):
    return _TORCH_CPLX[np.dtype(rd)]


def to_device(x, rd, *, complex_=False):
    """-> (contiguous CUDA tensor of the plan dtype, kind) with kind in {'torch','numpy'}."""
    require_cuda()
    want = cplx_t(rd) if complex_ else real_t(rd)
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.cuda()
        kind = "torch"
    elif isinstance(x, np.ndarray) or np.isscalar(x) or isinstance(x, (list, tuple)):
        a = np.ascontiguousarray(np.asarray(x))
        host = torch.from_numpy(a)
        try:
            host = host.pin_memory()
        except RuntimeError:
            pass
        t = host.to("cuda", non_blocking=True)
        kind = "numpy"
    elif hasattr(x, "__cuda_array_interface__"):
        t = torch.as_tensor(x, device="cuda")
        kind = "torch"
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
        t = t if t.is_cuda else t.cuda()
        kind = "torch"
    else:
        raise TypeError(f"unsupported array type {type(x)}")
    if t.device.index != torch.cuda.current_device():
        # plans, workspaces and the stream belong to the CURRENT device; a tensor living on another GPU would be
        # dereferenced by kernels running here (illegal address, or silent peer access with the tables elsewhere)
        raise ValueError(
            f"array lives on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}; "
            f"wrap the call in `with torch.cuda.device({t.device.index}):`")
    if t.dtype != want:
        if complex_ and not t.is_complex():
            t = t.to(real_t(rd)).to(want)
        elif not complex_ and t.is_complex():
            raise TypeError("expected a real-valued array")
        else:
            t = t.to(want)
    return t.contiguous(), kind


def from_device(t: torch.Tensor, kind: str):
    if kind == "numpy":
        return t.cpu().numpy()
    return t


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
