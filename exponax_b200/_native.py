"""
ctypes binding of libexb.so (include/exb.h).  There is no CPU fallback: if the shared
library is missing or no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EXB_LIB") or os.path.join(_HERE, "libexb.so")  # EXB_LIB: A/B experiments

EXB_F32, EXB_F64 = 0, 1
(NL_ZERO, NL_CONVECTION, NL_GRADIENT_NORM, NL_POLYNOMIAL, NL_VORTICITY_2D, NL_PROJECTED_3D, NL_GENERAL,
 NL_GRAY_SCOTT, NL_CAHN_HILLIARD) = range(9)
ROLLOUT_INCLUDE_INIT, ROLLOUT_LAYOUT_TB, ROLLOUT_FINAL_ONLY, ROLLOUT_SPECTRAL_CARRY = 1, 2, 4, 8
(SLAB_ROW_R2C, SLAB_ROW_C2R, SLAB_COL1_FWD, SLAB_COL1_INV, SLAB_COL0_FWD, SLAB_COL0_INV, SLAB_COL0_INV_PRO,
 SLAB_ROW_NL, SLAB_COL0_FWD_EPI, SLAB_COL1_INV_NL, SLAB_COL1_FWD_NL) = range(11)
SLAB_SEGMENTED = 0x100
EXB_MAX_POLY = 8


class ExbDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("num_spatial_dims", C.c_int32),
        ("num_points", C.c_int32),
        ("num_channels", C.c_int32),
        ("lin_channels", C.c_int32),
        ("order", C.c_int32),
        ("dtype", C.c_int32),
        ("dealias_kmax", C.c_int32),
        ("domain_extent", C.c_double),
        ("nl_kind", C.c_int32),
        ("single_channel", C.c_int32),
        ("conservative", C.c_int32),
        ("zero_mode_fix", C.c_int32),
        ("nl_scale", C.c_double),
        ("n_poly", C.c_int32),
        ("table_sets", C.c_int32),
        ("poly", C.c_double * EXB_MAX_POLY),
        ("general_scales", C.c_double * 3),
        ("has_injection", C.c_int32),
        ("injection_index", C.c_int32 * 3),
        ("injection_value", C.c_double),
        ("exp_term", C.c_void_p),
        ("half_exp_term", C.c_void_p),
        ("coef", C.c_void_p * 6),
        ("slab_nranks", C.c_int32),
        ("slab_rank", C.c_int32),
        ("tables_on_device", C.c_int32),
        ("slab_cyclic", C.c_int32),
        ("reserved2", C.c_int32),
        ("lin_matrix", C.c_int32),
    ]


_lib = None
_lock = threading.Lock()

# every symbol include/exb.h declares: (restype, argtypes)
SYMBOLS = {
    "exb_last_error": (C.c_char_p, []),
    "exb_version": (C.c_char_p, []),
    "exb_plan_create": (C.c_int, [C.POINTER(ExbDesc), C.POINTER(C.c_void_p)]),
    "exb_plan_destroy": (None, [C.c_void_p]),
    "exb_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "exb_fft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "exb_ifft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "exb_nonlinear_fun": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "exb_step_fourier": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "exb_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "exb_rollout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_uint32,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "exb_rollout_forced": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_uint32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double]),
    "exb_leray": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double]),
    "exb_launch_count": (C.c_int64, [C.c_void_p]),
    "exb_plan_fused_ok": (C.c_int, [C.c_void_p]),
    "exb_peak_fp32": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "exb_peak_smem": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "exb_slab_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "exb_slab_inv_pro_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "exb_slab_pass_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                     C.POINTER(C.c_void_p)]),
    "exb_plan_nl_fields": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "exb_plan_field_pitch": (C.c_int, [C.c_void_p]),
    "exb_ic_shape": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_double, C.c_double,
                               C.c_double]),
    "exb_ic_normalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p]),
    "exb_derivative": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_double]),
    "exb_fourier_sums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_int32, C.c_int32,
                                   C.c_double, C.c_double, C.c_void_p]),
    "exb_metric_sums": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                  C.c_void_p]),
    "exb_spectrum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                               C.c_void_p]),
}


def lib():
    """Load libexb.so (built by `__graft_entry__.build()`); raises if it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                        "g.build()'`. exponax_b200 has no CPU fallback."
                    )
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def _torch():
    import torch
    return torch


class ExbError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = lib().exb_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        if rc == -2:
            raise NotImplementedError(msg)
        raise ExbError(f"libexb error {rc}: {msg}")


class Plan:
    """Owns one exb_plan (device tables live until the object is garbage collected)."""

    def __init__(self, *, D, N, C_, E, order, dtype, L, kmax, nl, exp_term, half_exp_term=None, coefs=(),
                 slab=(1, 0), lin_matrix=False, table_sets=1, slab_cyclic=False):
        rd = np.float32 if dtype == np.float32 else np.float64
        cd = np.complex64 if rd == np.float32 else np.complex128
        d = ExbDesc()
        d.struct_size = C.sizeof(ExbDesc)
        d.num_spatial_dims, d.num_points, d.num_channels, d.lin_channels = D, N, C_, E
        d.order = order
        d.dtype = EXB_F32 if rd == np.float32 else EXB_F64
        d.dealias_kmax = kmax
        d.domain_extent = float(L)
        d.nl_kind = nl.get("kind", NL_ZERO)
        d.single_channel = int(nl.get("single_channel", 0))
        d.conservative = int(nl.get("conservative", 0))
        d.zero_mode_fix = int(nl.get("zero_mode_fix", 0))
        d.nl_scale = float(nl.get("scale", 1.0))
        poly = list(nl.get("poly", ()))
        if len(poly) > EXB_MAX_POLY:
            raise NotImplementedError(f"at most {EXB_MAX_POLY} polynomial coefficients are supported")
        d.n_poly = len(poly)
        for i, v in enumerate(poly):
            d.poly[i] = float(v)
        for i, v in enumerate(nl.get("general_scales", (0.0, 0.0, 0.0))):
            d.general_scales[i] = float(v)
        inj = nl.get("injection")
        if inj is not None:
            d.has_injection = 1
            idx = list(inj[0]) + [0] * (3 - len(inj[0]))
            for i in range(3):
                d.injection_index[i] = int(idx[i])
            d.injection_value = float(inj[1])
        d.slab_nranks, d.slab_rank = int(slab[0]), int(slab[1])
        d.lin_matrix = int(bool(lin_matrix))
        d.slab_cyclic = int(bool(slab_cyclic))
        d.table_sets = int(table_sets)
        self._keep = []

        on_device = hasattr(exp_term, "is_cuda") and exp_term.is_cuda
        d.tables_on_device = int(on_device)

        def host(a, dt):
            if on_device:  # torch CUDA tensors: passed by device pointer, kept alive by the plan
                t = a.contiguous()
                assert t.is_cuda and t.dtype == {np.complex64: _torch().complex64, np.complex128: _torch().complex128,
                                                 np.float32: _torch().float32, np.float64: _torch().float64}[dt]
                self._keep.append(t)
                return t.data_ptr()
            a = np.ascontiguousarray(np.asarray(a), dtype=dt)
            self._keep.append(a)
            return a.ctypes.data

        d.exp_term = host(exp_term, cd)
        if half_exp_term is not None:
            d.half_exp_term = host(half_exp_term, cd)
        for i, c in enumerate(coefs):
            d.coef[i] = host(c, rd)
        self.handle = C.c_void_p()
        check(lib().exb_plan_create(C.byref(d), C.byref(self.handle)))
        if not on_device:
            self._keep = None  # tables were copied to the device
        self.dtype = rd
        self.D, self.N, self.C, self.order = D, N, C_, order

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                lib().exb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def workspace_bytes(self, batch: int) -> int:
        return int(lib().exb_workspace_bytes(self.handle, int(batch)))

    def fused_ok(self) -> bool:
        """False for 1-D grids too large for the persistent shared-memory kernel (exb_plan_fused_ok)."""
        return bool(lib().exb_plan_fused_ok(self.handle))

    def launch_count(self) -> int:
        return int(lib().exb_launch_count(self.handle))
