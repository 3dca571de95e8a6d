#!/usr/bin/env python
"""Print the measured on-chip peaks (exb_peak_fp32 / exb_peak_smem) of cuda:0 as one JSON line."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from exponax_b200 import _native as nat
torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
f, s = (C.c_double * 2)(), (C.c_double * 2)()
nat.check(nat.lib().exb_peak_fp32(None, f))
nat.check(nat.lib().exb_peak_smem(None, s))
print(json.dumps({"fp32_tflops_ffma": f[0], "fp32_tflops_ffma2": f[1], "smem_gbs_lds64": s[0], "smem_gbs_lds128": s[1]}))
