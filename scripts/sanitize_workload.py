#!/usr/bin/env python
"""Small invocation of every fast kernel family for compute-sanitizer (racecheck / synccheck / memcheck):
1-D persistent kernel (N = 64, 256), 2-D fast passes incl. the persistent prologue (N = 128), 3-D fast passes incl.
the TMA-staged middle-axis pass (N = 128), generic kernels (N = 24).  Results are compared with the oracle so a
sanitizer run is also a parity run."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import exponax_b200 as ex
from oracle import exponax_np as ox

def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

which = sys.argv[1] if len(sys.argv) > 1 else "all"
L, dt = 2 * np.pi, 0.01
rng = np.random.default_rng(0)
if which in ("all", "1d"):
    for N in (64, 256):
        u0 = np.stack([ox.random_truncated_fourier_series(1, N, cutoff=5, seed=s, max_one=True) for s in range(5)])
        st, ost = ex.stepper.Burgers(1, L, N, dt, diffusivity=0.1), ox.Burgers(1, L, N, dt, diffusivity=0.1)
        got = ex.vmap(ex.rollout(st, 3))(torch.as_tensor(u0, device="cuda")).cpu().numpy()
        ref = np.stack([ox.rollout(ost, 3)(u) for u in u0])
        print("1d", N, rel(got, ref)); assert rel(got, ref) < 1e-5
        # the any-order instance (ETDRK4), a complex post-factor (conservative convection), an odd-derivative operator
        # (complex Nyquist coefficients) and an odd batch
        for name, kw in (("KuramotoSivashinskyConservative", dict(order=4)), ("KortewegDeVries", dict(order=3)),
                         ("KuramotoSivashinsky", dict(order=1)), ("FisherKPP", dict(order=0))):
            cls = getattr(ex.stepper, name, None) or getattr(ex.stepper.reaction, name)
            st, ost = cls(1, 20.0, N, 0.01, **kw), getattr(ox, name)(1, 20.0, N, 0.01, **kw)
            v0 = (0.3 * u0[:3]).astype(np.float32)
            got = ex.vmap(ex.rollout(st, 2))(torch.as_tensor(v0, device="cuda")).cpu().numpy()
            ref = np.stack([ox.rollout(ost, 2)(u) for u in v0])
            print("1d", name, N, rel(got, ref)); assert rel(got, ref) < 2e-5
if which in ("all", "2d"):
    N = 128
    u0 = (0.1 * rng.standard_normal((3, 1, N, N))).astype(np.float32)
    st, ost = ex.stepper.KolmogorovFlowVorticity(2, L, N, dt), ox.KolmogorovFlowVorticity(2, L, N, dt)
    got = ex.vmap(ex.repeat(st, 2))(torch.as_tensor(u0, device="cuda")).cpu().numpy()
    ref = np.stack([ox.repeat(ost, 2)(u) for u in u0])
    print("2d", N, rel(got, ref)); assert rel(got, ref) < 5e-5
    u0 = (0.1 * rng.standard_normal((2, 1, 24, 24))).astype(np.float32)
    st, ost = ex.stepper.KolmogorovFlowVorticity(2, L, 24, dt), ox.KolmogorovFlowVorticity(2, L, 24, dt)
    got = ex.vmap(st)(torch.as_tensor(u0, device="cuda")).cpu().numpy()
    print("2d generic", rel(got, np.stack([ost(u) for u in u0])))
if which in ("all", "3d"):
    N = 128
    u0 = (0.1 * rng.standard_normal((1, 3, N, N, N))).astype(np.float32)
    st, ost = ex.stepper.NavierStokesVelocity(3, L, N, 0.002), ox.NavierStokesVelocity(3, L, N, 0.002)
    got = ex.vmap(st)(torch.as_tensor(u0, device="cuda")).cpu().numpy()
    ref = np.stack([ost(u) for u in u0])
    print("3d", N, rel(got, ref)); assert rel(got, ref) < 1e-5
torch.cuda.synchronize()
print("sanitize workload OK")
