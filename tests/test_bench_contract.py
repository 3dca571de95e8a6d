"""Host-side checks of bench.py's contract pieces that need no GPU: the algorithmic-bytes pass model quoted in
DESIGN.md / SURVEY section 8d, the synthetic inputs, and the `--impl reference` arm (CPU oracle port)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _w(name):
    return bench.WORKLOADS[name]


def test_algorithmic_bytes_pass_model():
    # c2: the persistent kernel reads u0 once and writes every saved snapshot: 4 N (1 + T) bytes per trajectory
    w = _w("c2")
    assert bench.algorithmic_bytes_per_call(w) == 4 * w["N"] * (1 + w["T"]) * w["B"]
    # c3: 26 F per ETDRK2 step + 4 F per physical carry (every `substeps` steps), F = 4 N^2
    w = _w("c3")
    F = 4 * w["N"] ** 2
    steps, carries = w["T"], w["T"] // w["substeps"]
    assert bench.algorithmic_bytes_per_call(w) == pytest.approx((26 * steps + 4 * carries) * F * w["B"])
    # c4: 90 F per step + 12 F per carry (3 channels in, 3 out, both transforms), F = 4 N^3
    w = _w("c4")
    F = 4 * w["N"] ** 3
    steps, carries = w["T"], w["T"] // w["substeps"]
    assert bench.algorithmic_bytes_per_call(w) == pytest.approx((90 * steps + 12 * carries) * F * w["B"])


def test_synthetic_inputs_are_deterministic_and_bounded():
    for name, B in (("c2", 4), ("c3", 2)):
        a = bench.synth_ic(_w(name), B)
        b = bench.synth_ic(_w(name), B)
        assert a.dtype == np.float32 and np.array_equal(a, b)
        assert a.shape[0] == B and np.all(np.isfinite(a)) and np.abs(a).max() <= 1.16   # c3 / c4 scale sample b by 1 + 0.01 (b % 16)
        assert not np.array_equal(a[0], a[1])


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["metric"] == "ETDRK grid-point*steps/s" and line["unit"] == "grid-point*steps/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0          # no silent CPU fallback


def test_cufft_comparator_restates_the_reference_step():
    """bench.py's "library FFT on the same GPU" comparator (baseline/cufft_standin.py: torch.fft + unfused eager ops)
    must compute the reference's ETDRK2 step -- checked here against the oracle with torch on the CPU."""
    import torch
    import exponax_b200 as ex
    from baseline.cufft_standin import CufftEtdrk2
    from oracle import exponax_np as ox

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    rng = np.random.default_rng(0)
    for name, D, N, C, kw in (("Burgers", 1, 64, 1, dict(diffusivity=0.1)), ("KolmogorovFlowVorticity", 2, 32, 1, {}),
                              ("NavierStokesVelocity", 3, 16, 3, {})):
        st = getattr(ex.stepper, name)(D, 2 * np.pi, N, 0.01, **kw)
        ost = getattr(ox, name)(D, 2 * np.pi, N, 0.01, **kw)
        u0 = (0.3 * rng.standard_normal((2, C) + (N,) * D)).astype(np.float32)
        cf = CufftEtdrk2(st, device="cpu")
        assert rel(cf.step(torch.as_tensor(u0)).numpy(), np.stack([ost(u) for u in u0])) < 2e-6
        assert rel(cf.repeat(torch.as_tensor(u0), 2, substeps=2).numpy(), np.stack([ox.repeat(ost, 4)(u) for u in u0])) < 5e-6
        trj = cf.rollout(torch.as_tensor(u0), 3).numpy()
        assert trj.shape == (2, 3, C) + (N,) * D
        assert rel(trj, np.stack([ox.rollout(ost, 3)(u) for u in u0])) < 5e-6
