#!/usr/bin/env python
"""
bench.py -- ETDRK grid-point*steps/s of the exponax hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c1] [--impl reference]

A bench "step" = one pass of the hot path over one batch of synthetic input = ONE fused
rollout call (`ex.vmap(ex.rollout(stepper, T))(u0)` == one `exb_rollout`) of `B` trajectories
x `T` ETDRK steps.  Default workload = BASELINE.json configs[1] (c2): Burgers 1-D ETDRK2,
N=256, B=16384 per GPU, T=1000, every step saved (16.8 GB trajectory per call).

  value  : N^D * B * T * K / device-time, inputs resident in HBM, CUDA events, max over ranks
  e2e    : same call with HOST buffers (pinned), H2D of u0 and D2H of the result inside the timed
           region, through the public Python API
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement"

Multi-GPU: one process per GPU (torchrun), the batch axis is sharded, no data-path collective
(trajectories are independent) -> weak scaling: per-GPU batch fixed, value = all ranks' units /
max-over-ranks time.

`--impl reference` times the reference algorithm on the host cores: the NumPy/SciPy oracle
port of exponax (JAX is not installable here), bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (stepper name, D, L, N, dt, kwargs, C, B per GPU, T steps per call, save mode)
    "c1": dict(stepper="KuramotoSivashinskyConservative", D=1, L=100.0, N=200, dt=0.1, kw={}, C=1, B=1, T=500,
               final_only=False, desc="KS-conservative 1-D N=200 L=100 dt=0.1, 500-step rollout, batch 1"),
    "c2": dict(stepper="Burgers", D=1, L=2 * np.pi, N=256, dt=0.01, kw=dict(diffusivity=0.1), C=1, B=16384, T=1000,
               final_only=False,
               desc="Burgers 1-D ETDRK2 N=256, 16384 trajectories x 1000 steps per GPU, every step saved"),
    "c3": dict(stepper="KolmogorovFlowVorticity", D=2, L=2 * np.pi, N=512, dt=0.01, kw=dict(diffusivity=0.001), C=1,
               B=512, T=20, final_only=True, substeps=10,
               desc="KolmogorovFlowVorticity 2-D 512x512 ETDRK2 2/3-dealiased, batch 512 per GPU, "
                    "ex.repeat(ex.RepeatedStepper(stepper, 10), 2) = 20 ETDRK steps"),
    "c4": dict(stepper="NavierStokesVelocity", D=3, L=2 * np.pi, N=256, dt=0.005, kw=dict(diffusivity=0.01), C=3,
               B=16, T=4, final_only=True, substeps=2,
               desc="NavierStokesVelocity 3-D 256^3 Taylor-Green ETDRK2, batch 16 per GPU, "
                    "ex.repeat(ex.RepeatedStepper(stepper, 2), 2) = 4 ETDRK steps"),
    "readme": dict(stepper="KuramotoSivashinsky", D=2, L=30.0, N=128, dt=0.1, kw={}, C=1, B=50, T=200,
                   final_only=False,
                   desc="README.md:187-189 claim: 50 trajectories of 2-D Kuramoto-Sivashinsky, 128x128, 200 steps "
                        "('under a second on a modern GPU'); every step saved"),
    "c5": dict(stepper="KolmogorovFlowVelocity", D=3, L=2 * np.pi, N=2048, dt=1e-3, kw=dict(diffusivity=0.01), C=3,
               B=1, T=1, final_only=True,
               desc="KolmogorovFlowVelocity 3-D single field, slab-decomposed FFT (needs --gpus >= 2)"),
}


def run_c5(args, w, rank, world, local_rank):
    """Config c5: ONE 3-D field sharded over all ranks (strong scaling by construction)."""
    import torch
    import torch.distributed as dist

    import exponax_b200 as ex

    torch.cuda.set_device(local_rank)
    if world > 1:
        from exponax_b200._slab import init_process_group_nccl
        init_process_group_nccl(local_rank)
    N, L = w["N"], w["L"]
    t0 = time.time()
    # lean slab constructor: operator + ETDRK tables are assembled on the GPU for the local slab only
    slab = ex.SlabStepper.navier_stokes_velocity(L, N, w["dt"], injection_mode=4, **w["kw"])
    slab.plan()
    slab.overlap = not args.no_overlap
    slab.raw_exchange = not args.no_raw_exchange
    if args.no_peer_stores:
        slab.peer_stores = False
    if args.peer_stores:
        slab.peer_stores = True
    if args.exchange:
        slab.ce_exchange = args.exchange == "ce"
    t_ctor = time.time() - t0
    # Taylor-Green + small-mode perturbation generated on the device, slab by slab (never on the host)
    n = N // world
    x = (torch.arange(rank * n, (rank + 1) * n, device="cuda", dtype=torch.float32) * (L / N)).view(n, 1, 1)
    y = (torch.arange(N, device="cuda", dtype=torch.float32) * (L / N)).view(1, N, 1)
    z = (torch.arange(N, device="cuda", dtype=torch.float32) * (L / N)).view(1, 1, N)
    u = torch.empty((3, n, N, N), dtype=torch.float32, device="cuda")
    u[0] = torch.sin(x) * torch.cos(y) * torch.cos(z)
    u[1] = -torch.cos(x) * torch.sin(y) * torch.cos(z)
    u[2] = 0.05 * torch.sin(2 * x) * torch.cos(3 * y) * torch.ones_like(z)
    uh = slab.fft(u)
    del u
    slab.release_buffers()
    for _ in range(args.warmup):
        uh = slab.step_fourier(uh, inplace=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = slab.plan().launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        uh = slab.step_fourier(uh, inplace=True)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    finite = bool(torch.isfinite(torch.view_as_real(uh)).all())
    energy = torch.tensor([float((uh.abs() ** 2).sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(energy)
    launches = slab.plan().launch_count() - launches0
    mem_gb = torch.cuda.max_memory_allocated() / 1e9
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        F = 4 * N**3
        abytes = 90 * F                       # pass model per ETDRK2 step, whole field
        per_gpu = abytes / world / (total_ms / args.steps * 1e-3) / 1e9
        # all-to-all: 2 stages x (6 inverse + 3 forward) field transposes, (P-1)/P of each slab leaves the GPU
        # (payload = the compact field pitch: only the last-axis wavenumbers inside the dealiasing mask are shipped)
        kp = int(getattr(slab, "Kp", N // 2 + 1))
        a2a = 2 * 9 * (N * n * kp * 8) * (world - 1) / max(world, 1)
        a2a_unpruned = 2 * 9 * (N * n * (N // 2 + 1) * 8) * (world - 1) / max(world, 1)
        line = {"metric": "ETDRK grid-point*steps/s", "value": N**3 * args.steps / (total_ms * 1e-3),
                "unit": "grid-point*steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"c5: {w['desc']}", "N": N, "D": 3, "order": 2, "channels": 3,
                           "parallelism": f"slab decomposition x{world}",
                           "carry": "spectral (step_fourier loop)",
                           "transposes": ("copy-engine block copies into the peers' symmetric-memory buffers (cudaMemcpyAsync over "
                                          "NVLink on a second stream) + one barrier per transpose"
                                          if getattr(slab, "ce_exchange", False) and getattr(slab, "_peer", None) is not None
                                          else "fused into the pass kernels' stores over NVLink peer memory (symmetric memory) + barrier"
                                          if slab.peer_stores and getattr(slab, "_peer", None) is not None
                                          else "all_to_all_single (NCCL)" + (", pipelined per field on a second stream" if slab.overlap else "")),
                           "peer_store_fallback_reason": getattr(slab, "_peer_error", None)},
                "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak,
                             "traffic": None, "algorithmic_bytes_per_step_all_gpus": abytes,
                             "alltoall_bytes_per_gpu_per_step": a2a, "alltoall_bytes_unpruned": a2a_unpruned,
                             "axis1_distribution": "cyclic" if getattr(slab, "cyclic", False) else "block",
                             "alltoall_GBs_per_gpu": a2a / (total_ms / args.steps * 1e-3) / 1e9},
                "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "finite": finite,
                "spectral_energy": float(energy.item()), "ctor_seconds": t_ctor, "max_mem_gb_rank0": mem_gb}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def algorithmic_bytes_per_call(w, itemsize=4, spectral_carry=False):
    """SURVEY section 8(d) / BASELINE.md section 3: ALGORITHMIC bytes per rollout call (per GPU)."""
    N, D, C, B, T = w["N"], w["D"], w["C"], w["B"], w["T"]
    F = itemsize * N**D
    if D == 1:
        # persistent kernel: read u0 once, write each saved snapshot once
        saved = 1 if w["final_only"] else T
        return B * C * F * (1 + saved)
    # pass model per ETDRK2 step: stage = [2C + 2(D-1)(n_inv+n_fwd) + n_op] F, n_op = 0 / 2C
    n_inv, n_fwd = {"KolmogorovFlowVorticity": (4, 1), "NavierStokesVelocity": (6, 3),
                    "KuramotoSivashinsky": (2, 1)}[w["stepper"]]
    per_stage = 2 * C + 2 * (D - 1) * (n_inv + n_fwd)
    step = (2 * per_stage + 2 * C) * F          # c3: 26 F, c4: 90 F
    # physical carry (the step's own ifft + fft, +4 F per channel) once per SAVED step: sub-steps of a
    # RepeatedStepper carry the state in Fourier space (exponax/_repeated_stepper.py:56-102)
    carry = 0 if spectral_carry else 4 * C * F / w.get("substeps", 1)
    return B * T * (step + carry)


def synth_ic(w, B, seed0=0):
    """Synthetic `ex.ic`-style initial conditions (NumPy restatement, PCG64 seeds)."""
    from oracle import exponax_np as ox  # input generator only (harness), never on the timed path
    N, D, C = w["N"], w["D"], w["C"]
    rng = np.random.default_rng(seed0)
    if D == 1:
        # truncated Fourier series, cutoff 5, max |u| = 1 (vectorised over the batch)
        k = np.arange(1, 6)
        x = np.arange(N) * (2 * np.pi / N)
        a = rng.standard_normal((B, C, 5, 1)).astype(np.float32)
        b = rng.standard_normal((B, C, 5, 1)).astype(np.float32)
        u = (a * np.cos(k[:, None] * x) + b * np.sin(k[:, None] * x)).sum(axis=2)
        u /= np.abs(u).max(axis=-1, keepdims=True)
        return u.astype(np.float32)
    if D == 2:
        if w["stepper"] == "KuramotoSivashinsky":
            return np.stack([ox.random_truncated_fourier_series(2, N, cutoff=5, seed=seed0 + i, max_one=True)
                             for i in range(B)]).astype(np.float32)
        base = np.stack([ox.gaussian_random_field(2, N, powerlaw_exponent=3.5, seed=seed0 + i) for i in range(8)])
        reps = (B + 7) // 8
        amp = (1.0 + 0.01 * (np.arange(reps * 8, dtype=np.float32) % 16))[:B]
        return (np.tile(base, (reps, 1, 1, 1))[:B] * amp[:, None, None, None]).astype(np.float32)
    g = ox.make_grid(3, w["L"], N)
    tg = np.stack([np.sin(g[0]) * np.cos(g[1]) * np.cos(g[2]), -np.cos(g[0]) * np.sin(g[1]) * np.cos(g[2]),
                   np.zeros_like(g[0])]).astype(np.float32)
    amp = 1.0 + 0.01 * (np.arange(B, dtype=np.float32) % 16)
    return (tg[None] * amp[:, None, None, None, None]).astype(np.float32)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or \
               [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]),
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU reference
_CPU_SAMPLE = {  # bounded sample of each workload for the host-core arm: (trajectories, ETDRK steps per call)
    "c1": (1, 500), "c2": (1024, 50), "c3": (4, 5), "c4": (1, 1), "readme": (50, 20), "c5": (1, 2)}
_WORKER = {}


def _cpu_worker_init(w, fft_workers):
    from oracle import exponax_np as ox
    ox.set_fft_workers(fft_workers)
    _WORKER["ox"] = ox
    _WORKER["st"] = getattr(ox, w["stepper"])(w["D"], w["L"], w["N"], w["dt"], **w["kw"])
    _WORKER["w"] = w


def _cpu_worker_run(job):
    """One chunk of trajectories through the oracle port: `rollout` (every step stored) or `repeat`."""
    u0, T, final_only = job
    ox, st, w = _WORKER["ox"], _WORKER["st"], _WORKER["w"]
    batched = w["C"] == 1 and w["stepper"] in ("Burgers", "KolmogorovFlowVorticity", "KuramotoSivashinskyConservative")
    if batched:  # oracle classes broadcast over a batch axis placed between channel and space for C == 1
        x = np.ascontiguousarray(np.moveaxis(u0, 0, 1))
        if final_only:
            return float(np.abs(ox.repeat(st.step, T)(x)).sum())
        trj = np.empty((T,) + x.shape, dtype=x.dtype)       # the trajectory IS written (ex.rollout semantics)
        for t in range(T):
            x = st.step(x)
            trj[t] = x
        return float(np.abs(trj[-1]).sum())
    acc = 0.0
    for u in u0:
        r = ox.repeat(st.step, T)(u) if final_only else ox.rollout(st.step, T)(u)
        acc += float(np.abs(r[-1] if not final_only else r).sum())
    return acc


def cpu_reference_run(w, steps, warmup, budget_s=20.0):
    """Times the oracle port of the reference on ALL host cores: a bounded sample of the workload, the trajectories
    split over min(cores, sample batch) worker processes, scipy.fft threads filling the remaining cores."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    N, D = w["N"], w["D"]
    Bs, Ts = _CPU_SAMPLE[w["name"]]
    if w["name"] == "c5":
        # the reference cannot construct 2048^3: its CPU arm is sampled at 128^3
        N = 128
        w = dict(w, N=N)
    nproc = max(1, min(cores, Bs))
    fft_workers = max(1, cores // nproc)
    final_only = bool(w["final_only"])
    u0 = synth_ic(w, Bs)
    bounds = np.linspace(0, Bs, nproc + 1).astype(int)
    jobs = [(u0[bounds[i]:bounds[i + 1]], Ts, final_only) for i in range(nproc) if bounds[i + 1] > bounds[i]]
    wl = {k: w[k] for k in ("stepper", "D", "L", "N", "dt", "kw", "C")}
    times = []
    with ProcessPoolExecutor(nproc, mp_context=mp.get_context("fork"), initializer=_cpu_worker_init,
                             initargs=(wl, fft_workers)) as pool:
        list(pool.map(_cpu_worker_run, [(j[0][:1], 1, True) for j in jobs]))   # constructors + first touch, untimed
        for _ in range(max(0, min(warmup, 1))):
            list(pool.map(_cpu_worker_run, jobs))
        t_start = time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            chk = list(pool.map(_cpu_worker_run, jobs))
            times.append(time.perf_counter() - t0)
            assert all(np.isfinite(c) for c in chk)
            if time.perf_counter() - t_start > budget_s and len(times) >= 1:
                break
    units = (N**D) * Bs * Ts
    ms = 1e3 * sum(times) / len(times)
    call = "repeat (final state only)" if final_only else "rollout (every step stored)"
    return dict(value=units / (ms * 1e-3), unit="grid-point*steps/s", cores=cores, kind="port",
                sample=f"{w['stepper']} N={N}^{D}, batch {Bs} x {Ts} steps, {call}; {len(jobs)} worker processes x "
                       f"{fft_workers} scipy.fft threads on {cores} cores; {len(times)} timed passes",
                sample_batch=Bs, sample_steps=Ts, sample_N=N, ms_per_step=ms, steps=len(times))


# ------------------------------------------------------------------------------ cuFFT comparator
_CUFFT_SAMPLE_T = {"c2": 100, "c3": 20, "c4": 4, "readme": 50, "c1": 500}


def cufft_standin_run(w, stepper, u0, reps=3):
    """The same ETDRK2 workload through torch.fft (cuFFT) + unfused eager elementwise kernels on the SAME GPU
    (the stand-in for exponax on JAX-GPU, which is not installable here).  Same batch; the
    number of steps per call is bounded (c2: 100 of the 1000) -- every step costs the same."""
    import torch
    from baseline.cufft_standin import CufftEtdrk2
    N, D, B = w["N"], w["D"], u0.shape[0]
    cf = CufftEtdrk2(stepper)
    sub = w.get("substeps", 1)
    T = min(w["T"], _CUFFT_SAMPLE_T.get(w["name"], w["T"]))
    T -= T % sub
    out = None if w["final_only"] else torch.empty((B, T) + tuple(u0.shape[1:]), dtype=u0.dtype, device="cuda")

    def call():
        if w["final_only"]:
            return cf.repeat(u0, T // sub, substeps=sub)
        return cf.rollout(u0, T, out=out)

    r = call()
    torch.cuda.synchronize()
    assert torch.isfinite(r).all()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del out, r, cf
    torch.cuda.empty_cache()
    return dict(value=(N**D) * B * T / (best * 1e-3), unit="grid-point*steps/s", ms_per_call=best,
                kind="torch.fft.rfftn/irfftn (cuFFT) + unfused eager elementwise kernels, same GPU, same batch",
                sample=f"batch {B} x {T} ETDRK2 steps ({'repeat' if w['final_only'] else 'rollout, every step stored'}), "
                       f"best of {reps}")


# ------------------------------------------------------------------------------ main
def _load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def measure_exb(args, w, rank, world, local_rank, *, steps, warmup, with_e2e, with_clocks):
    """Device-timed (and optionally end-to-end) run of one batched workload on this rank's GPU."""
    import torch
    import torch.distributed as dist

    import exponax_b200 as ex

    N, D, C, B, T = w["N"], w["D"], w["C"], w["B"], w["T"]
    stepper = getattr(ex.stepper, w["stepper"])(D, w["L"], N, w["dt"], **w["kw"])
    u0_host = synth_ic(w, B, seed0=1000 * rank)
    u0 = torch.as_tensor(u0_host, device="cuda")
    sub = w.get("substeps", 1)
    if w["final_only"] and sub > 1:
        fn = ex.vmap(ex.repeat(ex.RepeatedStepper(stepper, sub), T // sub, spectral_carry=args.spectral_carry))
    elif w["final_only"]:
        fn = ex.vmap(ex.repeat(stepper, T, spectral_carry=args.spectral_carry))
    else:
        fn = ex.vmap(ex.rollout(stepper, T, spectral_carry=args.spectral_carry, cuda_graph=args.cuda_graph))
    plan = stepper._plan()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(warmup):
        out = fn(u0)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0 and with_clocks:
        sampler.start()
        time.sleep(0.3)
    launches0 = plan.launch_count()
    evs = []
    barrier()
    t_wall0 = time.time()
    for _ in range(steps):
        flush.zero_()  # evict L2 between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(u0)
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall1 = time.time()
    launches = plan.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop(t_wall0, t_wall1) if (rank == 0 and with_clocks) else None
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    units_per_call = (N**D) * B * T
    value = units_per_call * world * steps / (total_ms * 1e-3)
    ms_per_step = total_ms / steps
    assert torch.isfinite(out).all(), "non-finite result"

    # ---- roofline of the dominant kernel (per launch == per call for the 1-D persistent kernel;
    #      for N-D the pass kernels of one call are taken together) ----
    peaks = _load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    abytes = algorithmic_bytes_per_call(w, spectral_carry=args.spectral_carry)
    achieved = abytes / (ms_per_step * 1e-3) / 1e9
    traffic, ncu_ctx = None, {}
    try:  # DRAM bytes per launch/call measured once with `ncu --set full` (profiles/traffic.json)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[w["name"]]
        traffic = tr["bytes_per_trajectory_step"] * B * T
        ncu_ctx = {k: v for k, v in tr.items() if k.startswith("ncu_")}
    except Exception:
        ncu_ctx = {}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json (measured copy bandwidth)" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_call": abytes, "kernel_launches_per_call": launches / steps}
    if ncu_ctx:
        roofline["ncu"] = ncu_ctx  # what actually bounds the kernel (one `ncu --set full` capture, see profiles/)
    if D == 1:
        # The persistent 1-D kernel keeps the state on chip: HBM sees only the snapshots, so the binding resource is
        # FP32 issue / shared-memory bandwidth (ncu: profiles/*_full_c2.txt).  Report the fractions of the MEASURED
        # on-chip peaks (exb_peak_fp32 / exb_peak_smem micro-benchmarks, SURVEY 8d "measure it") and name the bound.
        import ctypes as Ct
        from exponax_b200 import _native as nat
        nfft = 8 if not args.spectral_carry else 7  # real N-point transforms per ETDRK2 rollout step (Burgers)
        flops = units_per_call / N * nfft * 2.5 * N * np.log2(N)
        tfl = flops / (ms_per_step * 1e-3) / 1e12
        roofline["hbm"] = {"achieved": achieved, "peak": peak, "frac": achieved / peak, "unit": "GB/s"}
        try:
            f2, s2 = (Ct.c_double * 2)(), (Ct.c_double * 2)()
            nat.check(nat.lib().exb_peak_fp32(None, f2))
            nat.check(nat.lib().exb_peak_smem(None, s2))
            # shared-memory bytes per trajectory step (fast kernel, Burgers ETDRK2): 8 complex-line transforms'
            # exchanges are shared by a PAIR of trajectories; per N-point complex transform one exchange
            # (8 B store + 8 B load per point) + twiddles (8 B per point) + state / stage / table traffic
            smem_bytes_per_pair_step = (4 * 24 + 96) * N
            if "ncu_shared_wavefronts_per_pair_step" in ncu_ctx:   # measured: 128 B per shared-memory wavefront
                smem_bytes_per_pair_step = 128 * ncu_ctx["ncu_shared_wavefronts_per_pair_step"]
            smem_gbs = (units_per_call / N / 2) * smem_bytes_per_pair_step / (ms_per_step * 1e-3) / 1e9
            roofline["fp32"] = {"achieved_tflops_fft_model": tfl, "peak_tflops_ffma": f2[0], "peak_tflops_ffma2": f2[1],
                                "frac_of_ffma_peak": tfl / f2[0],
                                "note": "2.5 N log2 N flop per real transform (adds and multiplies, not FMAs: an "
                                        "FFT butterfly can fuse at most ~1/3 of its flops, so ~0.5 of the FFMA peak "
                                        "is the ceiling of this instruction mix)"}
            roofline["smem"] = {"achieved_gbs_model": smem_gbs, "peak_gbs_lds64": s2[0], "peak_gbs_lds128": s2[1],
                                "frac_of_lds64_peak": smem_gbs / s2[0]}
            if "ncu_issue_slot_utilisation" in ncu_ctx:
                roofline["bound"] = ("shared-memory pipe / fp32 pipe (ncu: LSU pipe %.0f %%, FMA pipe %.0f %%, issue slots %.0f %%); "
                                     "HBM is %.0f %% used") % (
                    100 * ncu_ctx.get("ncu_shared_wavefront_utilisation", 0), 100 * ncu_ctx.get("ncu_fma_pipe_utilisation", 0),
                    100 * ncu_ctx["ncu_issue_slot_utilisation"], 100 * achieved / peak)
        except Exception as e:  # measurement utility only
            roofline["fp32"] = {"error": str(e)}
        roofline["fp32_fft_tflops"] = tfl
        roofline["note"] = ("1-D persistent kernel: state resident in shared memory, HBM sees only snapshots; "
                            "`achieved`/`frac` are the HBM figures the contract asks for, the binding resources are "
                            "under `fp32` / `smem` (SURVEY 8d)")

    # ---- the library-FFT comparator on the same GPU (rank 0, N = 1) ----
    lib = None
    if world == 1 and not args.no_cufft:
        try:
            lib = cufft_standin_run(w, stepper, u0)
            lib["speedup_of_this_repo"] = value / lib["value"]
        except Exception as e:
            lib = {"error": repr(e)[:300]}

    # ---- e2e through the public API with host buffers ----
    e2e = None
    if with_e2e:
        pinned_in = torch.from_numpy(u0_host).pin_memory()
        res_shape = tuple(out.shape)
        pinned_out = torch.empty(res_shape, dtype=out.dtype, pin_memory=True)
        nchunk = 8 if B >= 64 else 1
        bounds = np.linspace(0, B, nchunk + 1).astype(int)
        # two streams: the D2H copy of chunk i overlaps the compute of chunk i+1 (scratch is per stream)
        streams = [torch.cuda.Stream() for _ in range(2)]

        def e2e_call():
            # chunked over the batch so the D2H of chunk i overlaps the compute of chunk i+1
            for i in range(nchunk):
                s = streams[i % len(streams)]
                lo, hi = int(bounds[i]), int(bounds[i + 1])
                with torch.cuda.stream(s):
                    d = pinned_in[lo:hi].to("cuda", non_blocking=True)
                    r = fn(d)
                    pinned_out[lo:hi].copy_(r, non_blocking=True)
            for s in streams:
                s.synchronize()

        e2e_call()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(steps, 3))
        for _ in range(n_e2e):
            e2e_call()
        barrier()
        dt_e2e = time.perf_counter() - t0
        te = torch.tensor([dt_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": units_per_call * world * n_e2e / float(te.item()), "unit": "grid-point*steps/s",
               "h2d_bytes_per_step": int(u0_host.nbytes), "d2h_bytes_per_step": int(pinned_out.numel() * 4),
               "calls": n_e2e, "note": "pinned host buffers, 8 batch chunks on 2 streams (copy/compute overlap)"}
        del pinned_in, pinned_out
    del out, u0, flush
    torch.cuda.empty_cache()
    return dict(value=value, ms_per_step=ms_per_step, roofline=roofline, e2e=e2e, clocks=clocks,
                launches=int(launches), gpu_library_baseline=lib)


def workload_config(name, w, args, world):
    strong = args.scaling == "strong"
    return {"workload": f"{name}: {w['desc']}", "stepper": w["stepper"], "N": w["N"], "D": w["D"],
            "batch_per_gpu": w["B"], "etdrk_steps_per_call": w["T"], "order": 2, "precision": "f32",
            "parallelism": f"batch-shard x{world} (no collective)" + (
                f", STRONG scaling: total batch {w['B'] * world} divided over the ranks" if strong else ""),
            "save": "final state only (ex.repeat)" if w["final_only"] else "every step (ex.rollout)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c2 (the headline line) plus short c3 / c4 runs under `also`")
    ap.add_argument("--impl", default="exb", choices=["exb", "reference", "cufft"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: per-GPU batch fixed (default); strong: BASELINE's total batch divided over the ranks")
    ap.add_argument("--batch", type=int, default=None, help="override per-GPU batch")
    ap.add_argument("--T", type=int, default=None, help="override ETDRK steps per call")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cufft", action="store_true", help="skip the torch.fft (cuFFT) comparator")
    ap.add_argument("--no-also", action="store_true", help="default run: skip the short c3 / c4 runs")
    ap.add_argument("--spectral-carry", action="store_true")
    ap.add_argument("--N", type=int, default=None, help="override the grid size (c5)")
    ap.add_argument("--cuda-graph", action="store_true", help="replay the fused call from a captured CUDA graph")
    ap.add_argument("--no-raw-exchange", action="store_true", help="c5: pack / unpack copies around the all-to-all")
    ap.add_argument("--no-peer-stores", action="store_true",
                    help="c5: NCCL all-to-all transposes instead of pass kernels storing into peer memory")
    ap.add_argument("--peer-stores", action="store_true",
                    help="c5: pass kernels store into peer memory over NVLink instead of the NCCL all-to-all (opt-in)")
    ap.add_argument("--exchange", default=None, choices=["nccl", "ce"],
                    help="c5: transposes through NCCL all-to-all or as copy-engine peer copies (symmetric memory)")
    ap.add_argument("--no-overlap", action="store_true", help="c5: do not pipeline transposes against passes")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    default_run = args.workload is None
    name = args.workload or "c2"
    w = dict(WORKLOADS[name], name=name)
    if args.batch:
        w["B"] = args.batch
    if args.T:
        w["T"] = args.T
    if args.N:
        w["N"] = args.N
    if args.scaling == "strong" and name != "c5":
        if w["B"] % world:
            raise SystemExit(f"--scaling strong: batch {w['B']} is not divisible by {world} ranks")
        w["B"] //= world
    if name == "c5" and args.impl == "exb" and world < 2:
        raise SystemExit("workload c5 shards ONE field over the ranks: launch it with torchrun on >= 2 GPUs")
    if name == "c5" and args.impl == "exb":
        return run_c5(args, w, rank, world, local_rank)
    args.warmup = max(args.warmup, 3) if args.impl == "exb" else args.warmup
    config = workload_config(name, w, args, world)

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(w, args.steps, args.warmup, budget_s=120.0)
        config = dict(config, sample={"what": "bounded sample of this workload timed on the host cores",
                                      "batch": r["sample_batch"], "etdrk_steps_per_call": r["sample_steps"],
                                      "N": r["sample_N"],
                                      "call": "ex.repeat" if w["final_only"] else "ex.rollout (trajectory written)"})
        line = {"impl": "reference", "metric": "ETDRK grid-point*steps/s", "value": r["value"], "unit": r["unit"],
                "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "NumPy/SciPy port of exponax (oracle/) on all host cores; JAX cannot be installed here"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)

    if args.impl == "cufft":
        # the library-FFT comparator as its own arm (one GPU): same line format, "impl": "cufft"
        if rank != 0:
            return
        import exponax_b200 as ex
        stepper = getattr(ex.stepper, w["stepper"])(w["D"], w["L"], w["N"], w["dt"], **w["kw"])
        u0 = torch.as_tensor(synth_ic(w, w["B"]), device="cuda")
        r = cufft_standin_run(w, stepper, u0, reps=max(1, args.steps))
        line = {"impl": "cufft", "metric": "ETDRK grid-point*steps/s", "value": r["value"], "unit": r["unit"],
                "n_gpus": 1, "steps": max(1, args.steps), "warmup": 1, "ms_per_step": r["ms_per_call"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(config, sample=r["sample"]), "gpu_library_baseline": r,
                "note": "torch.fft (cuFFT) + eager elementwise restatement of the reference's step: the kernel "
                        "structure XLA emits for exponax on a GPU; none of this repo's kernels run here"}
        print(json.dumps(line))
        return

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    m = measure_exb(args, w, rank, world, local_rank, steps=args.steps, warmup=args.warmup,
                    with_e2e=not args.no_e2e, with_clocks=True)

    # ---- default run: the other two batched BASELINE configs, short, so the driver verifies them every round ----
    also = {}
    if default_run and not args.no_also:
        for other in ("c3", "c4"):
            wo = dict(WORKLOADS[other], name=other)
            if args.scaling == "strong":
                if wo["B"] % world:
                    continue
                wo["B"] //= world
            try:
                mo = measure_exb(args, wo, rank, world, local_rank, steps=3, warmup=3, with_e2e=False,
                                 with_clocks=False)
                also[other] = {"value": mo["value"], "unit": "grid-point*steps/s", "ms_per_step": mo["ms_per_step"],
                               "steps": 3, "warmup": 3, "config": workload_config(other, wo, args, world),
                               "roofline": {k: mo["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac",
                                                                           "traffic", "algorithmic_bytes_per_call")},
                               "gpu_launches": mo["launches"], "gpu_library_baseline": mo["gpu_library_baseline"]}
            except Exception as e:  # never lose the headline line to a secondary workload
                also[other] = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu and world == 1:
        r = cpu_reference_run(w, steps=3, warmup=1, budget_s=20.0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": "ETDRK grid-point*steps/s", "value": m["value"], "unit": "grid-point*steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, l2="256 MB flush between timed iterations; outputs >> L2"),
            "roofline": m["roofline"], "cpu_baseline": cpu, "clocks": m["clocks"], "e2e": m["e2e"],
            "gpu_launches": m["launches"], "trajectory_steps_per_s": m["value"] / (w["N"] ** w["D"]),
            "gpu_library_baseline": m["gpu_library_baseline"]}
    if also:
        line["also"] = also
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
