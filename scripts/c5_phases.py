"""Diagnostic (torchrun, >= 2 GPUs): one slab-decomposed ETDRK2 step of config c5 split into its phases.

The phases of `SlabStepper.step_fourier` (NCCL path) are run serially on one stream with CUDA events around each:
axis-0 inverse pass + prologue, the 6 inverse transposes, axis-1 inverse passes, row pass, axis-1 forward passes, the
3 forward transposes, axis-0 forward pass + ETDRK epilogue.  Prints the max over ranks per phase (ms per step) next to
the pipelined full step, and the all-to-all bandwidth seen by one rank.

usage: torchrun --nproc-per-node 8 scripts/c5_phases.py [N]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import exponax_b200 as ex  # noqa: E402
from exponax_b200 import _native as nat  # noqa: E402
from exponax_b200 import _slab  # noqa: E402
from exponax_b200.csrc_meta import etdrk_stage_input  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
_slab.init_process_group_nccl(torch.cuda.current_device())
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
L = 2 * np.pi
slab = ex.SlabStepper.navier_stokes_velocity(L, N, 1e-3, injection_mode=4, diffusivity=0.01)
slab.plan()
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


def full_step(reps):
    global uh
    for _ in range(2):
        uh = slab.step_fourier(uh, inplace=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        uh = slab.step_fourier(uh, inplace=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {"N": N, "world": world, "Kp": None}
slab.overlap = True
res["full_step_pipelined_ms"] = full_step(3)
slab.overlap = False
if not os.environ.get("C5_SKIP_SERIAL"):
    res["full_step_serial_ms"] = full_step(2)
res["Kp"] = slab.Kp

# ---- the serial path again, phase by phase (mirrors SlabStepper.step_fourier, NCCL branch, overlap off) ----
S = [slab._buf(f"S{i}", slab.Cn).view(slab.Cn, N, slab.n, slab.Nh) for i in range(slab.order)] + [None] * (4 - slab.order)
nb = max(slab.n_inv, slab.n_fwd)
wb = slab._buf("w_b", nb, fields=True)
winv_b = wb[:slab.n_inv].view(slab.n_inv, N, slab.n, slab.Kp)
wfwd_b = wb[:slab.n_fwd].view(slab.n_fwd, N, slab.n, slab.Kp)
winv_a = slab._buf("winv_a", slab.n_inv, fields=True)
wfwd_a = slab._buf("wfwd_a", slab.n_fwd, fields=True)
seg = nat.SLAB_SEGMENTED
kept = slab.kept
names = ["col0_inv_prologue", "transposes_inv(%d)" % slab.n_inv, "col1_inv", "row_nl", "col1_fwd",
         "transposes_fwd(%d)" % slab.n_fwd, "col0_fwd_epilogue"]
acc = {k: 0.0 for k in names}
reps = 2
for rep in range(reps + 1):
    for s in range(slab.order):
        si = etdrk_stage_input(slab.order, s)
        src = uh if si < 0 else S[si]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        slab._pass(nat.SLAB_COL0_INV_PRO, slab.n_inv, src, winv_b)
        ev[1].record()
        for f in range(slab.n_inv):
            _slab.exchange_b_to_a_raw(winv_b[f], winv_a[f], slab.group, kept)
        ev[2].record()
        slab._pass(nat.SLAB_COL1_INV_NL | seg, slab.n_inv, winv_a, winv_a)
        ev[3].record()
        slab._pass(nat.SLAB_ROW_NL, slab.n_inv, winv_a, wfwd_a)
        ev[4].record()
        slab._pass(nat.SLAB_COL1_FWD_NL | seg, slab.n_fwd, wfwd_a, wfwd_a)
        ev[5].record()
        for g in range(slab.n_fwd):
            _slab.exchange_a_to_b_raw(wfwd_a[g], wfwd_b[g], slab.group, kept)
        ev[6].record()
        slab._pass(nat.SLAB_COL0_FWD_EPI, slab.n_fwd, wfwd_b, None, stage=s, U=uh, OUT=uh, S=S)
        ev[7].record()
        torch.cuda.synchronize()
        if rep > 0:
            for i, k in enumerate(names):
                acc[k] += ev[i].elapsed_time(ev[i + 1]) / reps
t = torch.tensor([acc[k] for k in names], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
res["phases_ms_per_step_max_over_ranks"] = {k: round(float(v), 2) for k, v in zip(names, t.tolist())}
res["phases_sum_ms"] = round(float(t.sum()), 2)
# bytes one kept rank sends per transposed field: its slab minus the block it keeps, kept destinations only
nkept = sum(kept) if kept is not None else world
field_bytes = N * slab.n * slab.Kp * 8
sent = field_bytes * (nkept - 1) / world
tr_ms = (acc[names[1]] + acc[names[5]]) / (2 * (slab.n_inv + slab.n_fwd))      # per field transpose (this rank)
res["transpose_per_field_ms"] = round(tr_ms, 3)
res["all_to_all_send_GBs_per_rank"] = round(sent / (tr_ms * 1e-3) / 1e9, 1)
res["kept_ranks"] = kept
res["nccl_env"] = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
