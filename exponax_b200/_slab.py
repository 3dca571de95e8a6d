"""
Slab-decomposed 3-D ETDRK stepping of ONE field that is sharded over several GPUs
(BASELINE config c5).  One process per GPU (`torchrun`):

  physical slab  A: (F, N/P, N, N)        x-planes   [rank*N/P, (rank+1)*N/P)
  spectral slab  B: (F, N, N/P, N/2+1)    k1-indices [rank*N/P, (rank+1)*N/P)

A 3-D transform = local last-axis pass + local axis-1 pass on layout A, ONE all-to-all transpose
(NCCL over NVLink, `torch.distributed.all_to_all_single`), local axis-0 pass on layout B.  All
nonlinear / ETDRK arithmetic is per mode or per point and therefore local; the passes are the same
fused sm_100a kernels as on one GPU, driven through `exb_slab_pass` (include/exb.h).

The reference has no multi-device code at all: it cannot even construct the dense
operator arrays of a 2048^3 problem on one device.  Here every rank only ever holds its
own slab of the state, of the intermediates and of the ETDRK coefficient tables.
"""
from __future__ import annotations

import os

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _array as A
from . import _native as nat
from ._base_stepper import BaseStepper
from .csrc_meta import etdrk_stage_input


def _group_info(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def init_process_group_nccl(local_rank: int):
    """NCCL process group whose internal stream has high priority, so that the all-to-all kernels are
    scheduled ahead of the grid-filling pass kernels they overlap with."""
    opts = None
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    except Exception:
        pass
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)


def transpose_a_to_b(x: torch.Tensor, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """(F, n, N, K) split over x-planes  ->  (F, N, n, K) split over axis-1 indices (n = N/P).
    Field by field (bounded temporaries): one pack copy + one all-to-all, received in place."""
    rank, P = _group_info(group)
    F, n, N, K = x.shape
    if P == 1:
        if out is None:
            return x
        out.view(-1).copy_(x.reshape(-1))
        return out
    if out is None:
        out = torch.empty((F, N, n, K), dtype=x.dtype, device=x.device)
    from ._spectral import SLAB_CYCLIC
    for f in range(F):
        if SLAB_CYCLIC:   # axis-1 index i lives on rank i % P at local position i // P
            send = x[f].view(n, n, P, K).permute(2, 0, 1, 3).contiguous()  # [dest][x][k1 local][K]
        else:
            send = x[f].view(n, P, n, K).permute(1, 0, 2, 3).contiguous()  # [dest][x][k1][K]
        dist.all_to_all_single(out[f].view(P, n, n, K), send, group=group)  # [src][x_src][k1][K]: x_global = src*n + x
    return out


def transpose_b_to_a(x: torch.Tensor, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """(F, N, n, K) split over axis-1 indices  ->  (F, n, N, K) split over x-planes.
    Field by field: the send side is already contiguous per destination, one unpack copy."""
    rank, P = _group_info(group)
    F, N, n, K = x.shape
    if P == 1:
        if out is None:
            return x
        out.view(-1).copy_(x.reshape(-1))
        return out
    if out is None:
        out = torch.empty((F, n, N, K), dtype=x.dtype, device=x.device)
    from ._spectral import SLAB_CYCLIC
    recv = torch.empty((P, n, n, K), dtype=x.dtype, device=x.device)
    for f in range(F):
        dist.all_to_all_single(recv, x[f].view(P, n, n, K), group=group)   # [src][x][k1_src][K]
        if SLAB_CYCLIC:
            out[f].view(n, n, P, K).copy_(recv.permute(1, 2, 0, 3))         # k1_global = k1 * P + src
        else:
            out[f].view(n, P, n, K).copy_(recv.permute(1, 0, 2, 3))         # k1_global = src*n + k1
    return out


def kept_ranks(N: int, P: int, kmax: int):
    """ranks whose axis-1 index range [r N/P, (r+1) N/P) holds at least one wavenumber |k1| <= kmax (all when
    kmax < 0).  With the 2/3 rule on 8 ranks, ranks 3 and 4 own nothing but dealiased modes."""
    n = N // P
    if kmax < 0:
        return [True] * P
    i = np.arange(N)
    k = np.where(i < (N + 1) // 2, i, i - N)                 # fftfreq ordering (exponax/_spectral.py:40-41)
    return [bool((np.abs(k[r * n:(r + 1) * n]) <= kmax).any()) for r in range(P)]


def exchange_b_to_a_raw(x_b: torch.Tensor, out_a: torch.Tensor, group=None, kept=None):
    """All-to-all of ONE field from the spectral slab (N, n, K) straight into `out_a`, which then holds the
    raw peer-major layout [peer][x][k1 within peer][K] (EXB_SLAB_SEGMENTED): no pack, no unpack.
    `kept[r]` False: rank r's axis-1 range is entirely outside the dealiasing mask -- it neither sends nor is its
    block received (the consumers read masked line entries as zeros without loading them)."""
    rank, P = _group_info(group)
    N, n, K = x_b.shape
    if kept is None or all(kept):
        dist.all_to_all_single(out_a.view(P, n, n, K), x_b.view(P, n, n, K), group=group)
        return
    src, dst = x_b.view(P, n, n, K), out_a.view(P, n, n, K)
    empty = x_b.new_empty(0)
    send = [src[p] if kept[rank] else empty for p in range(P)]
    recv = [dst[p] if kept[p] else empty for p in range(P)]
    dist.all_to_all(recv, send, group=group)


def exchange_a_to_b_raw(x_a: torch.Tensor, out_b: torch.Tensor, group=None, kept=None):
    """Inverse of `exchange_b_to_a_raw`: raw peer-major A buffer -> spectral slab (N, n, K); blocks destined to
    ranks whose axis-1 range is dealiased away are not sent."""
    rank, P = _group_info(group)
    N, n, K = out_b.shape
    if kept is None or all(kept):
        dist.all_to_all_single(out_b.view(P, n, n, K), x_a.view(P, n, n, K), group=group)
        return
    src, dst = x_a.view(P, n, n, K), out_b.view(P, n, n, K)
    empty = x_a.new_empty(0)
    send = [src[p] if kept[p] else empty for p in range(P)]
    recv = [dst[p] if kept[rank] else empty for p in range(P)]
    dist.all_to_all(recv, send, group=group)


class _ShapeOnly:
    """stands in for a freed operator array where only its shape is consulted"""

    def __init__(self, shape):
        self.shape, self.ndim = shape, len(shape)


class _LeanNonlinearFun:
    """Descriptor-only nonlinear function (no operator / mask arrays are ever materialised)."""

    def __init__(self, desc: dict, kmax: int):
        self._desc, self._kmax = desc, kmax

    def _native_desc(self, num_channels):
        return self._desc


class _LeanStepper:
    """What `SlabStepper` needs of a stepper, without the dense host arrays of `BaseStepper`."""

    def __init__(self, N, L, dt, num_channels, dtype, integrator, nonlinear_fun, slab):
        self.num_spatial_dims, self.num_points, self.domain_extent, self.dt = 3, N, L, dt
        self.num_channels, self._dtype = num_channels, dtype
        self._integrator, self._nonlinear_fun, self._slab = integrator, nonlinear_fun, slab
        self.dx = L / N

    def _plan_available(self):
        return True


class SlabStepper:
    """Distributed counterpart of a 3-D `BaseStepper` for a single, slab-sharded field."""

    @classmethod
    def navier_stokes_velocity(cls, domain_extent: float, num_points: int, dt: float, *, diffusivity: float = 0.01,
                               drag: float = 0.0, injection_mode: int | None = None, injection_scale: float = 1.0,
                               order: int = 2, dealiasing_fraction: float = 2 / 3, num_circle_points: int = 16,
                               circle_radius: float = 1.0, group=None):
        """Lean constructor for `NavierStokesVelocity` (injection_mode=None) / `KolmogorovFlowVelocity`
        (exponax/stepper/_navier_stokes.py:330-598) on a slab: the linear operator
        `nu * laplace + drag` is assembled ON THE GPU for the local spectral slab from the 1-D
        wavenumber vectors (same float32 arithmetic as `build_derivative_operator` /
        `build_laplace_operator`), the ETDRK tables are built and kept on the device, and the
        nonlinear function is a pure descriptor.  Nothing of size N^3 ever exists on the host --
        the only way to set up the 2048^3 problem of config c5."""
        from . import _spectral as sp
        from . import etdrk
        from ._config import real_dtype

        rank, P = _group_info(group)
        rd = real_dtype()
        N, n = num_points, num_points // max(P, 1)
        td = A.real_t(rd)
        scale = rd(2 * np.pi / domain_extent)
        k_other = np.fft.fftfreq(N, 1 / N).astype(rd)
        k_last = np.fft.rfftfreq(N, 1 / N).astype(rd)
        d0 = torch.as_tensor(scale * k_other, device="cuda", dtype=td).view(N, 1, 1)
        mine = sp.slab_indices(rank, max(P, 1), N) if P > 1 else slice(None)
        d1 = torch.as_tensor(scale * k_other[mine], device="cuda", dtype=td).view(1, n, 1)
        d2 = torch.as_tensor(scale * k_last, device="cuda", dtype=td).view(1, 1, N // 2 + 1)
        lap = -(d0 * d0) - (d1 * d1) - (d2 * d2)          # sum_d (i k_d)^2, real
        lin = (rd(diffusivity) * lap + rd(drag)).to(A.cplx_t(rd)).unsqueeze(0)
        del lap
        cutoff = dealiasing_fraction * (N // 2) - 1         # nonlin_fun/_base.py:59-71
        kmax = sp.dealias_kmax(N, cutoff, rd)
        inj = None
        owned = (injection_mode is not None
                 and (P <= 1 or injection_mode in range(N)[sp.slab_indices(rank, max(P, 1), N)]))
        if injection_mode is not None and owned and 0 < injection_mode < N // 2:
            # (0, +k_f, 0) on channel 0, value gamma * N^3 / 2 (coef_extraction scaling), SURVEY App. B.14
            inj = ((0, int(injection_mode), 0), float(injection_scale) * N * (N / 2) * N)
        nl = _LeanNonlinearFun({"kind": nat.NL_PROJECTED_3D, "injection": inj}, kmax)
        cls_e = {0: etdrk.ETDRK0, 1: etdrk.ETDRK1, 2: etdrk.ETDRK2, 3: etdrk.ETDRK3, 4: etdrk.ETDRK4}[order]
        if order == 0:
            integ = cls_e(dt, lin)
        else:
            integ = cls_e(dt, lin, nl, num_circle_points=num_circle_points, circle_radius=circle_radius)
        integ._L_dev = None  # the operator itself is not needed after the tables exist
        integ._linear_operator = _ShapeOnly(tuple(lin.shape))
        del lin
        torch.cuda.empty_cache()
        return cls(_LeanStepper(N, domain_extent, dt, 3, rd, integ, nl, (rank, max(P, 1))), group=group)

    def __init__(self, stepper, group=None):
        if stepper.num_spatial_dims != 3:
            raise ValueError("SlabStepper needs a 3-D stepper")
        if not stepper._plan_available():
            raise NotImplementedError("SlabStepper needs a stepper with a native nonlinear function")
        self.stepper = stepper
        self.group = group
        self.rank, self.P = _group_info(group)
        N = stepper.num_points
        if N % self.P:
            raise ValueError(f"num_points={N} must be divisible by the number of ranks ({self.P})")
        self.N, self.n, self.Nh, self.Cn = N, N // self.P, N // 2 + 1, stepper.num_channels
        self._plan = None
        self._bufs = {}
        self.overlap = True   # pipeline the all-to-all transposes against the passes (two streams)
        self.raw_exchange = True  # no pack / unpack copies around the all-to-all (EXB_SLAB_SEGMENTED)
        # Transposes fused into the pass kernels' stores over NVLink peer memory (falls back to the all-to-all).
        # A pass CTA owns an [N x TW] tile, so a remote store is TW * 8 contiguous bytes: measured +7 % at
        # 1024^3 on 2 GPUs (TW = 4: 109 vs 117 ms) but -38 % at 2048^3 on 8 GPUs (TW = 2, 16-byte NVLink writes:
        # 539 vs 391 ms) -> on by default only up to N = 1024; EXB_SLAB_PEER_STORES=1 / 0 forces it on / off.
        # Round 2: with the compact field pitch the NCCL all-to-all ships 2/3 of the bytes and wins at every size
        # measured (1024^3 on 2 GPUs: 76.8 ms vs 83.5 ms with peer stores) -> peer stores are opt-in.
        env = os.environ.get("EXB_SLAB_PEER_STORES", "auto")
        self.peer_stores = env == "1"
        self.prune_peers = os.environ.get("EXB_SLAB_PRUNE_PEERS", "1") != "0"
        # Transposes as PEER COPIES of whole blocks into the peers' symmetric-memory buffers (cudaMemcpyAsync on a
        # second stream: the copy engines move the data over NVLink, no SM is taken from the pass kernels that run
        # at the same time -- NCCL's all-to-all is an SM kernel and competes with them) + one barrier per transpose.
        self.ce_exchange = os.environ.get("EXB_SLAB_EXCHANGE", "nccl") == "ce"
        from ._spectral import SLAB_CYCLIC
        self.cyclic = bool(SLAB_CYCLIC) and self.P > 1
        self.Kp, self.kept = self.Nh, None

    # ---- plan with the LOCAL slices of the coefficient tables --------------------------------
    def _local(self, arr):
        if self.stepper._slab is not None:  # built inside ex.spectral.slab_context: already local
            if tuple(self.stepper._slab) != (self.rank, self.P):
                raise ValueError(f"stepper was built for slab {self.stepper._slab}, process group says "
                                 f"{(self.rank, self.P)}")
            return arr if hasattr(arr, "is_cuda") else np.ascontiguousarray(arr)
        from ._spectral import slab_indices
        sl = slab_indices(self.rank, self.P, self.N)
        if hasattr(arr, "is_cuda"):
            return arr[:, :, sl, :].contiguous()
        return np.ascontiguousarray(arr[:, :, sl, :])

    def plan(self):
        if self._plan is None:
            st, it = self.stepper, self.stepper._integrator
            nl = st._nonlinear_fun
            order = it.order
            desc = nl._native_desc(st.num_channels) if order > 0 else {"kind": nat.NL_ZERO}
            half = it._half_exp()
            self._plan = nat.Plan(
                D=3, N=self.N, C_=self.Cn, E=it._linear_operator.shape[0], order=order, dtype=st._dtype,
                L=st.domain_extent, kmax=nl._kmax if order > 0 else -1, nl=desc,
                exp_term=self._local(it._exp_term), half_exp_term=None if half is None else self._local(half),
                coefs=[self._local(c) for c in it._coef_list()], slab=(max(self.P, 1), self.rank),
                slab_cyclic=self.cyclic)
            ni, nf = C.c_int32(), C.c_int32()
            nat.check(nat.lib().exb_plan_nl_fields(self._plan.handle, C.byref(ni), C.byref(nf)))
            self.n_inv, self.n_fwd = ni.value, nf.value
            self.order = order if desc.get("kind") != nat.NL_ZERO else 0
            # last-axis pitch of the exchanged field buffers: only the wavenumbers inside the dealiasing mask when
            # the fast kernels run (exb_plan_field_pitch) -> every transpose ships ~2/3 of the bytes; and the ranks
            # whose axis-1 range is dealiased away take no part in the transposes
            self.Kp = int(nat.lib().exb_plan_field_pitch(self._plan.handle))
            kmax = nl._kmax if order > 0 else -1
            # (block distribution only: under the cyclic one every rank owns kept wavenumbers)
            self.kept = (kept_ranks(self.N, self.P, kmax)
                         if (self.Kp != self.Nh and self.prune_peers and not self.cyclic) else None)
        return self._plan

    def release_buffers(self):
        """Drop the cached work buffers (they are re-created on demand)."""
        self._bufs.clear()
        torch.cuda.empty_cache()

    def _buf(self, name, nfields, real=False, fields=False):
        """fields: an exchanged field buffer (last-axis pitch `Kp`); otherwise a dense spectral / physical slab."""
        key = (name, nfields, real, fields)
        b = self._bufs.get(key)
        if b is None:
            rd = self.stepper._dtype
            if real:
                b = torch.zeros((nfields, self.n, self.N, self.N), dtype=A.real_t(rd), device="cuda")
            else:
                b = torch.zeros((nfields, self.n, self.N, self.Kp if fields else self.Nh), dtype=A.cplx_t(rd),
                                device="cuda")
            self._bufs[key] = b
        return b

    def _pass(self, kind, nfields, inp, out, *, stage=0, U=None, OUT=None, S=None):
        arr = (C.c_void_p * 4)(*[A.ptr(s) if s is not None else None for s in (S or [None] * 4)])
        nat.check(nat.lib().exb_slab_pass(self.plan().handle, A.stream_ptr(), kind, nfields, stage, A.ptr(inp),
                                          A.ptr(out), A.ptr(U), A.ptr(OUT), arr))

    # ---- data movement helpers (tests / examples) ---------------------------------------------
    def scatter(self, u_global):
        """Local physical slab (C, N/P, N, N) of a replicated global field (C, N, N, N)."""
        t, _ = A.to_device(u_global, self.stepper._dtype)
        return t[:, self.rank * self.n:(self.rank + 1) * self.n].contiguous()

    def gather(self, u_local):
        """Replicated global field from the local slabs (not on the timed path)."""
        if self.P == 1:
            return u_local
        parts = [torch.empty_like(u_local) for _ in range(self.P)]
        dist.all_gather(parts, u_local.contiguous(), group=self.group)
        return torch.cat(parts, dim=1)

    # ---- transforms ---------------------------------------------------------------------------
    def fft(self, u_local):
        """physical slab (F, N/P, N, N) -> spectral slab (F, N, N/P, N/2+1)."""
        self.plan()
        F = u_local.shape[0]
        a = self._buf("fft_a", F)
        self._pass(nat.SLAB_ROW_R2C, F, u_local.contiguous(), a)
        self._pass(nat.SLAB_COL1_FWD, F, a, a)
        b = transpose_a_to_b(a, self.group)
        out = torch.empty((F, self.N, self.n, self.Nh), dtype=b.dtype, device="cuda")
        self._pass(nat.SLAB_COL0_FWD, F, b.contiguous(), out)
        return out

    def ifft(self, uh_local):
        """spectral slab -> physical slab (the input is left untouched)."""
        self.plan()
        F = uh_local.shape[0]
        tmp = self._buf("ifft_b", F).view(F, self.N, self.n, self.Nh)
        self._pass(nat.SLAB_COL0_INV, F, uh_local.contiguous(), tmp)
        a = transpose_b_to_a(tmp, self.group).contiguous()
        self._pass(nat.SLAB_COL1_INV, F, a, a)
        out = torch.empty((F, self.n, self.N, self.N), dtype=A.real_t(self.stepper._dtype), device="cuda")
        self._pass(nat.SLAB_ROW_C2R, F, a, out)
        return out

    # ---- time stepping --------------------------------------------------------------------------
    def step_fourier(self, uh, *, inplace: bool = False):
        """One ETDRK step on the local spectral slab (C, N, N/P, N/2+1).  `inplace=True` overwrites
        `uh` (each mode is read before it is written) -- what long runs use to save memory."""
        self.plan()
        uh = uh.contiguous()
        out = uh if inplace else torch.empty_like(uh)
        if self.order == 0:
            self._pass(nat.SLAB_COL0_FWD_EPI, 0, None, None, U=uh, OUT=out)
            return out
        S = [self._buf(f"S{i}", self.Cn).view(self.Cn, self.N, self.n, self.Nh) for i in range(self.order)]
        S += [None] * (4 - self.order)
        peer = self._peer_buffers() if (self.peer_stores and self.raw_exchange and self.P > 1) else None
        if peer is not None:
            # Transposes fused into the stores of the passes that feed them: every rank writes its results
            # straight into the peers' buffers over NVLink (symmetric memory), a barrier replaces the
            # all-to-all.  Hazards: a peer overwrites my winv_a (wfwd_b) only after the barrier that follows
            # my last read of it in stream order, see DESIGN.md section 5.
            lib, h, st = nat.lib(), self.plan().handle, A.stream_ptr()
            wfwd_a = self._buf("wfwd_a", self.n_fwd, fields=True)
            try:
                for s in range(self.order):
                    si = etdrk_stage_input(self.order, s)
                    src = uh if si < 0 else S[si]
                    nat.check(lib.exb_slab_pass_peer(h, st, nat.SLAB_COL0_INV_PRO, 0, self.n_inv, A.ptr(src),
                                                     peer["ptrs_inv"]))
                    peer["hdl_inv"].barrier(channel=0)
                    self._pass(nat.SLAB_COL1_INV_NL | nat.SLAB_SEGMENTED, self.n_inv, peer["winv_a"], peer["winv_a"])
                    self._pass(nat.SLAB_ROW_NL, self.n_inv, peer["winv_a"], wfwd_a)
                    nat.check(lib.exb_slab_pass_peer(h, st, nat.SLAB_COL1_FWD_NL | nat.SLAB_SEGMENTED, 0, self.n_fwd,
                                                     A.ptr(wfwd_a), peer["ptrs_fwd"]))
                    peer["hdl_fwd"].barrier(channel=0)
                    self._pass(nat.SLAB_COL0_FWD_EPI, self.n_fwd, peer["wfwd_b"], None, stage=s, U=uh, OUT=out, S=S)
                return out
            except NotImplementedError:
                # generic kernels (grid size without a fast instantiation): raised by the FIRST call on every
                # rank alike, before anything was enqueued -> use the all-to-all path from now on
                self.peer_stores = False
        ce = self._peer_buffers() if (self.ce_exchange and self.raw_exchange and self.P > 1) else None
        if ce is not None:
            return self._step_fourier_ce(uh, out, S, ce)
        # the B-layout buffer is shared by the inverse fields and (later in the stage) the forward fields
        nb = max(self.n_inv, self.n_fwd)
        wb = self._buf("w_b", nb, fields=True)
        winv_b = wb[:self.n_inv].view(self.n_inv, self.N, self.n, self.Kp)
        wfwd_b = wb[:self.n_fwd].view(self.n_fwd, self.N, self.n, self.Kp)
        winv_a = self._buf("winv_a", self.n_inv, fields=True)
        wfwd_a = self._buf("wfwd_a", self.n_fwd, fields=True)
        overlap = self.overlap and self.P > 1
        # multi-rank: layout A stays in the raw all-to-all (peer-major) order for the whole N(u) evaluation --
        # the axis-1 passes address it with segmented lines, the row pass is order-agnostic
        seg = nat.SLAB_SEGMENTED if (self.raw_exchange and self.P > 1) else 0
        kept = self.kept
        b2a = ((lambda xb, oa, group=None: exchange_b_to_a_raw(xb, oa, group, kept)) if seg
               else (lambda xb, oa, group=None: transpose_b_to_a(xb[None], group, out=oa[None])))
        a2b = ((lambda xa, ob, group=None: exchange_a_to_b_raw(xa, ob, group, kept)) if seg
               else (lambda xa, ob, group=None: transpose_a_to_b(xa[None], group, out=ob[None])))
        for s in range(self.order):
            si = etdrk_stage_input(self.order, s)
            src = uh if si < 0 else S[si]
            if not overlap:
                self._pass(nat.SLAB_COL0_INV_PRO, self.n_inv, src, winv_b)
                for f in range(self.n_inv):
                    b2a(winv_b[f], winv_a[f], self.group)
                self._pass(nat.SLAB_COL1_INV_NL | seg, self.n_inv, winv_a, winv_a)
                self._pass(nat.SLAB_ROW_NL, self.n_inv, winv_a, wfwd_a)
                self._pass(nat.SLAB_COL1_FWD_NL | seg, self.n_fwd, wfwd_a, wfwd_a)
                for g in range(self.n_fwd):
                    a2b(wfwd_a[g], wfwd_b[g], self.group)
            else:
                # field-by-field pipeline on two streams: the all-to-all of field f (comm stream, NVLink)
                # runs while the compute stream works on the prologue of field f+1 / the axis-1 pass of
                # field f-1; events order producer -> transpose -> consumer per field.
                main, comm = torch.cuda.current_stream(), self._comm_stream()
                arrived = []
                for f in range(self.n_inv):
                    nat.check(nat.lib().exb_slab_inv_pro_fields(self.plan().handle, main.cuda_stream, f, 1,
                                                                A.ptr(src), A.ptr(winv_b)))
                    ready = torch.cuda.Event()
                    ready.record(main)
                    with torch.cuda.stream(comm):
                        comm.wait_event(ready)
                        b2a(winv_b[f], winv_a[f], self.group)
                        ev = torch.cuda.Event()
                        ev.record(comm)
                    arrived.append(ev)
                for f in range(self.n_inv):
                    main.wait_event(arrived[f])
                    self._pass(nat.SLAB_COL1_INV_NL | seg, 1, winv_a[f], winv_a[f])
                self._pass(nat.SLAB_ROW_NL, self.n_inv, winv_a, wfwd_a)
                arrived = []
                for g in range(self.n_fwd):
                    self._pass(nat.SLAB_COL1_FWD_NL | seg, 1, wfwd_a[g], wfwd_a[g])
                    ready = torch.cuda.Event()
                    ready.record(main)
                    with torch.cuda.stream(comm):
                        comm.wait_event(ready)
                        a2b(wfwd_a[g], wfwd_b[g], self.group)
                        ev = torch.cuda.Event()
                        ev.record(comm)
                    arrived.append(ev)
                for ev in arrived:
                    main.wait_event(ev)
            self._pass(nat.SLAB_COL0_FWD_EPI, self.n_fwd, wfwd_b, None, stage=s, U=uh, OUT=out, S=S)
            if overlap:
                # the next stage's prologue overwrites w_b: the comm stream must be done reading it (it is:
                # every transpose was awaited above), and must not start before this epilogue finished
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream())
                self._comm_stream().wait_event(done)
        return out

    def _step_fourier_ce(self, uh, out, S, peer):
        """One ETDRK step with copy-engine transposes: per field, the pass that produces it runs on the compute
        stream while the (P - 1) block copies of the previous field travel on the copy stream; a symmetric-memory
        barrier closes each transpose.  Buffers: winv_a / wfwd_b live in symmetric memory (written by the peers),
        winv_b / wfwd_a are local.  Hazards as for the peer-store path: a peer overwrites my
        winv_a (wfwd_b) only after the barrier that follows my last read of it in stream order."""
        lib, h = nat.lib(), self.plan().handle
        main, comm = torch.cuda.current_stream(), self._comm_stream()
        me, P = self.rank, self.P
        wb = self._buf("w_b", self.n_inv, fields=True)
        winv_b = wb.view(self.n_inv, P, self.n, self.n, self.Kp)          # [field][x-block of rank p][x][k1][K]
        wfwd_a = self._buf("wfwd_a", self.n_fwd, fields=True)
        wfwd_a_blk = wfwd_a.view(self.n_fwd, P, self.n, self.n, self.Kp)   # raw order [k1-owner p][x][k1 local][K]
        winv_a, wfwd_b = peer["winv_a"], peer["wfwd_b"]
        for s in range(self.order):
            si = etdrk_stage_input(self.order, s)
            src = uh if si < 0 else S[si]
            for f in range(self.n_inv):
                nat.check(lib.exb_slab_inv_pro_fields(h, main.cuda_stream, f, 1, A.ptr(src), A.ptr(wb)))
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(comm):
                    comm.wait_event(ready)
                    for d in range(P):                                   # my x-block for rank p -> its winv_a[f][me]
                        p = (me + d) % P
                        peer["remote_inv"][p][f, me].copy_(winv_b[f, p], non_blocking=True)
            with torch.cuda.stream(comm):
                peer["hdl_inv"].barrier(channel=0)
                done = torch.cuda.Event()
                done.record(comm)
            main.wait_event(done)
            self._pass(nat.SLAB_COL1_INV_NL | nat.SLAB_SEGMENTED, self.n_inv, winv_a, winv_a)
            self._pass(nat.SLAB_ROW_NL, self.n_inv, winv_a, wfwd_a)
            for g in range(self.n_fwd):
                self._pass(nat.SLAB_COL1_FWD_NL | nat.SLAB_SEGMENTED, 1, wfwd_a[g], wfwd_a[g])
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(comm):
                    comm.wait_event(ready)
                    for d in range(P):                                   # my k1-block of rank p -> its wfwd_b[g][me]
                        p = (me + d) % P
                        peer["remote_fwd"][p][g, me].copy_(wfwd_a_blk[g, p], non_blocking=True)
            with torch.cuda.stream(comm):
                peer["hdl_fwd"].barrier(channel=0)
                done = torch.cuda.Event()
                done.record(comm)
            main.wait_event(done)
            self._pass(nat.SLAB_COL0_FWD_EPI, self.n_fwd, wfwd_b, None, stage=s, U=uh, OUT=out, S=S)
            fin = torch.cuda.Event()
            fin.record(main)
            comm.wait_event(fin)     # the next stage's copies overwrite the peers' buffers only after everyone's barrier
        return out

    def _peer_buffers(self):
        """winv_a (n_inv fields, layout A) and wfwd_b (n_fwd fields, layout B) in symmetric memory + the peers'
        base pointers; None (and peer_stores switched off on EVERY rank) if symmetric memory is unavailable."""
        if getattr(self, "_peer", None) is not None or not (self.peer_stores or self.ce_exchange):
            return getattr(self, "_peer", None)
        import ctypes
        ok, peer = 1, None
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else dist.group.WORLD
            per_field = self.N * self.n * self.Kp
            rt = A.real_t(self.stepper._dtype)

            def make(nfields, shape):
                flat = symm.empty(nfields * per_field * 2, dtype=rt, device=torch.device("cuda", torch.cuda.current_device()))
                hdl = symm.rendezvous(flat, group)
                ptrs = (ctypes.c_void_p * self.P)(*[int(x) for x in hdl.buffer_ptrs])
                return torch.view_as_complex(flat.view(-1, 2)).view((nfields,) + shape), hdl, ptrs

            winv_a, hdl_inv, ptrs_inv = make(self.n_inv, (self.n, self.N, self.Kp))
            wfwd_b, hdl_fwd, ptrs_fwd = make(self.n_fwd, (self.N, self.n, self.Kp))
            peer = dict(winv_a=winv_a, hdl_inv=hdl_inv, ptrs_inv=ptrs_inv, wfwd_b=wfwd_b, hdl_fwd=hdl_fwd,
                        ptrs_fwd=ptrs_fwd)

            def remote(hdl, nfields):
                # every rank's buffer mapped into this process, viewed [field][block of rank q][n][n][Kp]
                out = []
                for r in range(self.P):
                    t = hdl.get_buffer(r, (nfields * per_field * 2,), rt)
                    out.append(torch.view_as_complex(t.view(-1, 2)).view(nfields, self.P, self.n, self.n, self.Kp))
                return out

            peer["remote_inv"] = remote(hdl_inv, self.n_inv)
            peer["remote_fwd"] = remote(hdl_fwd, self.n_fwd)
        except Exception as e:  # noqa: BLE001 -- any failure means: no peer memory here
            ok, self._peer_error = 0, repr(e)
        flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)     # all ranks take the same path
        if int(flag.item()) == 0:
            self.peer_stores, self.ce_exchange, peer = False, False, None
        self._peer = peer
        return peer

    def _comm_stream(self):
        if getattr(self, "_comm", None) is None:
            # high priority: the (small) NCCL / pack kernels must not queue behind the grid-filling pass kernels
            self._comm = torch.cuda.Stream(priority=-1)
        return self._comm

    def step(self, u_local):
        """One step of the physical slab (C, N/P, N, N): fft -> step_fourier -> ifft
        (exponax/_base_stepper.py:201-220, distributed)."""
        return self.ifft(self.step_fourier(self.fft(u_local)))

    def repeat(self, u_local, n: int, *, spectral_carry: bool = True):
        """n steps; with `spectral_carry` (default for the distributed path) the carry stays in
        Fourier space: 2 all-to-alls per step are saved."""
        if not spectral_carry:
            for _ in range(n):
                u_local = self.step(u_local)
            return u_local
        uh = self.fft(u_local)
        for _ in range(n):
            uh = self.step_fourier(uh, inplace=True)
        return self.ifft(uh)

    __call__ = step
