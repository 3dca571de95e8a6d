// Fast 2-D / 3-D pass kernels (f32, N in {128,256,512} and, for the slab-decomposed 3-D path, 1024 / 2048)
// built on the register FFT core exb_fft8.cuh.  Same pass structure and fusion as the generic kernels (exb_kernels_nd.cuh):
//   col_fast<MODE>:  strided axis; a CTA owns a [N x TW] tile; thread (j, w) holds 8 points of
//                    column w in registers, global loads/stores are coalesced across w, the
//                    shared-memory tile is only the exchange buffer between radix passes;
//                    COL_INV_PRO fuses mask / i*k / inverse Laplacian / curl, COL_FWD_EPI fuses
//                    mask / scale / Leray / injection and the ETDRK stage update.
//   row_fast<MODE>:  contiguous last axis; a group of N/8 threads owns a PAIR of rows
//                    (two-for-one), runs the n_inv inverse transforms, the pointwise
//                    nonlinearity in registers and the n_fwd forward transforms.
// The nonlinear function is a compile-time descriptor S (NlS<...>).
// Slab extras: COL_PLAIN addresses segmented lines (the raw all-to-all buffer of a slab transpose), and
// COL_PLAIN / COL_INV_PRO can store their results straight into the peers' buffers (ColParams::peer_out).
// Compile-time switches below are A/B knobs (scripts/build_variant.sh); their defaults are the measured best.
#pragma once
#include "exb_fft8.cuh"
#include "exb_kernels_nd.cuh"

#ifndef EXB_ROW_PREFETCH
#define EXB_ROW_PREFETCH 1
#endif
// COL_INV_PRO, single-channel state: park the 8 loaded modes of a thread in a thread-private slice of
// shared memory across the field loop instead of 16 registers (the kernel is capped at 64 registers for
// two 512-thread CTAs per SM and spilled to local memory otherwise)
// COL_FWD_EPI: prefetch the ETDRK stage operands into L2 ahead of the transforms.  Measured 2-3 % SLOWER
// on c3 / c4 (r01k: the kernels wait on DRAM queues, not on a cold L2), so it is compiled out.
#ifndef EXB_EPI_PREFETCH
#define EXB_EPI_PREFETCH 0
#endif
// COL_FWD_EPI / COL_FWD_NL with several channels: evaluate N(u) of all 8 modes first, then stream the update.
// Measured on c4 (r01l): 9.04e9 vs 9.42e9 interleaved (more spills at 64 registers) -> off.
#ifndef EXB_EPI_TWO_PHASE
#define EXB_EPI_TWO_PHASE 0
#endif
// resident CTAs per SM the multi-field epilogue kernels are compiled for (2: 64 registers with spills,
// 1: 96 registers without).  Measured on c4 (r01l): 1 -> 8.2e9, 2 -> 9.42e9: occupancy wins.
#ifndef EXB_EPI_MULTI_MINBLOCKS
#define EXB_EPI_MULTI_MINBLOCKS 2
#endif
#ifndef EXB_INVPRO_SMEM_U
#define EXB_INVPRO_SMEM_U 1
#endif
#ifndef EXB_EPI_BATCH_FASTEST
#define EXB_EPI_BATCH_FASTEST 1
#endif
// COL_FWD_EPI: dealiased modes take the closed-form update (etdrk_update_masked): no stage buffers, no N(u)
#ifndef EXB_EPI_MASKED_SKIP
#define EXB_EPI_MASKED_SKIP 1
#endif

namespace exb {

// 8-byte asynchronous global -> shared copy (LDGSTS) and its completion primitives
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One-channel 2-D prologue + inverse axis-0 transforms of the fields [f_begin, f_end) for the column owned by this
// thread row: the stage input sits in a thread-private stash `us[q * NT]` (zeros at dealiased entries).  Shared by
// the persistent prologue kernel and by the FUSED epilogue (col_fast_kernel<.., FUSE = 1>).
template <int N, int TW, class S, class Ex>
__device__ __forceinline__ void invpro_fields_1ch(const NlParams<float>& Pn, const cpx<float>* us, int NT, int j, bool col_keep,
                                                  unsigned rowmask, float kd1, int f_begin, int f_end, const Ex& ex,
                                                  const cpx<float>* tw, cpx<float>* dst0, size_t fM, size_t fqstride) {
  constexpr int P = N / 8;
  constexpr bool VORT = S::kind == EXB_NL_VORTICITY_2D;
  const cpx<float> zero(0.f, 0.f);
  auto kd0_of = [&](int q) { return Pn.dscale * (float)(j + P * q - (q >= 4 ? N : 0)); };
  for (int f = f_begin; f < f_end; ++f) {
    cpx<float> v[8];
    if (VORT) {
      // u = +d_y psi, v = -d_x psi, d_x w, d_y w with psi = w / laplacian (:= w at k = 0): i * (+-kd) * (psi or w)
      const bool use_k1 = f == 0 || f == 3;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float a = kd0_of(q);
        float kd = use_k1 ? kd1 : a;
        if (f == 1) kd = -kd;
        cpx<float> src = us[q * NT];
        if (f < 2) {
          const float lap = -(a * a) - (kd1 * kd1);
          src = (lap == 0.f ? 1.f : exb_rcp(lap)) * src;
        }
        v[q] = mul_i(kd * src);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        cpx<float> val = zero;
        if (col_keep && ((rowmask >> q) & 1u)) {
          ModeK<float> m;
          m.kd[0] = kd0_of(q);
          m.kd[1] = kd1;
          m.kd[2] = 0.f;
          m.keep = true;
          m.is_inj = false;
          m.is_dc = false;
          cpx<float> u[EXB_MAXC] = {us[q * NT], zero, zero};
          val = nl_inv_field<float, S>(Pn, f, u, m);
        }
        v[q] = val;
      }
    }
    fft8_run<N, +1>(v, ex, j, tw);
    if (col_keep) {
      cpx<float>* __restrict__ dst = dst0 + (size_t)f * fM;
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q * fqstride] = v[q];
    }
  }
}

#ifndef EXB_COL_LONG_LINE_REGS
#define EXB_COL_LONG_LINE_REGS 1
#endif
// Resident CTAs per SM the column kernels are compiled for (= the register cap).  The long lines (N >= 1024: four
// register-FFT passes) need more registers than the short ones: at the caps of the short lines the three-field
// epilogue spilled 300 B per thread and the plain pass 24 B (cuobjdump -res-usage), and the N = 2048 passes of config
// c5 ran at 1.5 TB/s (scripts/c5_phases.py).
template <int N, int TW, int NFWD, int MODE, int FUSE> __host__ __device__ constexpr int col_min_blocks() {
  constexpr int threads = (N / 8) * TW;
  if (MODE == COL_PLAIN) return (N >= 1024 && EXB_COL_LONG_LINE_REGS) ? 3 : 2048 / threads;
  if (MODE == COL_INV_PRO) return 2;
  if (FUSE) return 2;
  if (NFWD == 1) return 3;
  return (N >= 1024 && EXB_COL_LONG_LINE_REGS) ? 1 : EXB_EPI_MULTI_MINBLOCKS;
}
// ------------------------------------------------------------------------------- column pass
// Index bookkeeping is the expensive part of these kernels (ncu r01j: 50-60 % of the executed instructions of a
// column pass were integer / address / control, the butterflies 20-35 %), so everything that does not depend on
// the line entry is computed once per thread: 32-bit block decoding, a per-thread bit mask of the line entries
// inside the dealiasing mask, base pointers + a constant entry stride (P * line_stride elements between the 8
// entries of a thread), the wavenumber factors of the fixed axes.  ETDRK2 (the default order) gets its two stage
// updates as compile-time variants (STG) so that the operand pointers are not re-read per mode.
//   STG: 0 = order / stage at run time;  1 = ETDRK2 stage 0;  2 = ETDRK2 stage 1 (last)
//   FUSE (ETDRK2 epilogue of a one-channel 2-D kind): the value this pass produces -- stage 0: a = E u + c1 N(u),
//        stage 1: u+ -- is exactly the input of the NEXT evaluation's prologue pass, for the same column, held by
//        the same thread at the same line entries.  It is parked in a thread-private stash and the n_inv inverse
//        transforms of the next N(u) run right here (p.out = the inverse-field buffer): one launch, one read of the
//        stage input and the start-up latency of the prologue pass less per stage.
template <int N, int TW, class S, int NFWD, int MODE, int DIR, int STG = 0, int FUSE = 0>
__global__ void __launch_bounds__((N / 8) * TW, col_min_blocks<N, TW, NFWD, MODE, FUSE>())
col_fast_kernel(const ColParams<float> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int P = N / 8;
  cpx<float>* tw = reinterpret_cast<cpx<float>*>(smem_raw);
  cpx<float>* tile = tw + Fft8Tw<N>::SIZE;
  Fft8Tw<N>::fill(tw, p.tw);
  const int w = threadIdx.x % TW, j = threadIdx.x / TW;
  const unsigned inner = (unsigned)p.inner;
  const unsigned ntiles = (inner + TW - 1) / TW;
  // COL_FWD_EPI: the trajectory is the FASTEST block index when the coefficient tables do not fit the L2
  // (p.batch_fastest, set by the host: 3-D, E + c1 + c2 = 135 MB at 256^3): the CTAs that run together then read the
  // same tile of the tables once from HBM and batch - 1 times from L2.  With L2-resident tables the tile-fastest
  // order is better, neighbouring CTAs complete each other's DRAM pages (c3 lost 10 % the other way, r02a).
  const bool batch_fastest = (MODE == COL_FWD_EPI) && EXB_EPI_BATCH_FASTEST && p.batch_fastest;
  const unsigned nb = (unsigned)p.batch;
  const unsigned t = batch_fastest ? blockIdx.x / nb : blockIdx.x % ntiles;
  const unsigned rest = batch_fastest ? blockIdx.x % nb : blockIdx.x / ntiles;
  const unsigned w0 = t * TW;
  const unsigned iw = w0 + w;
  const bool act = iw < inner;
  const cpx<float> zero(0.f, 0.f);
  ExTile<TW> ex{tile + w};
  const long long ls = p.line_stride;
  const int kmax = p.P.kmax;
  // bit q: entry i = j + P*q of a (full, fftfreq-ordered) line is inside the dealiasing mask:  q < 4 -> k = i,
  // q >= 4 -> k = i - N  (i = N/2 is k = -N/2, always outside)
  unsigned rowmask = 0xffu;
  if (kmax >= 0) {
    rowmask = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ak = q < 4 ? j + P * q : N - (j + P * q);
      rowmask |= (ak <= kmax ? 1u : 0u) << q;
    }
  }
  // the fixed indices of this thread's column: 2-D axis 0 / 3-D axis 1: inner = last-axis wavenumber;
  // 3-D axis 0: inner = (i1, i2)
  const bool flat3 = inner != (unsigned)p.P.Nh;
  int i1 = (int)iw, i2 = 0, i1l = 0;   // i1l: LOCAL axis-1 index (slabs), i1: global
  if (flat3) {
    i1l = (int)(iw / (unsigned)p.P.Nh);
    i2 = (int)(iw - (unsigned)i1l * (unsigned)p.P.Nh);
    i1 = i1l * p.P.i1_mul + p.P.i1_off;
  }
  const int k1 = flat3 ? wavenumber_of(i1, N) : i1;  // (2-D: the last-axis index is the wavenumber)
  const bool col_in_mask = kmax < 0 || ((k1 < 0 ? -k1 : k1) <= kmax && i2 <= kmax);
  // are the fixed wavenumbers of this thread's column inside the mask (pruned passes only)
  const bool col_keep = act && (!((p.prune & PRUNE_COLS) && kmax >= 0) || col_in_mask);
  __syncthreads();
  const bool any_keep = __syncthreads_or(col_keep);
  const size_t qstride = (size_t)P * (size_t)ls;  // elements between the entries q and q + 1 of this thread
  // the same for the (possibly pitch-padded) field buffers of the prologue / epilogue passes: element offset of this
  // thread's entry q = 0 inside a field, entry stride, field stride
  const int fpitch = p.fpitch > 0 ? p.fpitch : p.P.Nh;
  const size_t fM = p.fpitch > 0 ? (size_t)p.fM : (size_t)p.M;
  const size_t fls = flat3 ? (size_t)(inner / (unsigned)p.P.Nh) * fpitch : (size_t)fpitch;  // (slabs: N/P axis-1 indices)
  const size_t foff = (flat3 ? (size_t)i1l * fpitch + i2 : (size_t)iw) + (size_t)j * fls;
  const size_t fqstride = (size_t)P * fls;

  if (MODE == COL_PLAIN) {
    const unsigned n_outer = (unsigned)p.n_outer;
    const unsigned o = rest % n_outer, fb = rest / n_outer;
    if (!any_keep) return;
    const size_t base = (size_t)fb * p.M + (size_t)o * p.outer_stride + iw;
    const unsigned in_mask = col_keep ? (((p.prune & PRUNE_IN_ROWS) && kmax >= 0) ? rowmask : 0xffu) : 0u;
    const unsigned out_mask = col_keep ? (((p.prune & PRUNE_OUT_ROWS) && kmax >= 0) ? rowmask : 0xffu) : 0u;
    cpx<float> v[8];
    if (p.seg_len > 0) {
      // segmented lines: the slab all-to-all buffers are used in place (entry i at (i / n) * seg_stride + (i % n) * ls)
      // owner rank / local index of line entry i: block (i / n, i % n) or cyclic (i % P, i / P) distribution
      auto seg_of = [&](int i) { return p.seg_cyclic ? i % p.seg_cyclic : i / p.seg_len; };
      auto loc_of = [&](int i) { return p.seg_cyclic ? i / p.seg_cyclic : i % p.seg_len; };
      auto line_off = [&](int i) -> size_t { return (size_t)seg_of(i) * p.seg_stride + (size_t)loc_of(i) * ls; };
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = ((in_mask >> q) & 1u) ? p.in[base + line_off(j + P * q)] : zero;
      fft8_run<N, DIR>(v, ex, j, tw);
      if (p.peer) {  // the store is the transpose: segment r of the line lives on rank r
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = j + P * q;
          if ((out_mask >> q) & 1u) p.peer_out[seg_of(i)][base + p.peer_off + (size_t)loc_of(i) * ls] = v[q];
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if ((out_mask >> q) & 1u) p.out[base + line_off(j + P * q)] = v[q];
      }
      return;
    }
    const cpx<float>* __restrict__ src = p.in + base + (size_t)j * ls;
    cpx<float>* __restrict__ dst = p.out + base + (size_t)j * ls;
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = ((in_mask >> q) & 1u) ? src[q * qstride] : zero;
    fft8_run<N, DIR>(v, ex, j, tw);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if ((out_mask >> q) & 1u) dst[q * qstride] = v[q];
    return;
  }

  const NlParams<float>& Pn = p.P;
  constexpr int C = S::C;
  const unsigned b = rest;
  // wavenumber factors: the derivative operator is i * kd  (_spectral.py:86-115: (2 pi / L) * k, rounded once)
  const float kd1 = Pn.dscale * (float)k1;
  const float kd2 = Pn.dscale * (float)i2;
  auto kd0_of = [&](int q) { return Pn.dscale * (float)(j + P * q - (q >= 4 ? N : 0)); };
  auto mode_of = [&](int q) {
    ModeK<float> m;
    m.kd[0] = kd0_of(q);
    m.kd[1] = kd1;
    m.kd[2] = S::D == 3 ? kd2 : 0.f;
    m.keep = col_in_mask && ((rowmask >> q) & 1u);
    m.is_inj = Pn.has_inj && (j + P * q) == Pn.inj_idx[0] && i1 == Pn.inj_idx[1] && (S::D < 3 || i2 == Pn.inj_idx[2]);
    m.is_dc = q == 0 && j == 0 && k1 == 0 && i2 == 0;
    return m;
  };

  if (MODE == COL_INV_PRO) {
    if (!any_keep) return;  // every column of this tile is dealiased away: nothing is written,
                            // the consumers never read masked columns (PRUNE_* contract)
    const unsigned ld_mask = col_keep ? rowmask : 0u;   // pre-dealiasing: modes outside the mask enter as zeros
    const cpx<float>* __restrict__ ubase = p.in + (size_t)b * C * p.M + iw + (size_t)j * ls;
    constexpr bool USM = (C == 1) && (EXB_INVPRO_SMEM_U != 0);
    constexpr bool VORT = S::kind == EXB_NL_VORTICITY_2D && USM;
    constexpr int NT = P * TW;
    constexpr int TROWS = ExTile<TW>::PAD ? N + N / 8 : N;
    // single-channel stage input: parked in a thread-private slice of shared memory across the field loop
    // (the kernel is capped at 64 registers for two 512-thread CTAs per SM); vorticity also parks the stream
    // function psi = w / laplacian, so the reciprocal is taken once per mode instead of once per field
    cpx<float>* ustash = tile + TROWS * TW + threadIdx.x;  // [q][thread], conflict-free
    cpx<float>* pstash = ustash + 8 * NT;
    cpx<float> u0[USM ? 1 : 8];
    if (C == 1) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const cpx<float> tq = ((ld_mask >> q) & 1u) ? ubase[q * qstride] : zero;
        if (USM) ustash[q * NT] = tq;
        else u0[q] = tq;
        if (VORT) {  // (_vorticity_convection.py:70-82: laplacian = -(kd0^2 + kd1^2), its inverse := 1 at k = 0)
          const float a = kd0_of(q);
          const float lap = -(a * a) - (kd1 * kd1);
          pstash[q * NT] = (lap == 0.f ? 1.f : exb_rcp(lap)) * tq;
        }
      }
    }
    const int f_begin = p.fcount > 0 ? p.f0 : 0, f_end = p.fcount > 0 ? p.f0 + p.fcount : Pn.n_inv;
    for (int f = f_begin; f < f_end; ++f) {
      cpx<float> v[8];
      if (VORT) {
        // u = +d_y psi, v = -d_x psi, d_x w, d_y w: every field is  i * (+-kd) * (psi or w)
        const cpx<float>* srcq = f < 2 ? pstash : ustash;
        const bool use_k1 = f == 0 || f == 3;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float kd = use_k1 ? kd1 : kd0_of(q);
          if (f == 1) kd = -kd;
          v[q] = mul_i(kd * srcq[q * NT]);   // masked entries were parked as zeros
        }
      } else if (S::kind == EXB_NL_PROJECTED_3D) {
        // fields 0..2: curl_f = i (kd_a u_b - kd_b u_a), (a, b) = (f+1, f+2) mod 3; fields 3..5: velocity
        const int ca = f >= 3 ? f - 3 : (f + 1) % 3, cb = (f + 2) % 3;
        const cpx<float>* __restrict__ ua = ubase + (size_t)ca * p.M;
        const cpx<float>* __restrict__ ub = ubase + (size_t)cb * p.M;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          cpx<float> val = zero;
          if ((ld_mask >> q) & 1u) {
            const cpx<float> xa = ua[q * qstride];
            if (f >= 3) {
              val = xa;
            } else {
              const cpx<float> xb = ub[q * qstride];
              const float k0 = kd0_of(q);
              const float ka = ca == 0 ? k0 : (ca == 1 ? kd1 : kd2), kb = cb == 0 ? k0 : (cb == 1 ? kd1 : kd2);
              val = mul_i(ka * xb - kb * xa);
            }
          }
          v[q] = val;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          cpx<float> val = zero;
          if ((ld_mask >> q) & 1u) {
            ModeK<float> m = mode_of(q);
            cpx<float> u[EXB_MAXC];
#pragma unroll
            for (int c = 0; c < EXB_MAXC; ++c)
              u[c] = c < C ? (C == 1 ? (USM ? ustash[q * NT] : u0[USM ? 0 : q]) : ubase[(size_t)c * p.M + q * qstride]) : zero;
            val = nl_inv_field<float, S>(Pn, f, u, m);
          }
          v[q] = val;
        }
      }
      fft8_run<N, DIR>(v, ex, j, tw);
      if (col_keep) {
        if (p.peer) {  // x-plane i of the result belongs to rank i / seg_len: store it there (NVLink)
          const size_t obase = ((size_t)b * Pn.n_inv + f) * fM + (foff - (size_t)j * fls);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int i = j + P * q;
            p.peer_out[i / p.seg_len][obase + p.peer_off + (size_t)(i % p.seg_len) * fls] = v[q];
          }
        } else {
          cpx<float>* __restrict__ dst = p.out + ((size_t)b * Pn.n_inv + f) * fM + foff;
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q * fqstride] = v[q];
        }
      }
    }
    return;
  }

  // COL_FWD_EPI / COL_FWD_NL
  // is the (single) injection mode outside the mask?  (then masked modes are not all N(u) == 0)
  bool inj_masked = false;
  if (Pn.has_inj && kmax >= 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (d < S::D) {
        int k = (d == S::D - 1) ? Pn.inj_idx[d] : wavenumber_of(Pn.inj_idx[d], N);
        inj_masked = inj_masked || (k < 0 ? -k : k) > kmax;
      }
    }
  }
  const int order = STG ? 2 : p.K.order;
  const int stage = STG ? STG - 1 : p.stage;
  const bool masked_skip = MODE == COL_FWD_EPI && EXB_EPI_MASKED_SKIP && kmax >= 0 && !inj_masked;
  // a tile of dealiased columns: N(u) == 0, the intermediate stages have nothing to do (and the last one neither
  // when the streaming pass etdrk_masked_linear_kernel advances the dealiased modes)
  const bool masked_here = stage == order - 1 && !p.masked_external;
  if (masked_skip && !any_keep && !masked_here) return;
  cpx<float> W[NFWD][8];
#pragma unroll
  for (int g = 0; g < NFWD; ++g) {
    const cpx<float>* __restrict__ src = p.in + ((size_t)b * NFWD + g) * fM + foff;
#pragma unroll
    for (int q = 0; q < 8; ++q) W[g][q] = col_keep ? src[q * fqstride] : zero;
  }
  if (any_keep) {  // masked columns only need the N = 0 update below
#pragma unroll
    for (int g = 0; g < NFWD; ++g) fft8_run<N, DIR>(W[g], ex, j, tw);
  }
  if (!act && !FUSE) return;
  // element offset of (trajectory b, channel 0, entry q = 0) and of the coefficient entry; + q * qstride + c * M
  const size_t off0 = (size_t)b * C * p.M + iw + (size_t)j * ls;
  const size_t ci0 = iw + (size_t)j * ls + (size_t)table_offset(p.K, (long long)b);
  const size_t cstep = p.K.E == 1 ? 0 : (size_t)p.K.M;
  constexpr int NTF = P * TW;
  constexpr int TROWSF = ExTile<TW>::PAD ? N + N / 8 : N;
  cpx<float>* fstash = tile + TROWSF * TW + threadIdx.x;   // FUSE: next stage input [q][thread], thread-private
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (FUSE) fstash[q * NTF] = zero;                       // dealiased / inactive entries enter the prologue as zeros
    if (!act) continue;
    const ModeK<float> m = mode_of(q);
    if (masked_skip && !m.keep) {
      if (masked_here) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const size_t off = off0 + q * qstride + (size_t)c * p.M;
          p.sb.OUT[off] = p.K.exp_term[ci0 + q * qstride + c * cstep] * p.sb.U[off];
        }
      }
      continue;
    }
    cpx<float> wq[NFWD], n[EXB_MAXC];
#pragma unroll
    for (int g = 0; g < NFWD; ++g) wq[g] = W[g][q];
    nl_from_fwd<float, S>(Pn, wq, m, n);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const size_t off = off0 + q * qstride + (size_t)c * p.M;
      const size_t ci = ci0 + q * qstride + c * cstep;
      if (MODE == COL_FWD_NL) {
        p.out[off] = n[c];
      } else if (STG == 1) {   // ETDRK2 stage 0 (_etdrk_2.py:96-98): a = E u + c1 N(u); keep N(u)
        const cpx<float> a = axpy(p.K.c[0][ci], n[c], p.K.exp_term[ci] * p.sb.U[off]);
        p.sb.S[0][off] = a;
        p.sb.S[1][off] = n[c];
        if (FUSE && m.keep) fstash[q * NTF] = a;
      } else if (STG == 2) {   // ETDRK2 stage 1 (_etdrk_2.py:99-101): u+ = a + c2 (N(a) - N(u))
        const cpx<float> un = axpy(p.K.c[1][ci], n[c] - p.sb.S[1][off], p.sb.S[0][off]);
        p.sb.OUT[off] = un;
        if (FUSE && m.keep) fstash[q * NTF] = un;
      } else {
        etdrk_update(p.K, stage, (long long)ci, off, n[c], p.sb);
      }
    }
  }
  if constexpr (FUSE != 0) {
    // the next evaluation's prologue pass for this column tile (one-channel 2-D kinds; tiles of dealiased columns
    // write nothing, exactly like the stand-alone pass)
    if (!any_keep) return;
    invpro_fields_1ch<N, TW, S>(Pn, fstash, NTF, j, col_keep, rowmask, kd1, 0, Pn.n_inv, ex, tw,
                                p.out + (size_t)b * Pn.n_inv * fM + foff, fM, fqstride);
  }
}

// ----------------------------------------------------------- persistent prologue pass (one-channel states)
// COL_INV_PRO for the one-channel kinds (2-D vorticity, gradient norm, polynomial) as a PERSISTENT kernel: the
// non-persistent version spends a third of its warp-stall samples at CTA start-up (ncu r02d: waiting for the
// twiddle table and for the first -- and only -- global loads of a CTA that lives for just n_inv transforms, with
// 2 CTAs per SM and nothing else to run).  Here a CTA fills its tables once and walks over (tile, trajectory) pairs
// with a stride of gridDim; the stage input of the NEXT pair is fetched with cp.async into the second half of a
// double-buffered, thread-private stash while the transforms of the current pair run.
template <int N, int TW, class S, int DIR>
__global__ void __launch_bounds__((N / 8) * TW, 2) col_invpro_persistent_kernel(const ColParams<float> p, unsigned nblk,
                                                                              unsigned ntiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  static_assert(S::C == 1, "one-channel states only");
  constexpr int P = N / 8, NT = P * TW;
  constexpr int TROWS = ExTile<TW>::PAD ? N + N / 8 : N;
  constexpr bool VORT = S::kind == EXB_NL_VORTICITY_2D;
  cpx<float>* tw = reinterpret_cast<cpx<float>*>(smem_raw);
  cpx<float>* tile = tw + Fft8Tw<N>::SIZE;
  cpx<float>* stash0 = tile + TROWS * TW + threadIdx.x;  // [buffer][q][thread]: thread-private, conflict-free
  Fft8Tw<N>::fill(tw, p.tw);
  const int w = threadIdx.x % TW, j = threadIdx.x / TW;
  const unsigned inner = (unsigned)p.inner;
  const cpx<float> zero(0.f, 0.f);
  ExTile<TW> ex{tile + w};
  const long long ls = p.line_stride;
  const NlParams<float>& Pn = p.P;
  const int kmax = Pn.kmax;
  unsigned rowmask = 0xffu;
  if (kmax >= 0) {
    rowmask = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ak = q < 4 ? j + P * q : N - (j + P * q);
      rowmask |= (ak <= kmax ? 1u : 0u) << q;
    }
  }
  const size_t qstride = (size_t)P * (size_t)ls;
  const int fpitch = p.fpitch > 0 ? p.fpitch : Pn.Nh;           // (2-D) pitch / field stride of the output fields
  const size_t fM = p.fpitch > 0 ? (size_t)p.fM : (size_t)p.M;
  const size_t fqstride = (size_t)P * fpitch;
  auto kd0_of = [&](int q) { return Pn.dscale * (float)(j + P * q - (q >= 4 ? N : 0)); };
  // asynchronous fetch of the stage input of logical block `blk` into stash buffer `buf`
  auto fetch = [&](unsigned blk, int buf) {
    const unsigned t = blk % ntiles, b = blk / ntiles;
    const unsigned iw = t * TW + w;
    const bool keep = iw < inner && (kmax < 0 || (int)iw <= kmax);
    const cpx<float>* __restrict__ ubase = p.in + (size_t)b * p.M + iw + (size_t)j * ls;
    cpx<float>* st = stash0 + buf * 8 * NT;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (keep && ((rowmask >> q) & 1u)) cp_async8(st + q * NT, ubase + q * qstride);
      else st[q * NT] = zero;   // pre-dealiasing: modes outside the mask enter as zeros
    }
    cp_async_commit();
  };
  int buf = 0;
  if (blockIdx.x < nblk) fetch(blockIdx.x, 0);
  __syncthreads();  // twiddle table
  for (unsigned blk = blockIdx.x; blk < nblk; blk += gridDim.x, buf ^= 1) {
    const unsigned t = blk % ntiles, b = blk / ntiles;
    const unsigned iw = t * TW + w;
    const int k1 = (int)iw;   // 2-D: the inner index is the last-axis wavenumber
    const bool col_keep = iw < inner && (kmax < 0 || k1 <= kmax);
    const float kd1 = Pn.dscale * (float)k1;
    cp_async_wait_all();      // this thread's slice of stash[buf] has landed (nobody else reads it)
    if (blk + gridDim.x < nblk) fetch(blk + gridDim.x, buf ^ 1);
    const cpx<float>* us = stash0 + buf * 8 * NT;
    const int f_begin = p.fcount > 0 ? p.f0 : 0, f_end = p.fcount > 0 ? p.f0 + p.fcount : Pn.n_inv;
    invpro_fields_1ch<N, TW, S>(Pn, us, NT, j, col_keep, rowmask, kd1, f_begin, f_end, ex, tw,
                                p.out + (size_t)b * Pn.n_inv * fM + iw + (size_t)j * fpitch, fM, fqstride);
  }
  cp_async_wait_all();
}

// ---------------------------------------------------------------------------------- row pass
// ROW_NL "streaming": the nonlinearities of the fast kinds are sums of products of one early and one
// late inverse field (u . grad w, u x curl u, |grad u|^2, (u . grad) u).  The first KST physical-space
// fields are parked in a thread-private slice of shared memory, every later field is folded into the
// forward-input accumulators as soon as its inverse transform is done, so only ONE inverse line plus
// the accumulators live in registers (instead of all NINV lines): 2-3x the resident warps per SM.
// row_stream_stash<S>() = KST (number of parked fields), -1: this kind is not streamed.
#ifndef EXB_ROW_STREAM
#define EXB_ROW_STREAM 1
#endif
template <class S> __host__ __device__ constexpr int row_stream_stash() {
  return !EXB_ROW_STREAM ? -1
         : S::kind == EXB_NL_VORTICITY_2D ? 2
         : S::kind == EXB_NL_PROJECTED_3D ? 3
         : (S::kind == EXB_NL_GRADIENT_NORM && S::C == 1) ? 0
         : (S::kind == EXB_NL_CONVECTION && S::var == 0 && S::C == S::D) ? S::C
         : -1;
}
template <class S, int NINV, int NFWD, int MODE> __host__ __device__ constexpr bool row_streams() {
  return MODE == ROW_NL && NINV >= 2 && row_stream_stash<S>() >= 0;
}
template <class S, int NINV, int NFWD, int MODE> __host__ __device__ constexpr int row_min_blocks() {
  return MODE != ROW_NL ? 4 : row_streams<S, NINV, NFWD, MODE>() ? (NFWD == 1 ? 3 : 2) : (NINV <= 4 ? 2 : 1);
}

// GROUPS row pairs per CTA, N/8 threads each.
template <int N, class S, int NINV, int NFWD, int MODE, int GROUPS>
__global__ void __launch_bounds__((N / 8) * GROUPS, row_min_blocks<S, NINV, NFWD, MODE>())
row_fast_kernel(const RowParams<float> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int P = N / 8, Nh = N / 2 + 1, XB = N + N / 8;
  const int ipitch = p.in_pitch > 0 ? p.in_pitch : Nh, opitch = p.out_pitch > 0 ? p.out_pitch : Nh;  // half-complex rows
  // ROW_NL with several inverse lines: the physical-space results of each line are parked in a
  // thread-private slice of shared memory (conflict-free, no synchronisation) instead of
  // NINV * 16 registers, which keeps the kernel at 2-3 CTAs per SM.
  constexpr bool STASH = false && (MODE == ROW_NL) && NINV >= 2;  // measured slower on B200 (r01g): off
  // ROW_NL: the half-complex rows of inverse line f+1 are fetched with cp.async into a small staging
  // buffer while line f is being transformed (global-load latency off the critical path, no registers).
  constexpr bool PREFETCH = (MODE == ROW_NL) && (EXB_ROW_PREFETCH != 0);
  constexpr int NHP = (Nh + 7) / 8 * 8;
  constexpr bool STREAM = row_streams<S, NINV, NFWD, MODE>();
  constexpr int KST = STREAM ? row_stream_stash<S>() : 0;
  constexpr int SLOT = XB + (STASH ? NINV * N : 0) + KST * N + (PREFETCH ? 2 * NHP : 0);  // complex elements per group
  cpx<float>* tw = reinterpret_cast<cpx<float>*>(smem_raw);
  Fft8Tw<N>::fill(tw, p.tw);
  const int g = threadIdx.x / P, j = threadIdx.x % P;
  cpx<float>* gbase = tw + Fft8Tw<N>::SIZE + (size_t)g * SLOT;
  ExLine ex{gbase, P <= 32 ? 0 : 1 + g, P};
  cpx<float>* stash = gbase + XB;
  cpx<float>* stage = gbase + XB + (STASH ? NINV * N : 0) + KST * N;
  __syncthreads();
  const NlParams<float>& Pn = p.P;
  const long long npairs = (p.rows + 1) / 2;
  const long long gp = (long long)blockIdx.x * GROUPS + g;   // global pair index
  const bool live = gp < npairs * p.batch;
  const long long b = live ? gp / npairs : 0;
  const long long rp = live ? gp - b * npairs : 0;
  const long long r1 = 2 * rp, r2 = r1 + 1;
  const bool has1 = live, has2 = live && r2 < p.rows;
  const cpx<float> zero(0.f, 0.f);
  const int kin = ((p.prune & PRUNE_IN_ROWS) && p.P.kmax >= 0) ? p.P.kmax : N;    // load k <= kin only
  const int kout = ((p.prune & PRUNE_OUT_ROWS) && p.P.kmax >= 0) ? p.P.kmax : N;  // store k <= kout only

  // two-for-one split of a forward line in registers; calls fn(k, X1, X2) for owned modes
  auto unpack_store = [&](cpx<float> (&v)[8], cpx<float>* o1, cpx<float>* o2) {
    ex.sync();
#pragma unroll
    for (int q = 4; q < 8; ++q) ex.st(j + P * q, v[q]);
    ex.sync();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = j + P * q;
      cpx<float> zk = v[q];
      cpx<float> zp = (k == 0) ? zk : ex.ld(N - k);
      if (has1 && k <= kout) o1[k] = 0.5f * (zk + conj(zp));
      if (has2 && k <= kout) o2[k] = mul_mi(0.5f * (zk - conj(zp)));
    }
    if (j == 0 && N / 2 <= kout) {
      if (has1) o1[N / 2] = cpx<float>(v[4].x, 0.f);
      if (has2) o2[N / 2] = cpx<float>(v[4].y, 0.f);
    }
  };
  // packed inverse line from the half-complex rows i1p, i2p (irfft semantics at DC / Nyquist)
  auto load_packed = [&](const cpx<float>* i1p, const cpx<float>* i2p, cpx<float> (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int n = j + P * q;
      const bool upper = n > N / 2;
      const int k = upper ? N - n : n;
      cpx<float> F1 = (has1 && k <= kin) ? i1p[k] : zero;
      cpx<float> F2 = (has2 && k <= kin) ? i2p[k] : zero;
      cpx<float> z;
      if (k == 0 || 2 * k == N) z = cpx<float>(F1.x, F2.x);
      else if (!upper) z = F1 + mul_i(F2);
      else z = conj(F1) + mul_i(conj(F2));
      v[q] = z;
    }
  };

  if (MODE == ROW_R2C) {
    const float* in = (const float*)p.in + (size_t)b * p.in_batch_stride;
    cpx<float>* out = (cpx<float>*)p.out + (size_t)b * p.out_batch_stride;
    for (int f = 0; f < p.nin; ++f) {
      cpx<float> v[8];
      const float* a = in + ((size_t)f * p.rows + r1) * N;
      const float* c = in + ((size_t)f * p.rows + r2) * N;
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = cpx<float>(has1 ? a[j + P * q] : 0.f, has2 ? c[j + P * q] : 0.f);
      fft8_run<N, -1>(v, ex, j, tw);
      unpack_store(v, out + ((size_t)f * p.rows + r1) * opitch, out + ((size_t)f * p.rows + r2) * opitch);
    }
    return;
  }
  if (MODE == ROW_C2R) {
    const cpx<float>* in = (const cpx<float>*)p.in + (size_t)b * p.in_batch_stride;
    float* out = (float*)p.out + (size_t)b * p.out_batch_stride;
    for (int f = 0; f < p.nin; ++f) {
      cpx<float> v[8];
      load_packed(in + ((size_t)f * p.rows + r1) * ipitch, in + ((size_t)f * p.rows + r2) * ipitch, v);
      fft8_run<N, +1>(v, ex, j, tw);
      float* a = out + ((size_t)f * p.rows + r1) * N;
      float* c = out + ((size_t)f * p.rows + r2) * N;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (has1) a[j + P * q] = v[q].x * Pn.inv_norm;
        if (has2) c[j + P * q] = v[q].y * Pn.inv_norm;
      }
    }
    return;
  }

  // ROW_NL
  const cpx<float>* in = (const cpx<float>*)p.in + (size_t)b * p.in_batch_stride;
  cpx<float>* out = (cpx<float>*)p.out + (size_t)b * p.out_batch_stride;
  cpx<float> wl[NFWD][8];
  if (STREAM) {
#pragma unroll
    for (int gg = 0; gg < NFWD; ++gg)
#pragma unroll
      for (int q = 0; q < 8; ++q) wl[gg][q] = zero;
    cpx<float>* park = stash + j;  // field g, point q of this thread at park[(g * 8 + q) * P]
    auto fetch = [&](int f) {      // asynchronous copy of the two half-complex rows of inverse field f
      const cpx<float>* a = in + ((size_t)f * p.rows + r1) * ipitch;
      const cpx<float>* c = in + ((size_t)f * p.rows + r2) * ipitch;
      for (int k = j; k < Nh && k <= kin; k += P) {
        if (has1) cp_async8(stage + k, a + k);
        if (has2) cp_async8(stage + NHP + k, c + k);
      }
      cp_async_commit();
    };
    if (PREFETCH) fetch(0);
#pragma unroll
    for (int f = 0; f < NINV; ++f) {
      cpx<float> z[8];
      if (PREFETCH) {
        cp_async_wait_all();
        ex.sync();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = j + P * q;
          const bool upper = n > N / 2;
          const int k = upper ? N - n : n;
          cpx<float> F1 = (has1 && k <= kin) ? stage[k] : zero;
          cpx<float> F2 = (has2 && k <= kin) ? stage[NHP + k] : zero;
          cpx<float> zz;
          if (k == 0 || 2 * k == N) zz = cpx<float>(F1.x, F2.x);
          else if (!upper) zz = F1 + mul_i(F2);
          else zz = conj(F1) + mul_i(conj(F2));
          z[q] = zz;
        }
        ex.sync();
        if (f + 1 < NINV) fetch(f + 1);
      } else {
        load_packed(in + ((size_t)f * p.rows + r1) * ipitch, in + ((size_t)f * p.rows + r2) * ipitch, z);
      }
      fft8_run<N, +1>(z, ex, j, tw);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const cpx<float> a = Pn.inv_norm * z[q];  // (row r1, row r2) at point j + P*q
        auto pk = [&](int g) { return park[(g * 8 + q) * P]; };
        auto acc = [&](int gg, cpx<float> s, float sign) {  // lane-wise (+-s) * a + wl: one packed FMA
          wl[gg][q] = lane_fma(sign < 0.f ? -s : s, a, wl[gg][q]);
        };
        if (f < KST) {
          park[(f * 8 + q) * P] = a;
        } else if (S::kind == EXB_NL_VORTICITY_2D) {    // u w_x + v w_y       (fields: u, v, w_x, w_y)
          acc(0, pk(f - 2), 1.f);
        } else if (S::kind == EXB_NL_PROJECTED_3D) {    // velocity x curl     (fields: curl 0..2, velocity 3..5)
          if (f == 3) { acc(1, pk(2), -1.f); acc(2, pk(1), 1.f); }
          if (f == 4) { acc(0, pk(2), 1.f); acc(2, pk(0), -1.f); }
          if (f == 5) { acc(0, pk(1), -1.f); acc(1, pk(0), 1.f); }
        } else if (S::kind == EXB_NL_GRADIENT_NORM) {   // sum_d (d_d u)^2
          acc(0, a, 1.f);
        } else if (S::kind == EXB_NL_CONVECTION) {      // out_c = sum_d u_d d_d u_c (fields: u_0.., then d_d u_c at C + c*D + d)
          constexpr int Cc = S::C, Dd = S::D;
          const int idx = f - Cc;
          acc(idx / Dd, pk(idx % Dd), 1.f);
        }
      }
    }
  } else if (STASH) {
    for (int f = 0; f < NINV; ++f) {
      cpx<float> z[8];
      load_packed(in + ((size_t)f * p.rows + r1) * ipitch, in + ((size_t)f * p.rows + r2) * ipitch, z);
      fft8_run<N, +1>(z, ex, j, tw);
#pragma unroll
      for (int q = 0; q < 8; ++q) stash[(size_t)f * N + j + P * q] = z[q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      f32x2 iv[NINV], ov[NFWD];
#pragma unroll
      for (int f = 0; f < NINV; ++f) iv[f] = lanes(Pn.inv_norm * stash[(size_t)f * N + j + P * q]);
      nl_pointwise<float, S, f32x2>(Pn, iv, ov);
#pragma unroll
      for (int gg = 0; gg < NFWD; ++gg) wl[gg][q] = as_cpx(ov[gg]);
    }
  } else {
    cpx<float> z[NINV][8];
    if (PREFETCH) {
      auto prefetch = [&](int f) {
        const cpx<float>* a = in + ((size_t)f * p.rows + r1) * ipitch;
        const cpx<float>* c = in + ((size_t)f * p.rows + r2) * ipitch;
        for (int k = j; k < Nh && k <= kin; k += P) {
          if (has1) cp_async8(stage + k, a + k);
          if (has2) cp_async8(stage + NHP + k, c + k);
        }
        cp_async_commit();
      };
      prefetch(0);
#pragma unroll
      for (int f = 0; f < NINV; ++f) {
        cp_async_wait_all();
        ex.sync();  // the staged rows are visible to the whole group
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = j + P * q;
          const bool upper = n > N / 2;
          const int k = upper ? N - n : n;
          cpx<float> F1 = (has1 && k <= kin) ? stage[k] : zero;
          cpx<float> F2 = (has2 && k <= kin) ? stage[NHP + k] : zero;
          cpx<float> zz;
          if (k == 0 || 2 * k == N) zz = cpx<float>(F1.x, F2.x);
          else if (!upper) zz = F1 + mul_i(F2);
          else zz = conj(F1) + mul_i(conj(F2));
          z[f][q] = zz;
        }
        ex.sync();  // staging buffer consumed
        if (f + 1 < NINV) prefetch(f + 1);
        fft8_run<N, +1>(z[f], ex, j, tw);
      }
    } else {
#pragma unroll
      for (int f = 0; f < NINV; ++f) {
        load_packed(in + ((size_t)f * p.rows + r1) * ipitch, in + ((size_t)f * p.rows + r2) * ipitch, z[f]);
        fft8_run<N, +1>(z[f], ex, j, tw);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      f32x2 iv[NINV], ov[NFWD];   // the two lanes = the same grid point of rows r1 and r2
#pragma unroll
      for (int f = 0; f < NINV; ++f) iv[f] = lanes(Pn.inv_norm * z[f][q]);
      nl_pointwise<float, S, f32x2>(Pn, iv, ov);
#pragma unroll
      for (int gg = 0; gg < NFWD; ++gg) wl[gg][q] = as_cpx(ov[gg]);
    }
  }
#pragma unroll
  for (int gg = 0; gg < NFWD; ++gg) {
    fft8_run<N, -1>(wl[gg], ex, j, tw);
    unpack_store(wl[gg], out + ((size_t)gg * p.rows + r1) * opitch, out + ((size_t)gg * p.rows + r2) * opitch);
  }
}

}  // namespace exb
