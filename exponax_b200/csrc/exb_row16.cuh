// 16-points-per-thread row pass for N = 256 (config c4), the default since round 2 (-DEXB_ROW16=0 restores the
// 8-points-per-thread row_fast_kernel for this size).
//
// Why: the ROW_NL pass is bound by the shared-memory pipe (86-92 % LSU wavefront utilisation): an 8-points-per-thread
// line FFT exchanges every point twice (three radix passes).  With 16 points per thread N = 256 = 16 x 16 needs ONE
// exchange (the 1-D kernel's fft_reg<16>), the twiddle loads per point halve too; the price is 2x the registers per
// thread (1 CTA of 256 threads per SM, 16 row pairs in flight per SM as before).  Same math, same streaming
// accumulation, same pruning as row_fast_kernel; covered by every 3-D test at N = 256 (tests/test_gpu_parity.py:
// test_full_size_c4_navier_stokes_256_benchmarked_instantiation) and the 2-D fast-kind tests at N = 256.
#pragma once
#include "exb_kernels_1d_fast.cuh"
#include "exb_kernels_nd_fast.cuh"

namespace exb {

template <class S, int NINV, int NFWD>
__global__ void __launch_bounds__(256, 1) row16_kernel(const RowParams<float> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int R = 16, N = R * R, Nh = N / 2 + 1, P = R, GROUPS = 256 / P;
  constexpr int XB = (R + 1) * R;                 // exchange buffer, slot(i) = i + i / 16
  constexpr int NHP = (Nh + 7) / 8 * 8;
  constexpr int KST = row_stream_stash<S>();
  static_assert(KST >= 0 && NINV >= 2, "row16_kernel: streaming nonlinear functions only");
  constexpr int SLOT = XB + KST * N + 2 * NHP;    // complex elements per group
  cpx<float>* tw2 = reinterpret_cast<cpx<float>*>(smem_raw);
  for (int q = threadIdx.x; q < R * R; q += blockDim.x) tw2[q] = p.tw[((q / R) * (q % R)) % N];
  const int g = threadIdx.x / P, j = threadIdx.x % P;
  cpx<float>* xb = tw2 + R * R + (size_t)g * SLOT;
  cpx<float>* park = xb + XB + j;                 // field f, point r of this thread at park[(f * R + r) * P]
  cpx<float>* stage = xb + XB + KST * N;
  __syncthreads();
  const NlParams<float>& Pn = p.P;
  const long long npairs = (p.rows + 1) / 2;
  const long long gp = (long long)blockIdx.x * GROUPS + g;
  const bool live = gp < npairs * p.batch;
  const long long b = live ? gp / npairs : 0;
  const long long rp = live ? gp - b * npairs : 0;
  const long long r1 = 2 * rp, r2 = r1 + 1;
  const bool has1 = live, has2 = live && r2 < p.rows;
  const cpx<float> zero(0.f, 0.f);
  const int kin = ((p.prune & PRUNE_IN_ROWS) && Pn.kmax >= 0) ? Pn.kmax : N;
  const int kout = ((p.prune & PRUNE_OUT_ROWS) && Pn.kmax >= 0) ? Pn.kmax : N;
  const cpx<float>* in = (const cpx<float>*)p.in + (size_t)b * p.in_batch_stride;
  cpx<float>* out = (cpx<float>*)p.out + (size_t)b * p.out_batch_stride;
  const int ipitch = p.in_pitch > 0 ? p.in_pitch : Nh, opitch = p.out_pitch > 0 ? p.out_pitch : Nh;

  auto fetch = [&](int f) {
    const cpx<float>* a = in + ((size_t)f * p.rows + r1) * ipitch;
    const cpx<float>* c = in + ((size_t)f * p.rows + r2) * ipitch;
    for (int k = j; k < Nh && k <= kin; k += P) {
      if (has1) cp_async8(stage + k, a + k);
      if (has2) cp_async8(stage + NHP + k, c + k);
    }
    cp_async_commit();
  };
  cpx<float> wl[NFWD][R];
#pragma unroll
  for (int gg = 0; gg < NFWD; ++gg)
#pragma unroll
    for (int r = 0; r < R; ++r) wl[gg][r] = zero;
  fetch(0);
#pragma unroll
  for (int f = 0; f < NINV; ++f) {
    cpx<float> z[R];
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int n = j + P * r;
      const bool upper = n > N / 2;
      const int k = upper ? N - n : n;
      cpx<float> F1 = (has1 && k <= kin) ? stage[k] : zero;
      cpx<float> F2 = (has2 && k <= kin) ? stage[NHP + k] : zero;
      cpx<float> zz;
      if (k == 0 || 2 * k == N) zz = cpx<float>(F1.x, F2.x);
      else if (!upper) zz = cpx<float>(F1.x - F2.y, F1.y + F2.x);
      else zz = cpx<float>(F1.x + F2.y, F2.x - F1.y);
      z[r] = zz;
    }
    __syncwarp();
    if (f + 1 < NINV) fetch(f + 1);
    fft_reg<R, +1>(z, xb, j, tw2);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const cpx<float> a(z[r].x * Pn.inv_norm, z[r].y * Pn.inv_norm);
      auto pk = [&](int fld) { return park[(fld * R + r) * P]; };
      auto acc = [&](int gg, cpx<float> s, float sign) {
        wl[gg][r].x = fmaf(sign * s.x, a.x, wl[gg][r].x);
        wl[gg][r].y = fmaf(sign * s.y, a.y, wl[gg][r].y);
      };
      if (f < KST) {
        park[(f * R + r) * P] = a;
      } else if (S::kind == EXB_NL_VORTICITY_2D) {
        acc(0, pk(f - 2), 1.f);
      } else if (S::kind == EXB_NL_PROJECTED_3D) {
        if (f == 3) { acc(1, pk(2), -1.f); acc(2, pk(1), 1.f); }
        if (f == 4) { acc(0, pk(2), 1.f); acc(2, pk(0), -1.f); }
        if (f == 5) { acc(0, pk(1), -1.f); acc(1, pk(0), 1.f); }
      } else if (S::kind == EXB_NL_GRADIENT_NORM) {
        acc(0, a, 1.f);
      } else if (S::kind == EXB_NL_CONVECTION) {
        constexpr int Cc = S::C, Dd = S::D;
        const int idx = f - Cc;
        acc(idx / Dd, pk(idx % Dd), 1.f);
      }
    }
  }
  auto slot = [](int i) { return i + (i >> 4); };
#pragma unroll
  for (int gg = 0; gg < NFWD; ++gg) {
    fft_reg<R, -1>(wl[gg], xb, j, tw2);
    cpx<float>* o1 = out + ((size_t)gg * p.rows + r1) * opitch;
    cpx<float>* o2 = out + ((size_t)gg * p.rows + r2) * opitch;
    __syncwarp();
#pragma unroll
    for (int r = R / 2; r < R; ++r) xb[slot(j + P * r)] = wl[gg][r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R / 2; ++r) {
      const int k = j + P * r;
      const cpx<float> zk = wl[gg][r];
      const cpx<float> zp = (k == 0) ? zk : xb[slot(N - k)];
      if (has1 && k <= kout) o1[k] = cpx<float>(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
      if (has2 && k <= kout) o2[k] = cpx<float>(0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x));
    }
    if (j == 0 && N / 2 <= kout) {
      if (has1) o1[N / 2] = cpx<float>(wl[gg][R / 2].x, 0.f);
      if (has2) o2[N / 2] = cpx<float>(wl[gg][R / 2].y, 0.f);
    }
  }
}

template <class S, int NINV, int NFWD> constexpr size_t row16_smem() {
  constexpr int R = 16, N = 256, NHP = (N / 2 + 1 + 7) / 8 * 8;
  return (size_t)(R * R + (256 / R) * ((R + 1) * R + row_stream_stash<S>() * N + 2 * NHP)) * sizeof(cpx<float>);
}

}  // namespace exb
