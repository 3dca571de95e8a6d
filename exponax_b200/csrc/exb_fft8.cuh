// Register-resident FFT core of the fast 2-D / 3-D kernels: an N-point complex line
// (N = 64, 128, 256, 512; N = 8 * P) is held by P threads, 8 points each.  Two or three
// in-register radix passes (8 x 8 x N/64) are joined by exchanges through shared memory; the
// caller supplies the exchange addressing (contiguous padded line, or a column of a [N][TW]
// tile) and the synchronisation (warp, named barrier or CTA).
//   entry: v[q] = x[j + P*q]      exit: v[q] = X[j + P*q]      (unnormalised; DIR=-1 forward)
#pragma once
#include "exb_fft.cuh"

namespace exb {

template <int DIR, int R> __device__ __forceinline__ void dft_small(cpx<float>* a) {
  if (R == 8) dft8<float, DIR>(a);
  if (R == 4) dft4<float, DIR>(a);
  if (R == 2) dft2<float, DIR>(a[0], a[1]);
}

// Pass structure: N = 64: 8 x 8;  N = 128/256/512: 8 x 8 x (N/64);  N = 1024/2048: 8 x 8 x 8 x (N/512).
template <int N> struct Fft8Cfg {
  static constexpr int P = N / 8;
  static constexpr int NP = N == 64 ? 2 : (N <= 512 ? 3 : 4);          // number of passes
  static constexpr int RL = N == 64 ? 8 : (N <= 512 ? N / 64 : N / 512);  // radix of the last pass
};

// Twiddle table (shared memory, forward sign), arranged per pass so that the lanes of a warp
// read consecutive entries (bank-conflict free):
//   T2[(r-1)*8 + k]    = w_64^(k*r)        r = 1..7, k = 0..7                     (pass 2)
//   3 passes: T3[(r-1)*64 + b]  = w_N^(b*r)    r = 1..RL-1, b = 0..63             (pass 3, last)
//   4 passes: T3[(r-1)*64 + k]  = w_512^(k*r)  r = 1..7,    k = 0..63             (pass 3)
//             T4[(r-1)*512 + b] = w_N^(b*r)    r = 1..RL-1, b = 0..511            (pass 4, last)
template <int N> struct Fft8Tw {
  using Cfg = Fft8Cfg<N>;
  static constexpr int T3 = 56;
  static constexpr int N3 = Cfg::NP == 2 ? 0 : (Cfg::NP == 3 ? (Cfg::RL - 1) * 64 : 7 * 64);
  static constexpr int T4 = T3 + N3;
  static constexpr int N4 = Cfg::NP == 4 ? (Cfg::RL - 1) * 512 : 0;
  static constexpr int SIZE = T4 + N4;
  // index into the N-th roots of unity exp(-2 pi i j / N) of table entry q
  static __host__ __device__ constexpr int root_index(int q) {
    return q < T3   ? (N / 64) * (q % 8) * (q / 8 + 1)
           : q < T4 ? (Cfg::NP == 4 ? (N / 512) : 1) * ((q - T3) % 64) * ((q - T3) / 64 + 1)
                    : ((q - T4) % 512) * ((q - T4) / 512 + 1);
  }
  // The plan uploads the table ALREADY ARRANGED right behind the N roots (exb_fastnd_tw_arrange), so a CTA fills its
  // shared-memory copy with one coalesced pass -- the former per-CTA gather (index arithmetic + scattered loads)
  // was 5 % of the instructions of a column-pass CTA (ncu r01j, exb_fft8.cuh:40-46).
  static __device__ __forceinline__ void fill(cpx<float>* t, const cpx<float>* __restrict__ roots) {
    const cpx<float>* __restrict__ arranged = roots + N;
    for (int q = threadIdx.x; q < SIZE; q += blockDim.x) t[q] = arranged[q];
  }
};

// The R - 1 twiddles w^r (r = 1 .. R-1) of one butterfly.  DERIVE: only w^1, w^2, w^4 are read from the shared-memory
// table, the others are products of two of them (w^3 = w^1 w^2, w^5 = w^1 w^4, w^6 = w^2 w^4, w^7 = w^3 w^4: at most
// two extra roundings).  The row kernels are bound by the shared-memory pipe (88-92 % LSU wavefront utilisation,
// 17 % of the wavefronts were twiddle loads -- ncu r02d) with the FMA pipe at 35 %: 4 loads become 8 packed FP ops.
template <int R, int DIR, bool DERIVE>
__device__ __forceinline__ void load_twiddles(cpx<float> (&w)[8], const cpx<float>* __restrict__ t, int stride) {
  if (!DERIVE || R < 4) {
#pragma unroll
    for (int r = 1; r < R; ++r) w[r] = twd<float, DIR>(t[(r - 1) * stride]);
    return;
  }
  w[1] = twd<float, DIR>(t[0]);
  w[2] = twd<float, DIR>(t[stride]);
  w[3] = w[1] * w[2];
  if (R == 8) {
    w[4] = twd<float, DIR>(t[3 * stride]);
    w[5] = w[1] * w[4];
    w[6] = w[2] * w[4];
    w[7] = w[3] * w[4];
  }
}

// last pass: radix RL with Ns = N / RL; 8/RL butterflies per thread, results stay in registers
template <int N, int DIR, int TOFF, bool DERIVE>
__device__ __forceinline__ void fft8_last_pass(cpx<float> (&v)[8], int j, const cpx<float>* __restrict__ tw) {
  constexpr int P = N / 8, RL = Fft8Cfg<N>::RL, NS = N / RL, NB = 8 / RL;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int b = j + P * i;
    cpx<float> a[RL], w[8];
    load_twiddles<RL, DIR, DERIVE>(w, tw + TOFF + b, NS);
#pragma unroll
    for (int r = 0; r < RL; ++r) {
      a[r] = v[i + NB * r];
      if (r > 0) a[r] = a[r] * w[r];
    }
    dft_small<DIR, RL>(a);
#pragma unroll
    for (int r = 0; r < RL; ++r) v[i + NB * r] = a[r];
  }
}

// Ex: struct with  void st(int i, cpx<float>) const;  cpx<float> ld(int i) const;  void sync() const;
template <int N, int DIR, class Ex>
__device__ __forceinline__ void fft8_run(cpx<float> (&v)[8], const Ex& ex, int j, const cpx<float>* __restrict__ tw) {
  using Cfg = Fft8Cfg<N>;
  using Tw = Fft8Tw<N>;
  constexpr int P = Cfg::P;
  static_assert(N == 64 || N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048, "fft8_run: unsupported N");
  // ---- pass 1: radix 8, Ns = 1 ----
  dft8<float, DIR>(v);
  ex.sync();
#pragma unroll
  for (int r = 0; r < 8; ++r) ex.st(8 * j + r, v[r]);
  ex.sync();
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = ex.ld(j + P * q);
  // ---- pass 2: radix 8, Ns = 8 ----
  {
    const int k = j & 7;
    {
      cpx<float> w[8];
      load_twiddles<8, DIR, Ex::kDeriveTw>(w, tw + k, 8);
#pragma unroll
      for (int r = 1; r < 8; ++r) v[r] = v[r] * w[r];
    }
    dft8<float, DIR>(v);
    if (Cfg::NP == 2) return;  // N = 64: outputs already at j + 8*r
    const int j0 = (j - k) * 8 + k;
    ex.sync();
#pragma unroll
    for (int r = 0; r < 8; ++r) ex.st(j0 + 8 * r, v[r]);
    ex.sync();
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = ex.ld(j + P * q);
  }
  if (Cfg::NP == 3) {
    fft8_last_pass<N, DIR, Tw::T3, Ex::kDeriveTw>(v, j, tw);
    return;
  }
  // ---- pass 3 of 4: radix 8, Ns = 64 ----
  {
    const int k = j & 63;
    {
      cpx<float> w[8];
      load_twiddles<8, DIR, Ex::kDeriveTw>(w, tw + Tw::T3 + k, 64);
#pragma unroll
      for (int r = 1; r < 8; ++r) v[r] = v[r] * w[r];
    }
    dft8<float, DIR>(v);
    const int j0 = (j - k) * 8 + k;
    ex.sync();
#pragma unroll
    for (int r = 0; r < 8; ++r) ex.st(j0 + 64 * r, v[r]);
    ex.sync();
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = ex.ld(j + P * q);
  }
  fft8_last_pass<N, DIR, Tw::T4, Ex::kDeriveTw>(v, j, tw);
}

// exchange through a contiguous line buffer (row kernels).  Element i lives at the XOR-swizzled slot
//   i ^ ((i >> 4) & 7) ^ ((i >> 3) & 8)
// which makes every access pattern of fft8_run conflict-free per half-warp (8-byte elements, 16 bank
// pairs): the stride-8 stores of pass 1, the k + 64 m stores of pass 2, the stride-64 stores of pass 3
// and the unit-stride loads.  (The former padding i + i/8 cost the loads a second wavefront: 16
// consecutive elements spanned 17 slots.)  EXB_LINE_SWIZZLE=0 restores the padded layout.
#ifndef EXB_LINE_SWIZZLE
#define EXB_LINE_SWIZZLE 1
#endif
#ifndef EXB_ROW_DERIVE_TW
#define EXB_ROW_DERIVE_TW 1
#endif
struct ExLine {
  static constexpr bool kDeriveTw = EXB_ROW_DERIVE_TW != 0;   // row kernels: shared-memory-pipe bound
  cpx<float>* base;
  int sync_kind;  // 0: __syncwarp, otherwise named barrier id
  int nthreads;
  static __device__ __forceinline__ int slot(int i) {
    return EXB_LINE_SWIZZLE ? (i ^ ((i >> 4) & 7) ^ ((i >> 3) & 8)) : (i + (i >> 3));
  }
  __device__ __forceinline__ void st(int i, cpx<float> x) const { base[slot(i)] = x; }
  __device__ __forceinline__ cpx<float> ld(int i) const { return base[slot(i)]; }
  __device__ __forceinline__ void sync() const {
    if (sync_kind == 0) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(sync_kind), "r"(nthreads) : "memory");
  }
};

// exchange through column w of a [N][TW] tile (column kernels): element i at i*TW + w
// (rows narrower than 128 B get one pad row per 8 rows so that the stride-8-row stores of
// pass 1 spread over all banks)
template <int TW> struct ExTile {
  static constexpr bool kDeriveTw = false;                     // column kernels: not bound by the shared-memory pipe
  static constexpr bool PAD = TW < 16;
  static constexpr int rows(int n) { return PAD ? n + n / 8 : n; }
  cpx<float>* base;  // tile + w
  __device__ __forceinline__ void st(int i, cpx<float> x) const { base[(PAD ? i + (i >> 3) : i) * TW] = x; }
  __device__ __forceinline__ cpx<float> ld(int i) const { return base[(PAD ? i + (i >> 3) : i) * TW]; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

}  // namespace exb
