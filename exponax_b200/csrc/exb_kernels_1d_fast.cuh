// Fast 1-D persistent rollout kernel for N = R*R (R = 16 -> N = 256, R = 8 -> N = 64), f32, C = 1.
//
// Work decomposition (B200: 148 SMs, 64K regs, 227 KB smem / SM):
//   * a PAIR of trajectories shares every complex FFT (two-for-one: z = x1 + i*x2);
//   * a pair is owned by R threads (a half-warp for N = 256) for the whole rollout: every
//     synchronisation is a __syncwarp, there is no __syncthreads after the table load;
//   * each thread keeps R points of the line in registers; an N-point FFT is two in-register
//     radix-R passes with ONE shared-memory exchange in between (padded, conflict-free);
//   * the pointwise nonlinearity happens in registers between the inverse and forward
//     transforms (the output mapping of the inverse FFT is the input mapping of the forward);
//   * spectral state, ETDRK stage buffers and coefficient tables live in shared memory; HBM
//     sees the initial condition once and the saved snapshots;
//   * the nonlinear function is a compile-time descriptor S (NlS<...>): no per-element branching.
// Reference semantics identical to k1d_kernel (exb_kernels_1d.cuh); parity is tested against
// the same oracle.
#pragma once
#include "exb_kernels_1d.cuh"

namespace exb {

struct FastLayout {
  int off_tw2;       // [R][R] inter-pass twiddles w_N^(j*r), forward sign
  int off_exp;       // cpx[Nh]
  int off_hexp;      // cpx[Nh]
  int off_c[6];      // float[Nh] each
  int off_h;         // float[Nh]: mask * post-factor / 2 of the nonlinear function (see unpack_scaled)
  int off_mk;        // float2[Nh]: (keep / N, keep * kd / N) -- pre-dealiasing mask, derivative scale and 1/N of the inverse transform
  int off_pairs;     // start of per-pair storage
  int pair_bytes;    // bytes per pair
  int nstate;        // spectral state arrays per pair (1 + scratch)
  int nhp;           // padded Nh (complex elements) per trajectory
};

// ---- in-register DFTs; output position p holds X[xidx<R>(p)] --------------------------------
template <int R> __host__ __device__ constexpr int xidx(int p) { return R == 16 ? 4 * (p & 3) + (p >> 2) : p; }

template <int DIR> __device__ __forceinline__ void dft16_reg(cpx<float>* v) {
  // n = n1 + 4*n2, k = 4*k1 + k2 :  w16^(nk) = w4^(n1 k1) * w16^(n1 k2) * w4^(n2 k2)
  const float c = 0.92387953251128675613f, s = 0.38268343236508977173f, h = 0.70710678118654752440f;
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    cpx<float> t[4] = {v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]};
    dft4<float, DIR>(t);
    v[n1] = t[0];
    v[n1 + 4] = t[1];
    v[n1 + 8] = t[2];
    v[n1 + 12] = t[3];
  }
  // twiddles w16^(n1*k2) on v[n1 + 4*k2]; forward twiddle = (wr, -wi)
  // (a complex product with a compile-time constant: two packed instructions with immediate operands)
#define EXB_MULW(a, wr, wi) a = a * (DIR < 0 ? cpx<float>((wr), -(wi)) : cpx<float>((wr), (wi)))
  EXB_MULW(v[1 + 4], c, s);     // w^1
  EXB_MULW(v[2 + 4], h, h);     // w^2
  EXB_MULW(v[3 + 4], s, c);     // w^3
  EXB_MULW(v[1 + 8], h, h);     // w^2
  v[2 + 8] = rot90<float, DIR>(v[2 + 8]);  // w^4 = -i
  EXB_MULW(v[3 + 8], -h, h);    // w^6
  EXB_MULW(v[1 + 12], s, c);    // w^3
  EXB_MULW(v[2 + 12], -h, h);   // w^6
  EXB_MULW(v[3 + 12], -c, -s);  // w^9
#undef EXB_MULW
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) dft4<float, DIR>(v + 4 * k2);  // v[4*k2 + k1] = X[4*k1 + k2]
}

template <int R, int DIR> __device__ __forceinline__ void dft_reg(cpx<float>* v) {
  if (R == 16) dft16_reg<DIR>(v);
  if (R == 8) dft8<float, DIR>(v);
}

// (round 2, lean kernel: the shared-memory pipe is the busiest unit (ncu 87 %), so deriving 11 of the 15 inter-pass
// twiddles from w, w^2, w^4, w^8 pays: c2 1.73e11 -> 1.79e11; with the round-1 kernel, issue-bound, it was a loss)
#ifndef EXB_1D_DERIVE_TW
#define EXB_1D_DERIVE_TW 1
#endif
// N = R*R point FFT of the line held as v[r] = x[j + R*r] by the R threads j of a group.
// On return v[r] = X[j + R*r].  xb: per-pair exchange buffer ((R+1)*R complex, padded).
template <int R, int DIR>
__device__ __forceinline__ void fft_reg_body(cpx<float> (&v)[R], cpx<float>* xb, int j, const cpx<float>* tw2) {
  dft_reg<R, DIR>(v);
  __syncwarp();
#pragma unroll
  for (int p = 0; p < R; ++p) xb[(R + 1) * j + xidx<R>(p)] = v[p];  // pad(R*j + X-index)
  __syncwarp();
#if EXB_1D_DERIVE_TW
  {
    // inter-pass twiddles w^(r j), r = 1 .. R-1: only the powers of two are read from shared memory, the others are
    // products of at most 3 of them (the shared-memory pipe carries 17 % twiddle loads, the FMA pipe has headroom)
    cpx<float> w[R];
#pragma unroll
    for (int b = 1; b < R; b <<= 1) w[b] = twd<float, DIR>(tw2[b * R + j]);
#pragma unroll
    for (int r = 3; r < R; ++r)
      if (r & (r - 1)) w[r] = w[r & (r - 1)] * w[r & -r];          // r = (r without its lowest bit) + lowest bit
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cpx<float> t = xb[j + (R + 1) * r];
      v[r] = r > 0 ? t * w[r] : t;
    }
  }
#else
#pragma unroll
  for (int r = 0; r < R; ++r) {
    cpx<float> t = xb[j + (R + 1) * r];                             // pad(j + R*r)
    if (r > 0) t = t * twd<float, DIR>(tw2[r * R + j]);
    v[r] = t;
  }
#endif
  dft_reg<R, DIR>(v);
  if (R == 16) {
    cpx<float> o[R];
#pragma unroll
    for (int p = 0; p < R; ++p) o[xidx<R>(p)] = v[p];
#pragma unroll
    for (int p = 0; p < R; ++p) v[p] = o[p];
  }
}

#ifndef EXB_1D_SHFL_UNPACK
#define EXB_1D_SHFL_UNPACK 1
#endif
#ifndef EXB_1D_CTA_SYNC
#define EXB_1D_CTA_SYNC 0
#endif
#ifndef EXB_1D_SPLIT_BAR
#define EXB_1D_SPLIT_BAR 0
#endif
#ifndef EXB_1D_SKEW_NS
#define EXB_1D_SKEW_NS 0
#endif

// The step body calls the transform from five places.  Round 1 (scalar FP32, per-trajectory masks and selects in the
// line builders) the inlined body was > 64 KB and the warps stalled on instruction fetch, so the transform was a real
// function (EXB_1D_FFT_CALL=1: the line travels by value in registers) and a CTA barrier per stage kept the warps on
// the same instruction-cache lines (EXB_1D_CTA_SYNC=1).  With packed FP32 and the lean line builders the inlined
// step is small enough: measured on B200 (c2, profiles/r02v_c2_lean_1d.md) call + barrier 1.51e11, inlined + barrier
// 1.62e11, inlined without the barrier 1.70e11 grid-point*steps/s -- the call's argument marshalling (about 60 MOVs
// between two transforms) and the barrier's phase-locking of the warps cost more than the instruction fetch.
#ifndef EXB_1D_FFT_CALL
#define EXB_1D_FFT_CALL 0
#endif
#ifndef EXB_1D_FFT_UNIFY
#define EXB_1D_FFT_UNIFY 1
#endif
// The line crosses the call as R 64-bit values (one register pair per complex point): passed as 2R scalar floats,
// every pair is re-formed with two MOVs in front of its first packed instruction (32 of the body's 293 instructions).
template <int R> struct RegLine { unsigned long long q[R]; };
__device__ __forceinline__ unsigned long long pair_of(cpx<float> a) {
  unsigned long long q;
  asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(a.x), "f"(a.y));
  return q;
}
__device__ __forceinline__ cpx<float> cpx_of(unsigned long long q) {
  cpx<float> a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(q));
  return a;
}
template <int R, int DIR>
__device__ __noinline__ RegLine<R> fft_reg_call(RegLine<R> a, cpx<float>* xb, int j, const cpx<float>* tw2) {
  cpx<float> v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = cpx_of(a.q[r]);
  fft_reg_body<R, DIR>(v, xb, j, tw2);
#pragma unroll
  for (int r = 0; r < R; ++r) a.q[r] = pair_of(v[r]);
  return a;
}
template <int R, int DIR>
__device__ __forceinline__ void fft_reg(cpx<float> (&v)[R], cpx<float>* xb, int j, const cpx<float>* tw2) {
#if EXB_1D_FFT_CALL
  // EXB_1D_FFT_UNIFY: the inverse transform is conj(forward(conj(x))) -- sign flips are exact, so one
  // function body serves both directions
  constexpr bool CONJ = (EXB_1D_FFT_UNIFY != 0) && DIR > 0;
  RegLine<R> a;
#pragma unroll
  for (int r = 0; r < R; ++r) a.q[r] = pair_of(CONJ ? cpx<float>(v[r].x, -v[r].y) : v[r]);
  a = fft_reg_call<R, CONJ ? -1 : DIR>(a, xb, j, tw2);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const cpx<float> t = cpx_of(a.q[r]);
    v[r] = CONJ ? cpx<float>(t.x, -t.y) : t;
  }
#else
  fft_reg_body<R, DIR>(v, xb, j, tw2);
#endif
}

template <int R, class S, int NINV, int NFWD> struct Fast1d {
  static constexpr int N = R * R;
  static constexpr int Nh = N / 2 + 1;
  static constexpr int NOWN = R / 2 + 1;  // modes owned per thread: k = j + R*r (r < R/2) [+ N/2 for j == 0]

  const K1dParams<float>& p;
  const FastLayout& lay;
  const cpx<float>* tw2;
  cpx<float>* st;   // pair state base: [array][traj][nhp]
  cpx<float>* xb;   // exchange buffer
  int j;            // thread index within the pair group
  const cpx<float>* sE;
  const cpx<float>* sEh;
  const float* sc[6];
  const float2* sMK;
  const float* sH;

  __device__ Fast1d(const K1dParams<float>& p_, const FastLayout& l_) : p(p_), lay(l_) {}

  __device__ __forceinline__ cpx<float>* state(int a) const { return st + (size_t)a * 2 * lay.nhp; }

  // packed line element n from the half-complex fields F1, F2 of the two trajectories
  template <int KINDN>  // 0: k = n < N/2 (direct), 1: DC or Nyquist, 2: upper (conjugate of mode N - n)
  static __device__ __forceinline__ cpx<float> pack(cpx<float> F1, cpx<float> F2) {
    if (KINDN == 1) return cpx<float>(F1.x, F2.x);                 // irfft drops these imaginary parts
    if (KINDN == 0) return F1 + mul_i(F2);                         // F1 + i F2        (one packed add each:
    return conj(F1) + mul_i(conj(F2));                             // conj(F1) + i conj(F2)   swaps / signs are operand modifiers)
  }

  // Build the packed inverse lines from spectral state `src`.  NLF: apply the nonlinear
  // function's prologue (mask, i*k, ...); otherwise the plain state (output transform).
  template <int NL, bool NLF>
  __device__ __forceinline__ void build_lines(const cpx<float>* src, cpx<float> (&z)[NL][R]) const {
    const NlParams<float>& P = p.P;
    auto elem = [&](int r, int k, int kindn) {
      cpx<float> u1[EXB_MAXC], u2[EXB_MAXC];
      u1[0] = src[k];
      u2[0] = src[lay.nhp + k];
      u1[1] = u1[2] = u2[1] = u2[2] = cpx<float>(0.f, 0.f);
      ModeK<float> m = make_mode<float, S>(P, k, 0, 0);
#pragma unroll
      for (int f = 0; f < NL; ++f) {
        cpx<float> F1 = NLF ? nl_inv_field<float, S>(P, f, u1, m) : u1[0];
        cpx<float> F2 = NLF ? nl_inv_field<float, S>(P, f, u2, m) : u2[0];
        z[f][r] = kindn == 0 ? pack<0>(F1, F2) : (kindn == 1 ? pack<1>(F1, F2) : pack<2>(F1, F2));
      }
    };
    // r = 0: n = j (DC for j == 0)
    elem(0, j, j == 0 ? 1 : 0);
#pragma unroll
    for (int r = 1; r < R / 2; ++r) elem(r, j + R * r, 0);
    // r = R/2: n = N/2 + j: Nyquist for j == 0, otherwise the conjugate of mode N/2 - j
    elem(R / 2, j == 0 ? N / 2 : N / 2 - j, j == 0 ? 1 : 2);
#pragma unroll
    for (int r = R / 2 + 1; r < R; ++r) elem(r, N - (j + R * r), 2);
  }

  // Inverse-transform inputs of the nonlinear function (C = 1, D = 1): every field is mask(k) * {1 | i kd} * u_hat, the
  // same factor for both trajectories of the pair, so the packed line is factor * (u1 + i u2) -- one packed add and one
  // packed multiply per point instead of per-trajectory masks, derivatives and selects.  The factor also carries the
  // 1/N of the inverse transform (N is a power of two here: scaling before or after the FFT is bit-identical).
  static __device__ __forceinline__ constexpr bool field_is_derivative(int f) {
    return (S::kind == EXB_NL_CONVECTION && !(S::var & 1) && f >= 1) || S::kind == EXB_NL_GRADIENT_NORM ||
           (S::kind == EXB_NL_GENERAL && f >= 1);
  }
  template <int NL>
  __device__ __forceinline__ void build_lines_nl(const cpx<float>* src, cpx<float> (&z)[NL][R]) const {
    // upper: n > N/2, the value is the conjugate of mode k = N - n.  edge: for j == 0 this slot is the DC / Nyquist
    // mode, of which irfft keeps the real part only
    auto elem = [&](int r, int k, bool upper, bool edge) {
      const cpx<float> u1 = src[k], u2 = src[lay.nhp + k];
      const float2 t = sMK[k];
      const cpx<float> z0 = upper ? conj(u1) + mul_i(conj(u2)) : u1 + mul_i(u2);
#pragma unroll
      for (int f = 0; f < NL; ++f) {
        cpx<float> v;
        if (field_is_derivative(f)) {
          const cpx<float> w = t.y * z0;                 // i kd w (lower), conj(i kd) w (upper)
          v = upper ? mul_mi(w) : mul_i(w);
          if (edge && j == 0) v = cpx<float>(-t.y * u1.y, -t.y * u2.y);   // Re(i kd u)
        } else {
          v = t.x * z0;
          if (edge && j == 0) v = cpx<float>(t.x * u1.x, t.x * u2.x);
        }
        z[f][r] = v;
      }
    };
    elem(0, j, false, true);
#pragma unroll
    for (int r = 1; r < R / 2; ++r) elem(r, j + R * r, false, false);
    elem(R / 2, j == 0 ? N / 2 : N / 2 - j, true, true);
#pragma unroll
    for (int r = R / 2 + 1; r < R; ++r) elem(r, N - (j + R * r), true, false);
  }

  // Z[N - k] for the owned mode k = j + R*r: it lives in the registers of thread (R - j) % R of the same group, slot
  // R - 1 - r (for j == 0 in the thread's own slot R - r), so the two-for-one split needs no shared-memory round trip:
  // two shuffles per mode instead of a 64-bit store + load (the shared-memory pipe is this kernel's busiest unit; the
  // staggered j == 0 row also cost a 2-way bank conflict on every partner load).
  __device__ __forceinline__ cpx<float> partner_of(const cpx<float> (&v)[R], int r) const {
#if EXB_1D_SHFL_UNPACK
    const int src = ((threadIdx.x & 31) & ~(R - 1)) | ((R - j) & (R - 1));
    const cpx<float> send = v[R - 1 - r];
    cpx<float> t;
    t.x = __shfl_sync(0xffffffffu, send.x, src);
    t.y = __shfl_sync(0xffffffffu, send.y, src);
    const cpx<float> own = r == 0 ? v[0] : v[R - r];
    return j == 0 ? own : t;
#else
    const int pidx = (j == 0) ? (R + 1) * (R - r) : (R - j) + (R + 1) * (R - 1 - r);
    return (r == 0 && j == 0) ? v[r] : xb[pidx];
#endif
  }
  __device__ __forceinline__ void stage_partners(const cpx<float> (&v)[R]) const {
#if !EXB_1D_SHFL_UNPACK
    __syncwarp();
#pragma unroll
    for (int r = R / 2; r < R; ++r) xb[j + (R + 1) * r] = v[r];  // upper half: partners of the owned modes
    __syncwarp();
#endif
  }

  // Two-for-one split of a forward-transformed line held in registers: for every owned mode slot
  // (k = j + R*slot, slot < R/2; slot R/2 = Nyquist, valid for j == 0 only) X1, X2.
  __device__ __forceinline__ void unpack_owned(const cpx<float> (&v)[R], cpx<float> (&X1)[NOWN],
                                               cpx<float> (&X2)[NOWN]) const {
    stage_partners(v);
#pragma unroll
    for (int r = 0; r < R / 2; ++r) {
      // partner of k = j + R*r is n' = N - k = (R - j) + R*(R - 1 - r)  (j > 0);  N - R*r = R*(R - r) (j == 0)
      cpx<float> zk = v[r];
      cpx<float> zp = partner_of(v, r);
      X1[r] = 0.5f * (zk + conj(zp));              // ( zk + conj(zp)) / 2
      X2[r] = mul_mi(0.5f * (zk - conj(zp)));      // (zk - conj(zp)) / (2 i)
    }
    X1[R / 2] = cpx<float>(v[R / 2].x, 0.f);  // Nyquist (meaningful for j == 0)
    X2[R / 2] = cpx<float>(v[R / 2].y, 0.f);
  }

  // Nonlinear functions whose post-processing is ONE real factor per mode, c(k) = mask(k) * {-scale | 1 | -scale / 2}:
  // the factor rides on the 1/2 of the two-for-one split (h = c / 2 is exact, the product rounds exactly as
  // c * (sum / 2) did), so N(u) of both trajectories is two packed multiplies per mode on top of the split.
  static constexpr bool kRealPost =
      NFWD == 1 && ((S::kind == EXB_NL_CONVECTION && S::var >= 0 && !(S::var & 1)) || S::kind == EXB_NL_POLYNOMIAL ||
                    S::kind == EXB_NL_GRADIENT_NORM);
  static __device__ __forceinline__ float post_factor_half(const NlParams<float>& P, int k) {
    if (P.kmax >= 0 && k > P.kmax) return 0.f;
    if (S::kind == EXB_NL_POLYNOMIAL) return 0.5f;
    if (S::kind == EXB_NL_GRADIENT_NORM) return (P.zero_mode_fix && k == 0) ? 0.f : -0.25f * P.scale;
    return -0.5f * P.scale;
  }
  __device__ __forceinline__ void unpack_scaled(const cpx<float> (&v)[R], cpx<float> (&n1)[NOWN],
                                                cpx<float> (&n2)[NOWN]) const {
    stage_partners(v);
#pragma unroll
    for (int r = 0; r < R / 2; ++r) {
      const float h = sH[j + R * r];
      cpx<float> zk = v[r];
      cpx<float> zp = partner_of(v, r);
      n1[r] = h * (zk + conj(zp));
      n2[r] = mul_mi(h * (zk - conj(zp)));
    }
    const float c = 2.f * sH[N / 2];
    n1[R / 2] = c * cpx<float>(v[R / 2].x, 0.f);  // Nyquist (meaningful for j == 0)
    n2[R / 2] = c * cpx<float>(v[R / 2].y, 0.f);
  }

  // N(src) -> per owned mode, both trajectories
  __device__ __forceinline__ void eval_nl(const cpx<float>* src, cpx<float> (&n1)[NOWN], cpx<float> (&n2)[NOWN]) const {
    const NlParams<float>& P = p.P;
    cpx<float> w[NFWD][R];
    {
      cpx<float> z[NINV][R];
      build_lines_nl<NINV>(src, z);
#pragma unroll
      for (int f = 0; f < NINV; ++f) fft_reg<R, +1>(z[f], xb, j, tw2);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        // the two lanes of a packed line value are the same grid point of the two trajectories
        f32x2 iv[NINV], ov[NFWD];
#pragma unroll
        for (int f = 0; f < NINV; ++f) iv[f] = lanes(z[f][r]);   // 1/N is in the line factors
        nl_pointwise<float, S, f32x2>(P, iv, ov);
#pragma unroll
        for (int g = 0; g < NFWD; ++g) w[g][r] = as_cpx(ov[g]);
      }
    }
    if constexpr (kRealPost) {
      fft_reg<R, -1>(w[0], xb, j, tw2);
      unpack_scaled(w[0], n1, n2);
      return;
    }
    cpx<float> W1[NFWD][NOWN], W2[NFWD][NOWN];
#pragma unroll
    for (int g = 0; g < NFWD; ++g) {
      fft_reg<R, -1>(w[g], xb, j, tw2);
      unpack_owned(w[g], W1[g], W2[g]);
    }
#pragma unroll
    for (int sl = 0; sl < NOWN; ++sl) {
      const int k = sl < R / 2 ? j + R * sl : N / 2;
      ModeK<float> m = make_mode<float, S>(P, k, 0, 0);
      cpx<float> wa[NFWD], wb[NFWD], a[EXB_MAXC], b[EXB_MAXC];
#pragma unroll
      for (int g = 0; g < NFWD; ++g) {
        wa[g] = W1[g][sl];
        wb[g] = W2[g][sl];
      }
      nl_from_fwd<float, S>(P, wa, m, a);
      nl_from_fwd<float, S>(P, wb, m, b);
      n1[sl] = a[0];
      n2[sl] = b[0];
    }
  }

  // one ETDRK step on the shared-memory state (array 0); stage formulas: exponax/etdrk/_etdrk_{0..4}.py
  __device__ __forceinline__ void etdrk_step(int order) {
    cpx<float>* U = state(0);
    const int nhp = lay.nhp;
    if (order == 0) {
#pragma unroll
      for (int sl = 0; sl < NOWN; ++sl) {
        const int k = sl < R / 2 ? j + R * sl : N / 2;
        if (sl < R / 2 || j == 0) {
          cpx<float> e = sE[k];
          U[k] = e * U[k];
          U[nhp + k] = e * U[nhp + k];
        }
      }
      __syncwarp();
      return;
    }
    for (int s = 0; s < order; ++s) {
      // order 2 runs in place (a overwrites u): 2 state arrays instead of 3
      const int si = order == 2 ? -1 : etdrk_stage_input(order, s);
      const cpx<float>* src = si < 0 ? U : state(1 + si);
      cpx<float> n1[NOWN], n2[NOWN];
      eval_nl(src, n1, n2);
      cpx<float>* S0 = state(lay.nstate > 1 ? 1 : 0);
      cpx<float>* S1 = state(lay.nstate > 2 ? 2 : 0);
      cpx<float>* S2 = state(lay.nstate > 3 ? 3 : 0);
      cpx<float>* S3 = state(lay.nstate > 4 ? 4 : 0);
      // One (order, stage) dispatch per stage, not per mode; per owned mode the coefficients are read once for both
      // trajectories and every load precedes the stores (the compiler cannot reorder them across the aliasing stores).
      auto for_owned = [&](auto&& fn) {
#pragma unroll
        for (int sl = 0; sl < NOWN; ++sl) {
          const int k = sl < R / 2 ? j + R * sl : N / 2;
          if (sl < R / 2 || j == 0) fn(k, nhp + k, n1[sl], n2[sl]);
        }
      };
      switch (order * 4 + s) {
        case 4:   // ETDRK1 (_etdrk_1.py:78-82)
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> e = sE[a];
            const float c0 = sc[0][a];
            const cpx<float> ua = U[a], ub = U[b];
            U[a] = e * ua + c0 * na;
            U[b] = e * ub + c0 * nb;
          });
          break;
        case 8:   // ETDRK2 (_etdrk_2.py:91-102), a overwrites u
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> e = sE[a];
            const float c0 = sc[0][a];
            const cpx<float> ua = U[a], ub = U[b];
            U[a] = e * ua + c0 * na;
            S0[a] = na;
            U[b] = e * ub + c0 * nb;
            S0[b] = nb;
          });
          break;
        case 9:
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const float c1 = sc[1][a];
            const cpx<float> ua = U[a], ub = U[b], pa = S0[a], pb = S0[b];
            U[a] = ua + c1 * (na - pa);
            U[b] = ub + c1 * (nb - pb);
          });
          break;
        case 12:  // ETDRK3 (_etdrk_3.py:191-212)
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> eh = sEh[a];
            const float c0 = sc[0][a];
            const cpx<float> ua = U[a], ub = U[b];
            S0[a] = eh * ua + c0 * na;
            S1[a] = na;
            S0[b] = eh * ub + c0 * nb;
            S1[b] = nb;
          });
          break;
        case 13:
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> e = sE[a];
            const float c1 = sc[1][a];
            const cpx<float> ua = U[a], ub = U[b], pa = S1[a], pb = S1[b];
            S0[a] = e * ua + c1 * (2.f * na - pa);
            S2[a] = na;
            S0[b] = e * ub + c1 * (2.f * nb - pb);
            S2[b] = nb;
          });
          break;
        case 14:
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> e = sE[a];
            const float c2 = sc[2][a], c3 = sc[3][a], c4 = sc[4][a];
            const cpx<float> ua = U[a], ub = U[b], pa = S1[a], pb = S1[b], qa = S2[a], qb = S2[b];
            U[a] = e * ua + c2 * pa + c3 * qa + c4 * na;
            U[b] = e * ub + c2 * pb + c3 * qb + c4 * nb;
          });
          break;
        case 16:  // ETDRK4 (_etdrk_4.py:198-224)
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> eh = sEh[a];
            const float c0 = sc[0][a];
            const cpx<float> ua = U[a], ub = U[b];
            S0[a] = eh * ua + c0 * na;
            S1[a] = na;
            S0[b] = eh * ub + c0 * nb;
            S1[b] = nb;
          });
          break;
        case 17:
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> eh = sEh[a];
            const float c1 = sc[1][a];
            const cpx<float> ua = U[a], ub = U[b];
            S2[a] = eh * ua + c1 * na;
            S3[a] = na;
            S2[b] = eh * ub + c1 * nb;
            S3[b] = nb;
          });
          break;
        case 18:
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> eh = sEh[a];
            const float c2 = sc[2][a];
            const cpx<float> aa = S0[a], ab = S0[b], pa = S1[a], pb = S1[b], qa = S3[a], qb = S3[b];
            S2[a] = eh * aa + c2 * (2.f * na - pa);
            S3[a] = qa + na;
            S2[b] = eh * ab + c2 * (2.f * nb - pb);
            S3[b] = qb + nb;
          });
          break;
        default:  // 19
          for_owned([&](int a, int b, cpx<float> na, cpx<float> nb) {
            const cpx<float> e = sE[a];
            const float c3 = sc[3][a], c4 = sc[4][a], c5 = sc[5][a];
            const cpx<float> ua = U[a], ub = U[b], pa = S1[a], pb = S1[b], qa = S3[a], qb = S3[b];
            U[a] = e * ua + c3 * pa + c4 * (2.f * qa) + c5 * na;
            U[b] = e * ub + c3 * pb + c4 * (2.f * qb) + c5 * nb;
          });
          break;
      }
      __syncwarp();
#if EXB_1D_CTA_SYNC
      // No data is shared between warps: the barrier only keeps the warps of the CTA at the same place of
      // the (fully unrolled, > 64 KB) step body, so that they share instruction-cache lines instead of
      // each streaming the body from L2 on its own (ncu r01: `no_instruction` was the top stall reason).
#if EXB_1D_SPLIT_BAR
      // two half-CTA barriers (even / odd warps) instead of one: the halves drift apart, so that the FMA-pipe phases
      // (butterflies) of one half overlap the shared-memory phases (exchanges, state updates) of the other
      {
        const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const int half = warp & 1, cnt = ((nw + 1 - half) >> 1) * 32;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(cnt) : "memory");
      }
#else
      __syncthreads();
#endif
#if EXB_1D_SKEW_NS
      // de-phase the odd warps by a fraction of a transform: the butterflies of one half then overlap the exchanges
      // of the other instead of all warps saturating the FMA pipe and the shared-memory pipe in turns
      if ((threadIdx.x >> 5) & 1) __nanosleep(EXB_1D_SKEW_NS);
#endif
#endif
    }
  }
};

template <int R, class S, int NINV, int NFWD>
__global__ void __launch_bounds__(256, 2) k1d_fast_kernel(const K1dParams<float> p, const FastLayout lay) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = R * R, Nh = N / 2 + 1;
  constexpr int GROUPS_PER_WARP = 32 / R;
  using F1 = Fast1d<R, S, NINV, NFWD>;
  constexpr int NOWN = F1::NOWN;
  // ---- cooperative load of the tables ----
  {
    cpx<float>* tw2 = (cpx<float>*)(smem_raw + lay.off_tw2);
    for (int q = threadIdx.x; q < R * R; q += blockDim.x) {
      int r = q / R, jj = q - r * R;
      tw2[q] = p.tw[(r * jj) % N];
    }
    cpx<float>* e = (cpx<float>*)(smem_raw + lay.off_exp);
    cpx<float>* eh = (cpx<float>*)(smem_raw + lay.off_hexp);
    for (int q = threadIdx.x; q < Nh; q += blockDim.x) {
      e[q] = p.K.exp_term[q];
      if (p.K.half_exp) eh[q] = p.K.half_exp[q];
      for (int i = 0; i < 6; ++i)
        if (p.K.c[i]) ((float*)(smem_raw + lay.off_c[i]))[q] = p.K.c[i][q];
      const bool keep = !(p.P.kmax >= 0 && q > p.P.kmax);
      ((float*)(smem_raw + lay.off_h))[q] = F1::post_factor_half(p.P, q);
      ((float2*)(smem_raw + lay.off_mk))[q] =
          keep ? make_float2(p.P.inv_norm, (p.P.dscale * (float)q) * p.P.inv_norm) : make_float2(0.f, 0.f);
    }
  }
  __syncthreads();

  F1 F(p, lay);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = (blockIdx.x * (blockDim.x >> 5) + warp) * GROUPS_PER_WARP + lane / R;
  const int local_group = warp * GROUPS_PER_WARP + lane / R;
  const int j = lane % R;
  F.j = j;
  F.tw2 = (const cpx<float>*)(smem_raw + lay.off_tw2);
  unsigned char* pb = smem_raw + lay.off_pairs + (size_t)local_group * lay.pair_bytes;
  F.st = (cpx<float>*)pb;
  F.xb = (cpx<float>*)(pb + (size_t)lay.nstate * 2 * lay.nhp * sizeof(cpx<float>));
  F.sE = (const cpx<float>*)(smem_raw + lay.off_exp);
  F.sEh = (const cpx<float>*)(smem_raw + lay.off_hexp);
  for (int i = 0; i < 6; ++i) F.sc[i] = (const float*)(smem_raw + lay.off_c[i]);
  F.sMK = (const float2*)(smem_raw + lay.off_mk);
  F.sH = (const float*)(smem_raw + lay.off_h);
  const int order = p.K.order;

  const long long t1 = 2ll * group, t2 = t1 + 1;
  const bool act1 = t1 < p.batch, act2 = t2 < p.batch;
  const bool include_init = (p.flags & EXB_ROLLOUT_INCLUDE_INIT) != 0;
  const bool layout_tb = (p.flags & EXB_ROLLOUT_LAYOUT_TB) != 0;
  const bool final_only = (p.flags & EXB_ROLLOUT_FINAL_ONLY) != 0;
  const bool spectral_carry = (p.flags & EXB_ROLLOUT_SPECTRAL_CARRY) != 0;
  const long long Tn = final_only ? 1 : p.n_saved + (include_init ? 1 : 0);
  float* out = (float*)p.out;
  // output pointers of slot 0 and the slot stride (elements)
  float* o1 = nullptr;
  float* o2 = nullptr;
  size_t slot_stride = 0;
  if (final_only) {
    o1 = out + (size_t)t1 * N;
    o2 = out + (size_t)t2 * N;
  } else if (layout_tb) {
    o1 = out + (size_t)t1 * N;
    o2 = out + (size_t)t2 * N;
    slot_stride = (size_t)p.batch * N;
  } else {
    o1 = out + (size_t)t1 * Tn * N;
    o2 = out + (size_t)t2 * Tn * N;
    slot_stride = N;
  }
  o1 += j;
  o2 += j;
  auto store_phys = [&](const cpx<float> (&v)[R], long long slot, float scale) {
    float* a = o1 + (size_t)slot * slot_stride;
    float* b = o2 + (size_t)slot * slot_stride;
    if (act1) {
#pragma unroll
      for (int r = 0; r < R; ++r) a[R * r] = v[r].x * scale;
    }
    if (act2) {
#pragma unroll
      for (int r = 0; r < R; ++r) b[R * r] = v[r].y * scale;
    }
  };

  cpx<float>* U = F.state(0);
  auto to_state = [&](cpx<float> (&line)[R]) {
    fft_reg<R, -1>(line, F.xb, j, F.tw2);
    cpx<float> X1[NOWN], X2[NOWN];
    F.unpack_owned(line, X1, X2);
#pragma unroll
    for (int sl = 0; sl < R / 2; ++sl) {
      U[j + R * sl] = X1[sl];
      U[lay.nhp + j + R * sl] = X2[sl];
    }
    if (j == 0) {
      U[N / 2] = X1[R / 2];
      U[lay.nhp + N / 2] = X2[R / 2];
    }
    __syncwarp();
  };

  // ---- u0 -> spectral state ----
  {
    cpx<float> v[R];
    const float* in = (const float*)p.in;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float a = act1 ? in[(size_t)t1 * N + j + R * r] : 0.f;
      float b = act2 ? in[(size_t)t2 * N + j + R * r] : 0.f;
      v[r] = cpx<float>(a, b);
    }
    if (include_init && !final_only) store_phys(v, 0, 1.0f);
    to_state(v);
  }

  const float invN = p.P.inv_norm;
  for (long long s = 0; s < p.n_saved; ++s) {
    if (p.forcing) {  // u_hat += dt * f_hat (ForcedStepper), owned modes of both trajectories
      const cpx<float>* f1 = p.forcing + s * p.fstep + t1 * p.fbatch;
      const cpx<float>* f2 = p.forcing + s * p.fstep + t2 * p.fbatch;
#pragma unroll
      for (int sl = 0; sl < R / 2; ++sl) {
        const int k = j + R * sl;
        if (act1) U[k] = axpy(p.fscale, f1[k], U[k]);
        if (act2) U[lay.nhp + k] = axpy(p.fscale, f2[k], U[lay.nhp + k]);
      }
      if (j == 0) {
        if (act1) U[N / 2] = axpy(p.fscale, f1[N / 2], U[N / 2]);
        if (act2) U[lay.nhp + N / 2] = axpy(p.fscale, f2[N / 2], U[lay.nhp + N / 2]);
      }
      __syncwarp();
    }
    for (int sub = 0; sub < p.substeps; ++sub) F.etdrk_step(order);
    const bool last = (s == p.n_saved - 1);
    const bool store = !final_only || last;
    if (store || !spectral_carry) {
      cpx<float> z[1][R];
      F.template build_lines<1, false>(U, z);
      fft_reg<R, +1>(z[0], F.xb, j, F.tw2);
#pragma unroll
      for (int r = 0; r < R; ++r) z[0][r] = invN * z[0][r];    // one packed scaling serves the snapshot and the carry
      if (store) store_phys(z[0], final_only ? 0 : s + (include_init ? 1 : 0), 1.0f);
      if (last) break;
      if (!spectral_carry) {
        to_state(z[0]);
        continue;
      }
    }
    if (last) break;
    // spectral carry: the reference's irfft -> rfft round trip == Hermitian projection
    if (j == 0) {
      U[0].y = 0.f;
      U[N / 2].y = 0.f;
      U[lay.nhp].y = 0.f;
      U[lay.nhp + N / 2].y = 0.f;
    }
    __syncwarp();
  }
}

}  // namespace exb
