// Fast 1-D persistent rollout kernel for N = R*R (R = 16 -> N = 256, R = 8 -> N = 64), f32, C = 1.
//
// Work decomposition (B200: 148 SMs, 64K regs, 227 KB smem / SM):
//   * a PAIR of trajectories is carried as ONE complex trajectory z = x1 + i*x2: the state is the packed spectrum
//     Z[n] = X1[n] + i X2[n] over all N modes (see Fast1d below) -- no two-for-one split or packing anywhere;
//   * a pair is owned by R threads (a half-warp for N = 256) for the whole rollout; thread j owns the modes / points
//     n = j + R*r, so every state access is thread-private and the only synchronisation is the __syncwarp pair around
//     the transform's exchange (no __syncthreads after the table load);
//   * each thread keeps R points of the line in registers; an N-point FFT is two in-register radix-R passes with ONE
//     shared-memory exchange in between (padded, conflict-free); 11 of the 15 inter-pass twiddles are derived;
//   * the pointwise nonlinearity happens in registers between the inverse and forward transforms (the output
//     mapping of the inverse FFT is the input mapping of the forward);
//   * packed spectral state, ETDRK stage buffers, coefficient and factor tables live in shared memory; HBM sees the
//     initial condition once and the saved snapshots;
//   * the nonlinear function is a compile-time descriptor S (NlS<...>): no per-element branching; ETDRK2 (the
//     reference's default order) has its own kernel instance (ORD = 2).
// How this structure was arrived at, step by step with measurements: profiles/r02v_c2_lean_1d.md, DESIGN.md section 6.
// Reference semantics identical to k1d_kernel (exb_kernels_1d.cuh); parity is tested against the same oracle.
#pragma once
#include "exb_kernels_1d.cuh"

namespace exb {

struct FastLayout {
  int off_tw2;       // [R][R] inter-pass twiddles w_N^(j*r), forward sign
  int off_exp;       // cpx[Nh]
  int off_hexp;      // cpx[Nh]
  int off_c[6];      // float[Nh] each
  int off_mk;        // float2[Nh]: (keep / N, keep * kd / N) -- pre-dealiasing mask, derivative scale and 1/N of the inverse transform
  int off_pairs;     // start of per-pair storage
  int pair_bytes;    // bytes per pair
  int nstate;        // spectral state arrays per pair (1 + scratch), each N + 2 complex values (Fast1d::ZS)
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

#ifndef EXB_1D_CTA_SYNC
#define EXB_1D_CTA_SYNC 0
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

// NL independent lines through one transform, phase by phase: all first passes, then the exchanges (one buffer, one
// line after the other), then all second passes.  The lines share the inter-pass twiddles (loaded / derived once), and
// the scheduler may overlap the exchange latency of one line with the butterflies of the other -- called line by line,
// the __syncwarp()s around each exchange keep the transforms strictly one after the other.
#ifndef EXB_1D_FFT_LINES
#define EXB_1D_FFT_LINES 1
#endif
template <int R, int DIR, int NL>
__device__ __forceinline__ void fft_reg_lines(cpx<float> (&v)[NL][R], cpx<float>* xb, int j, const cpx<float>* tw2) {
#if EXB_1D_FFT_LINES && !EXB_1D_FFT_CALL
#pragma unroll
  for (int f = 0; f < NL; ++f) dft_reg<R, DIR>(v[f]);
  cpx<float> w[R];
#if EXB_1D_DERIVE_TW
#pragma unroll
  for (int b = 1; b < R; b <<= 1) w[b] = twd<float, DIR>(tw2[b * R + j]);
#pragma unroll
  for (int r = 3; r < R; ++r)
    if (r & (r - 1)) w[r] = w[r & (r - 1)] * w[r & -r];
#else
#pragma unroll
  for (int r = 1; r < R; ++r) w[r] = twd<float, DIR>(tw2[r * R + j]);
#endif
#pragma unroll
  for (int f = 0; f < NL; ++f) {
    __syncwarp();
#pragma unroll
    for (int p = 0; p < R; ++p) xb[(R + 1) * j + xidx<R>(p)] = v[f][p];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cpx<float> t = xb[j + (R + 1) * r];
      v[f][r] = r > 0 ? t * w[r] : t;
    }
  }
#pragma unroll
  for (int f = 0; f < NL; ++f) {
    dft_reg<R, DIR>(v[f]);
    if (R == 16) {
      cpx<float> o[R];
#pragma unroll
      for (int p = 0; p < R; ++p) o[xidx<R>(p)] = v[f][p];
#pragma unroll
      for (int p = 0; p < R; ++p) v[f][p] = o[p];
    }
  }
#else
#pragma unroll
  for (int f = 0; f < NL; ++f) fft_reg<R, DIR>(v[f], xb, j, tw2);
#endif
}

// ETDRK2 instance: the two stages as a 2-iteration loop around ONE copy of the nonlinear evaluation (three transforms)
// instead of straight-line code with two copies: 3048 instead of 3944 instructions, and c2 2.06e11 -> 2.32e11 -- the step
// body has to stay inside the instruction cache (profiles/r02v_c2_lean_1d.md)
#ifndef EXB_1D_O2_LOOP
#define EXB_1D_O2_LOOP 1
#endif
#ifndef EXB_1D_UPD_BATCH
#define EXB_1D_UPD_BATCH 8
#endif
// what one slot of an ETDRK stage update reads: coefficients and up to three state / stage values (unused members vanish)
struct SlotLd {
  cpx<float> e, u, p, q;
  float f0, f1, f2;
};

// ---- the pair as ONE complex trajectory -------------------------------------------------------------------------
// z = x1 + i x2 has the spectrum Z[n] = X1[n] + i X2[n], n = 0 .. N-1, and because X1, X2 are spectra of real signals,
// Z[N - k] = conj(X1[k]) + i conj(X2[k]).  Every operation of the step is complex-linear per mode with coefficients
// that satisfy c(-k) = conj(c(k)) (real PDE operators: polynomials in i k with real coefficients), so the ETDRK update,
// the derivative / mask factors and the post-processing of the nonlinear function act on Z[n] directly:
//     Z+[k]     = e(k) Z[k]           + c0(k) f(k) W[k]          (n = k <= N/2)
//     Z+[N - k] = conj(e(k)) Z[N - k] + c0(k) conj(f(k)) W[N - k]
// where W is the forward transform of the packed pointwise product.  The two-for-one split (and its inverse, the
// packing of the inverse-transform inputs) never happens: the state IS the packed spectrum.  Thread j of the group
// owns n = j + R*r, r = 0 .. R-1 -- exactly the elements its registers hold after the forward transform and must hold
// before the inverse one, so every state access is thread-private (no partner exchange, no __syncwarp, no selects).
// The one exception is the Nyquist mode: within a step U1[N/2], U2[N/2] are genuinely complex (odd-derivative
// operators) while irfft keeps their real parts only, so they cannot share one complex slot; lane j == 0 carries them
// as two extra values behind Z (slots N, N+1) and the packed slot N/2 is overridden wherever a line is built.
// (The DC slot needs nothing: Im U[0] is exactly zero throughout -- e(0) and f(0) are real -- so Z[0] = (U1[0], U2[0]).)
template <int R, class S, int NINV, int NFWD> struct Fast1d {
  static constexpr int N = R * R;
  static constexpr int Nh = N / 2 + 1;
  static constexpr int ZS = N + 2;   // complex values per state array: Z[0 .. N-1], U1[N/2], U2[N/2]

  const K1dParams<float>& p;
  const FastLayout& lay;
  const cpx<float>* tw2;
  cpx<float>* st;   // pair state base: [array][ZS]
  cpx<float>* xb;   // exchange buffer
  int j;            // thread index within the pair group
  const cpx<float>* sE;
  const cpx<float>* sEh;
  const float* sc[6];
  const float2* sMK;

  __device__ Fast1d(const K1dParams<float>& p_, const FastLayout& l_) : p(p_), lay(l_) {}

  __device__ __forceinline__ cpx<float>* state(int a) const { return st + (size_t)a * ZS; }
  // slot r of this thread: line / state index n, index k = |wavenumber| into the coefficient tables, and whether the
  // slot is a negative wavenumber (coefficients conjugated)
  __device__ __forceinline__ int nidx(int r) const { return j + R * r; }
  __device__ __forceinline__ int kidx(int r) const { return r < R / 2 ? j + R * r : N - (j + R * r); }
  static __device__ __forceinline__ constexpr bool upper(int r) { return r >= R / 2; }
  // c(+-k) * u for a table coefficient c(k)
  static __device__ __forceinline__ cpx<float> cmul(cpx<float> c, cpx<float> u, bool up) {
    return up ? mul_conj(u, c) : c * u;
  }

  // inverse-transform input of the plain state (snapshots)
  __device__ __forceinline__ void build_plain(const cpx<float>* src, cpx<float> (&z)[R]) const {
#pragma unroll
    for (int r = 0; r < R; ++r) z[r] = src[nidx(r)];
    if (j == 0) z[R / 2] = cpx<float>(src[N].x, src[N + 1].x);       // irfft keeps Re U[N/2]
  }

  // Inverse-transform inputs of the nonlinear function (C = 1, D = 1): every field is mask(k) * {1 | i kd} * u_hat, so
  // the packed line is factor(+-k) * Z[n]: one shared-memory read of the state, one of the factor pair (which also
  // carries the 1/N of the inverse transform -- N is a power of two, scaling before or after is bit-identical) and one
  // packed multiply per field and point.
  static __device__ __forceinline__ constexpr bool field_is_derivative(int f) {
    return (S::kind == EXB_NL_CONVECTION && !(S::var & 1) && f >= 1) || S::kind == EXB_NL_GRADIENT_NORM ||
           (S::kind == EXB_NL_GENERAL && f >= 1);
  }
  // zr (REGS): the thread's slots of `src` are already in registers (the previous update / transform left them there)
  template <int NL, bool REGS = false>
  __device__ __forceinline__ void build_nl(const cpx<float>* src, cpx<float> (&z)[NL][R], const cpx<float>* zr = nullptr) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const cpx<float> z0 = REGS ? zr[r] : src[nidx(r)];
      const float2 t = sMK[kidx(r)];
#pragma unroll
      for (int f = 0; f < NL; ++f) {
        if (field_is_derivative(f)) {
          const cpx<float> w = t.y * z0;                   // (+-i kd) Z
          z[f][r] = upper(r) ? mul_mi(w) : mul_i(w);
        } else {
          z[f][r] = t.x * z0;
        }
      }
    }
    if (j == 0) {
      const cpx<float> u1 = src[N], u2 = src[N + 1];
      const float2 t = sMK[N / 2];
#pragma unroll
      for (int f = 0; f < NL; ++f)
        z[f][R / 2] = field_is_derivative(f) ? cpx<float>(-t.y * u1.y, -t.y * u2.y)   // Re(i kd U)
                                             : cpx<float>(t.x * u1.x, t.x * u2.x);
    }
  }

  // Nonlinear functions whose post-processing is ONE real factor per mode, c(k) = mask(k) * {-scale | 1 | -scale / 2}
  static constexpr bool kRealPost =
      NFWD == 1 && ((S::kind == EXB_NL_CONVECTION && S::var >= 0 && !(S::var & 1)) || S::kind == EXB_NL_POLYNOMIAL ||
                    S::kind == EXB_NL_GRADIENT_NORM);
  static __device__ __forceinline__ float post_factor(const NlParams<float>& P, int k) {
    if (P.kmax >= 0 && k > P.kmax) return 0.f;
    if (S::kind == EXB_NL_POLYNOMIAL) return 1.f;
    if (S::kind == EXB_NL_GRADIENT_NORM) return (P.zero_mode_fix && k == 0) ? 0.f : -0.5f * P.scale;
    return -P.scale;
  }

  // N(src): the packed nonlinear term at the thread's R slots, plus the two Nyquist values (meaningful for j == 0)
  template <bool REGS = false>
  __device__ __forceinline__ void eval_nl(const cpx<float>* src, cpx<float> (&npk)[R], cpx<float>& ny1,
                                          cpx<float>& ny2, const cpx<float>* zr = nullptr) const {
    const NlParams<float>& P = p.P;
    cpx<float> w[NFWD][R];
    {
      cpx<float> z[NINV][R];
      build_nl<NINV, REGS>(src, z, zr);
      fft_reg_lines<R, +1, NINV>(z, xb, j, tw2);
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
    fft_reg_lines<R, -1, NFWD>(w, xb, j, tw2);
    if constexpr (kRealPost) {
      // the real post-factor c(k) is folded into the ETDRK coefficient tables (every update term is c_i(k) * N, see the
      // table load in the kernel prologue): the transformed product is the nonlinear term as far as this kernel goes
#pragma unroll
      for (int r = 0; r < R; ++r) npk[r] = w[0][r];
      ny1 = cpx<float>(w[0][R / 2].x, 0.f);
      ny2 = cpx<float>(w[0][R / 2].y, 0.f);
    } else {
      // the post-processing (nl_from_fwd) is complex-linear in the transformed fields with coefficients that are
      // polynomials in i kd: evaluated at -kd for the negative-wavenumber slots it acts on the packed values as is
      auto post = [&](int k, bool up, bool nyq, int lane, int r, cpx<float>& out) {
        ModeK<float> m;
        m.kd[0] = (up ? -1.f : 1.f) * (P.dscale * (float)k);
        m.kd[1] = m.kd[2] = 0.f;
        m.keep = !(P.kmax >= 0 && k > P.kmax);
        m.is_inj = false;
        m.is_dc = k == 0;
        cpx<float> wa[NFWD], a[EXB_MAXC];
#pragma unroll
        for (int g = 0; g < NFWD; ++g)
          wa[g] = nyq ? cpx<float>(lane == 0 ? w[g][r].x : w[g][r].y, 0.f) : w[g][r];
        nl_from_fwd<float, S>(P, wa, m, a);
        out = a[0];
      };
#pragma unroll
      for (int r = 0; r < R; ++r) post(kidx(r), upper(r), false, 0, r, npk[r]);
      post(N / 2, false, true, 0, R / 2, ny1);
      post(N / 2, false, true, 1, R / 2, ny2);
    }
  }

  // ETDRK2 (_etdrk_2.py:91-102) with the thread's slots of the state handed over in registers from the update (or
  // transform) that produces them to the line builder that consumes them (zr: in = u, out = u+).  The state arrays in
  // shared memory are then read by the updates only: three line builds per step cost no shared-memory traffic.
  __device__ __forceinline__ void etdrk2_step(cpx<float> (&zr)[R]) {
    cpx<float>* __restrict__ U = state(0);
    cpx<float>* __restrict__ S0 = state(1);
    constexpr int B = EXB_1D_UPD_BATCH;
    cpx<float> npk[R], ny1, ny2;
#if EXB_1D_O2_LOOP
    // one copy of the nonlinear evaluation (three transforms) in the instruction stream instead of two
#pragma unroll 1
    for (int stg = 0; stg < 2; ++stg) {
      eval_nl<true>(U, npk, ny1, ny2, zr);
      if (stg == 0) {
    #pragma unroll
        for (int r0 = 0; r0 < R; r0 += B) {     // a = e u + c0 N(u), overwrites u; N(u) kept for stage 1
          cpx<float> e[B], u[B];
          float c0[B];
    #pragma unroll
          for (int q = 0; q < B; ++q) {
            e[q] = sE[kidx(r0 + q)];
            c0[q] = sc[0][kidx(r0 + q)];
            u[q] = U[nidx(r0 + q)];
          }
    #pragma unroll
          for (int q = 0; q < B; ++q) {
            const cpx<float> a = cmul(e[q], u[q], upper(r0 + q)) + c0[q] * npk[r0 + q];
            U[nidx(r0 + q)] = a;
            S0[nidx(r0 + q)] = npk[r0 + q];
            zr[r0 + q] = a;
          }
        }
        if (j == 0) {
          const cpx<float> e = sE[N / 2];
          const float c0 = sc[0][N / 2];
          const cpx<float> u1 = U[N], u2 = U[N + 1];
          U[N] = e * u1 + c0 * ny1;
          S0[N] = ny1;
          U[N + 1] = e * u2 + c0 * ny2;
          S0[N + 1] = ny2;
        }
      } else {
    #pragma unroll
        for (int r0 = 0; r0 < R; r0 += B) {     // u+ = a + c1 (N(a) - N(u))
          cpx<float> u[B], n0[B];
          float c1[B];
    #pragma unroll
          for (int q = 0; q < B; ++q) {
            c1[q] = sc[1][kidx(r0 + q)];
            u[q] = U[nidx(r0 + q)];
            n0[q] = S0[nidx(r0 + q)];
          }
    #pragma unroll
          for (int q = 0; q < B; ++q) {
            const cpx<float> un = u[q] + c1[q] * (npk[r0 + q] - n0[q]);
            U[nidx(r0 + q)] = un;
            zr[r0 + q] = un;
          }
        }
        if (j == 0) {
          const float c1 = sc[1][N / 2];
          const cpx<float> u1 = U[N], u2 = U[N + 1], p1 = S0[N], p2 = S0[N + 1];
          U[N] = u1 + c1 * (ny1 - p1);
          U[N + 1] = u2 + c1 * (ny2 - p2);
        }
      }
    }
#else
    eval_nl<true>(U, npk, ny1, ny2, zr);
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += B) {     // a = e u + c0 N(u), overwrites u; N(u) kept for stage 1
      cpx<float> e[B], u[B];
      float c0[B];
#pragma unroll
      for (int q = 0; q < B; ++q) {
        e[q] = sE[kidx(r0 + q)];
        c0[q] = sc[0][kidx(r0 + q)];
        u[q] = U[nidx(r0 + q)];
      }
#pragma unroll
      for (int q = 0; q < B; ++q) {
        const cpx<float> a = cmul(e[q], u[q], upper(r0 + q)) + c0[q] * npk[r0 + q];
        U[nidx(r0 + q)] = a;
        S0[nidx(r0 + q)] = npk[r0 + q];
        zr[r0 + q] = a;
      }
    }
    if (j == 0) {
      const cpx<float> e = sE[N / 2];
      const float c0 = sc[0][N / 2];
      const cpx<float> u1 = U[N], u2 = U[N + 1];
      U[N] = e * u1 + c0 * ny1;
      S0[N] = ny1;
      U[N + 1] = e * u2 + c0 * ny2;
      S0[N + 1] = ny2;
    }
    eval_nl<true>(U, npk, ny1, ny2, zr);
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += B) {     // u+ = a + c1 (N(a) - N(u))
      cpx<float> u[B], n0[B];
      float c1[B];
#pragma unroll
      for (int q = 0; q < B; ++q) {
        c1[q] = sc[1][kidx(r0 + q)];
        u[q] = U[nidx(r0 + q)];
        n0[q] = S0[nidx(r0 + q)];
      }
#pragma unroll
      for (int q = 0; q < B; ++q) {
        const cpx<float> un = u[q] + c1[q] * (npk[r0 + q] - n0[q]);
        U[nidx(r0 + q)] = un;
        zr[r0 + q] = un;
      }
    }
    if (j == 0) {
      const float c1 = sc[1][N / 2];
      const cpx<float> u1 = U[N], u2 = U[N + 1], p1 = S0[N], p2 = S0[N + 1];
      U[N] = u1 + c1 * (ny1 - p1);
      U[N + 1] = u2 + c1 * (ny2 - p2);
    }
#endif
  }

  // one ETDRK step on the shared-memory state (array 0); stage formulas: exponax/etdrk/_etdrk_{0..4}.py
  __device__ __forceinline__ void etdrk_step(int order) {
    cpx<float>* __restrict__ U = state(0);
    if (order == 0) {
      cpx<float> u[R], e[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        u[r] = U[nidx(r)];
        e[r] = sE[kidx(r)];
      }
#pragma unroll
      for (int r = 0; r < R; ++r) U[nidx(r)] = cmul(e[r], u[r], upper(r));
      if (j == 0) {
        const cpx<float> e = sE[N / 2];
        U[N] = e * U[N];
        U[N + 1] = e * U[N + 1];
      }
      return;
    }
    for (int s = 0; s < order; ++s) {
      // order 2 runs in place (a overwrites u): 2 state arrays instead of 3
      const int si = order == 2 ? -1 : etdrk_stage_input(order, s);
      const cpx<float>* src = si < 0 ? U : state(1 + si);
      cpx<float> npk[R], ny1, ny2;
      eval_nl(src, npk, ny1, ny2);
      cpx<float>* __restrict__ S0 = state(lay.nstate > 1 ? 1 : 0);
      cpx<float>* __restrict__ S1 = state(lay.nstate > 2 ? 2 : 0);
      cpx<float>* __restrict__ S2 = state(lay.nstate > 3 ? 3 : 0);
      cpx<float>* __restrict__ S3 = state(lay.nstate > 4 ? 4 : 0);
      // One (order, stage) dispatch per stage.  Each case is a load function ld(i, k) -> SlotLd (state slot i, table
      // index k) and an apply function ap(i, up, L, n) (up: conjugated coefficients, n: packed nonlinear term).  The
      // loads of EXB_1D_UPD_BATCH slots are issued ahead of their arithmetic and stores: the compiler cannot hoist a
      // shared-memory load over the (possibly aliasing) store of the previous slot by itself, and one slot at a time
      // exposes the full load latency 16 times per stage.  The Nyquist pair of lane 0 uses the same formulas.
      auto slots = [&](auto&& ld, auto&& ap) {
        constexpr int B = EXB_1D_UPD_BATCH;
#pragma unroll
        for (int r0 = 0; r0 < R; r0 += B) {
          SlotLd L[B];
#pragma unroll
          for (int q = 0; q < B; ++q) L[q] = ld(nidx(r0 + q), kidx(r0 + q));
#pragma unroll
          for (int q = 0; q < B; ++q) ap(nidx(r0 + q), upper(r0 + q), L[q], npk[r0 + q]);
        }
        if (j == 0) {
          const SlotLd a = ld(N, N / 2), b = ld(N + 1, N / 2);
          ap(N, false, a, ny1);
          ap(N + 1, false, b, ny2);
        }
      };
      switch (order * 4 + s) {
        case 4:   // ETDRK1 (_etdrk_1.py:78-82)
          slots([&](int i, int k) { SlotLd L; L.e = sE[k]; L.f0 = sc[0][k]; L.u = U[i]; return L; },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) { U[i] = cmul(L.e, L.u, up) + L.f0 * n; });
          break;
        case 8:   // ETDRK2 (_etdrk_2.py:91-102), a overwrites u
          slots([&](int i, int k) { SlotLd L; L.e = sE[k]; L.f0 = sc[0][k]; L.u = U[i]; return L; },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) {
                  U[i] = cmul(L.e, L.u, up) + L.f0 * n;
                  S0[i] = n;
                });
          break;
        case 9:
          slots([&](int i, int k) { SlotLd L; L.f0 = sc[1][k]; L.u = U[i]; L.p = S0[i]; return L; },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) { U[i] = L.u + L.f0 * (n - L.p); });
          break;
        case 12:  // ETDRK3 (_etdrk_3.py:191-212)
        case 16:  // ETDRK4 (_etdrk_4.py:198-224), first stage: the same formula
          slots([&](int i, int k) { SlotLd L; L.e = sEh[k]; L.f0 = sc[0][k]; L.u = U[i]; return L; },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) {
                  S0[i] = cmul(L.e, L.u, up) + L.f0 * n;
                  S1[i] = n;
                });
          break;
        case 13:
          slots([&](int i, int k) { SlotLd L; L.e = sE[k]; L.f0 = sc[1][k]; L.u = U[i]; L.p = S1[i]; return L; },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) {
                  S0[i] = cmul(L.e, L.u, up) + L.f0 * (2.f * n - L.p);
                  S2[i] = n;
                });
          break;
        case 14:
          slots([&](int i, int k) {
                  SlotLd L; L.e = sE[k]; L.f0 = sc[2][k]; L.f1 = sc[3][k]; L.f2 = sc[4][k];
                  L.u = U[i]; L.p = S1[i]; L.q = S2[i]; return L;
                },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) {
                  U[i] = cmul(L.e, L.u, up) + L.f0 * L.p + L.f1 * L.q + L.f2 * n;
                });
          break;
        case 17:
          slots([&](int i, int k) { SlotLd L; L.e = sEh[k]; L.f0 = sc[1][k]; L.u = U[i]; return L; },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) {
                  S2[i] = cmul(L.e, L.u, up) + L.f0 * n;
                  S3[i] = n;
                });
          break;
        case 18:
          slots([&](int i, int k) {
                  SlotLd L; L.e = sEh[k]; L.f0 = sc[2][k]; L.u = S0[i]; L.p = S1[i]; L.q = S3[i]; return L;
                },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) {
                  S2[i] = cmul(L.e, L.u, up) + L.f0 * (2.f * n - L.p);
                  S3[i] = L.q + n;
                });
          break;
        default:  // 19
          slots([&](int i, int k) {
                  SlotLd L; L.e = sE[k]; L.f0 = sc[3][k]; L.f1 = sc[4][k]; L.f2 = sc[5][k];
                  L.u = U[i]; L.p = S1[i]; L.q = S3[i]; return L;
                },
                [&](int i, bool up, const SlotLd& L, cpx<float> n) {
                  U[i] = cmul(L.e, L.u, up) + L.f0 * L.p + L.f1 * (2.f * L.q) + L.f2 * n;
                });
          break;
      }
#if EXB_1D_CTA_SYNC
      // No data is shared between warps: the barrier only keeps the warps of the CTA at the same place of the step
      // body so that they share instruction-cache lines (round 1: `no_instruction` was the top stall reason; with the
      // lean body it costs more than it saves -- off by default).
      __syncthreads();
#endif
    }
  }
};

// ORD = 2: the ETDRK2 instance (register hand-over between updates and line builders, no order dispatch); ORD = 0: any order
template <int R, class S, int NINV, int NFWD, int ORD>
__global__ void __launch_bounds__(256, 2) k1d_fast_kernel(const K1dParams<float> p, const FastLayout lay) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = R * R, Nh = N / 2 + 1;
  constexpr int GROUPS_PER_WARP = 32 / R;
  using F1 = Fast1d<R, S, NINV, NFWD>;
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
      const float post = F1::kRealPost ? F1::post_factor(p.P, q) : 1.f;
      for (int i = 0; i < 6; ++i)
        if (p.K.c[i]) ((float*)(smem_raw + lay.off_c[i]))[q] = p.K.c[i][q] * post;
      const bool keep = !(p.P.kmax >= 0 && q > p.P.kmax);
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
  F.xb = (cpx<float>*)(pb + (size_t)lay.nstate * F1::ZS * sizeof(cpx<float>));
  F.sE = (const cpx<float>*)(smem_raw + lay.off_exp);
  F.sEh = (const cpx<float>*)(smem_raw + lay.off_hexp);
  for (int i = 0; i < 6; ++i) F.sc[i] = (const float*)(smem_raw + lay.off_c[i]);
  F.sMK = (const float2*)(smem_raw + lay.off_mk);
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
  auto store_phys = [&](const cpx<float> (&v)[R], long long slot) {
    float* a = o1 + (size_t)slot * slot_stride;
    float* b = o2 + (size_t)slot * slot_stride;
    if (act1) {
#pragma unroll
      for (int r = 0; r < R; ++r) a[R * r] = v[r].x;
    }
    if (act2) {
#pragma unroll
      for (int r = 0; r < R; ++r) b[R * r] = v[r].y;
    }
  };

  cpx<float>* U = F.state(0);
  cpx<float> zr[R];   // ORD == 2: the thread's slots of U, valid whenever U is
  // physical pair line -> packed spectral state
  auto to_state = [&](cpx<float> (&line)[R]) {
    fft_reg<R, -1>(line, F.xb, j, F.tw2);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      U[j + R * r] = line[r];
      if (ORD == 2) zr[r] = line[r];
    }
    if (j == 0) {   // X1[N/2], X2[N/2] are real
      U[N] = cpx<float>(line[R / 2].x, 0.f);
      U[N + 1] = cpx<float>(line[R / 2].y, 0.f);
    }
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
    if (include_init && !final_only) store_phys(v, 0);
    to_state(v);
  }

  const float invN = p.P.inv_norm;
  for (long long s = 0; s < p.n_saved; ++s) {
    if (p.forcing) {  // u_hat += dt * f_hat (ForcedStepper): packed F1[k] + i F2[k], conjugated inputs for the negative half
      const cpx<float>* f1 = p.forcing + s * p.fstep + t1 * p.fbatch;
      const cpx<float>* f2 = p.forcing + s * p.fstep + t2 * p.fbatch;
      const cpx<float> zero(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = F.kidx(r);
        const cpx<float> a = act1 ? f1[k] : zero, b = act2 ? f2[k] : zero;
        cpx<float> add = F1::upper(r) ? conj(a) + mul_i(conj(b)) : a + mul_i(b);
        if (r == 0 && j == 0) add = cpx<float>(a.x, b.x);   // DC: real parts only (irfft)
        const cpx<float> uf = axpy(p.fscale, add, U[j + R * r]);
        U[j + R * r] = uf;
        if (ORD == 2) zr[r] = uf;
      }
      if (j == 0) {
        if (act1) U[N] = axpy(p.fscale, f1[N / 2], U[N]);
        if (act2) U[N + 1] = axpy(p.fscale, f2[N / 2], U[N + 1]);
      }
    }
    for (int sub = 0; sub < p.substeps; ++sub) {
      if constexpr (ORD == 2) F.etdrk2_step(zr);
      else F.etdrk_step(order);
    }
    const bool last = (s == p.n_saved - 1);
    const bool store = !final_only || last;
    if (store || !spectral_carry) {
      cpx<float> z[R];
      F.build_plain(U, z);
      if constexpr (ORD == 2) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r != R / 2) z[r] = zr[r];
        if (j != 0) z[R / 2] = zr[R / 2];
      }
      fft_reg<R, +1>(z, F.xb, j, F.tw2);
#pragma unroll
      for (int r = 0; r < R; ++r) z[r] = invN * z[r];    // one packed scaling serves the snapshot and the carry
      if (store) store_phys(z, final_only ? 0 : s + (include_init ? 1 : 0));
      if (last) break;
      if (!spectral_carry) {
        to_state(z);
        continue;
      }
    }
    if (last) break;
    // spectral carry: the reference's irfft -> rfft round trip == Hermitian projection (the packed DC slot holds the
    // real parts only by construction; the Nyquist values lose their imaginary parts)
    if (j == 0) {
      U[N].y = 0.f;
      U[N + 1].y = 0.f;
    }
  }
}

}  // namespace exb
