// Per-mode and per-point operators of the nonlinear functions and of the ETDRK stage updates.
// Everything here is pure register arithmetic shared by the 1-D persistent kernel and the
// 2-D / 3-D pass kernels.  Reference semantics: exponax/nonlin_fun/*.py, exponax/etdrk/*.py.
#pragma once
#include "exb_common.cuh"
#include "../../include/exb.h"

namespace exb {

// Compile-time description of the nonlinear function.  Members < 0 (or 0 for D / C) mean "read it
// from NlParams at run time" (generic kernels); the fast kernels pass fully static descriptors so
// that every `switch (kind)` below folds away.
//   VAR (convection only): bit 0 = conservative, bit 1 = single_channel
template <int KIND, int VAR, int DD, int CC> struct NlS {
  static constexpr int kind = KIND, var = VAR, D = DD, C = CC;
};
using NlDyn = NlS<-1, -1, 0, 0>;
#define EXB_S_KIND (S::kind >= 0 ? S::kind : P.kind)
#define EXB_S_D (S::D > 0 ? S::D : P.D)
#define EXB_S_C (S::C > 0 ? S::C : P.C)
#define EXB_S_CONS (S::var >= 0 ? (S::var & 1) : P.conservative)
#define EXB_S_SINGLE (S::var >= 0 ? ((S::var >> 1) & 1) : P.single_channel)

// wavenumber data of one spectral mode
template <class T> struct ModeK {
  T kd[3];     // (2*pi/L) * k_d ; derivative operator is i*kd[d]   (_spectral.py:86-115)
  bool keep;   // inside the dealiasing box |k_d| <= kmax (true when no mask) (_spectral.py:333-336)
  bool is_inj; // the single Kolmogorov injection mode
  bool is_dc;
};

// idx[d]: array index along spatial axis d (last axis: 0..N/2 directly)
template <class T, class S = NlDyn>
__device__ __forceinline__ ModeK<T> make_mode(const NlParams<T>& P, int i0, int i1, int i2) {
  ModeK<T> m;
  int idx[3] = {i0, i1, i2};
  bool keep = true, inj = P.has_inj != 0, dc = true;
  const int D = EXB_S_D;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (d < D) {
      int k = (d == D - 1) ? idx[d] : wavenumber_of(idx[d], P.N);
      m.kd[d] = P.dscale * (T)k;
      int ak = k < 0 ? -k : k;
      if (P.kmax >= 0 && ak > P.kmax) keep = false;
      if (idx[d] != P.inj_idx[d]) inj = false;
      if (k != 0) dc = false;
    } else {
      m.kd[d] = (T)0;
    }
  }
  m.keep = keep;
  m.is_inj = inj;
  m.is_dc = dc;
  return m;
}

// 1/x, correctly rounded (identical to the IEEE division the reference's `1 / laplacian` performs) but
// a single MUFU + Newton step instead of the full division sequence
__device__ __forceinline__ float exb_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double exb_rcp(double x) { return 1.0 / x; }

template <class T> __device__ __forceinline__ T pick3(const T* a, int i) {
  return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]);
}
template <class T> __device__ __forceinline__ cpx<T> pick3c(const cpx<T>* a, int i) {
  return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]);
}

// ---- inverse-transform input field f of the nonlinear function, at one mode ----------------
// uh[c]: stage input, channel c (c < C <= 3).  Pre-dealiasing (nonlin_fun/_base.py:117-137)
// is applied here: modes outside the mask yield 0.
template <class T, class S = NlDyn>
__device__ __forceinline__ cpx<T> nl_inv_field(const NlParams<T>& P, int f, const cpx<T>* uh,
                                               const ModeK<T>& m) {
  const cpx<T> zero((T)0, (T)0);
  if (!m.keep) return zero;
  const int C = EXB_S_C, D = EXB_S_D;
  switch (EXB_S_KIND) {
    case EXB_NL_CONVECTION: {
      if (EXB_S_CONS) return pick3c(uh, f);  // single or multi: u only
      if (EXB_S_SINGLE) {                     // u, d_d u   (_convection.py:207-217)
        if (f == 0) return uh[0];
        return mul_i(pick3(m.kd, f - 1) * uh[0]);
      }
      if (f < C) return pick3c(uh, f);            // (_convection.py:165-190)
      int c = (f - C) / D, d = (f - C) - c * D;
      return mul_i(pick3(m.kd, d) * pick3c(uh, c));
    }
    case EXB_NL_GRADIENT_NORM: {                  // (_gradient_norm.py:84-101)
      int c = f / D, d = f - c * D;
      return mul_i(pick3(m.kd, d) * pick3c(uh, c));
    }
    case EXB_NL_POLYNOMIAL:
    case EXB_NL_GRAY_SCOTT:
    case EXB_NL_CAHN_HILLIARD:
      return pick3c(uh, f);
    case EXB_NL_VORTICITY_2D: {                   // (_vorticity_convection.py:78-99)
      // laplacian = (i kd0)^2 + (i kd1)^2 ; inv = where(lap == 0, 1, 1/lap)
      T lap = -(m.kd[0] * m.kd[0]) - (m.kd[1] * m.kd[1]);
      T inv = (lap == (T)0) ? (T)1 : exb_rcp(lap);
      cpx<T> w = uh[0];
      cpx<T> psi = inv * w;
      if (f == 0) return mul_i(m.kd[1] * psi);    // u = +d_y psi
      if (f == 1) return mul_mi(m.kd[0] * psi);   // v = -d_x psi
      if (f == 2) return mul_i(m.kd[0] * w);      // d_x omega
      return mul_i(m.kd[1] * w);                  // d_y omega
    }
    case EXB_NL_PROJECTED_3D: {                   // (_projected_convection.py:114-136)
      if (f >= 3) return uh[f - 3];               // velocity
      // curl_hat = (i kd) x u_hat
      int a = (f + 1) % 3, b = (f + 2) % 3;
      cpx<T> t = pick3(m.kd, a) * pick3c(uh, b) - pick3(m.kd, b) * pick3c(uh, a);
      return mul_i(t);
    }
    case EXB_NL_GENERAL: {                        // fused (_general_nonlinear.py:111-119)
      if (f < C) return pick3c(uh, f);
      int c = (f - C) / D, d = (f - C) - c * D;
      return mul_i(pick3(m.kd, d) * pick3c(uh, c));
    }
    default:
      return zero;
  }
}

// ---- pointwise products in physical space ----------------------------------------------------
// in[f]: the n_inv inverse-transformed fields at one grid point (already scaled by 1/N^D);
// out[g]: the n_fwd fields to be forward-transformed.  V = T (one grid point) or f32x2 (the same grid point of
// the two rows / trajectories that share a two-for-one line: every operation is one packed instruction).
template <class T, class S = NlDyn, class V = T>
__device__ __forceinline__ void nl_pointwise(const NlParams<T>& P, const V* in, V* out) {
  const int C = EXB_S_C, D = EXB_S_D;
  switch (EXB_S_KIND) {
    case EXB_NL_CONVECTION: {
      if (EXB_S_CONS) {
        if (EXB_S_SINGLE) {
#pragma unroll
          for (int c = 0; c < EXB_MAXC; ++c)
            if (c < C) out[c] = in[c] * in[c];
        } else {
#pragma unroll
          for (int c = 0; c < EXB_MAXC; ++c)
#pragma unroll
            for (int d = 0; d < EXB_MAXC; ++d)
              if (c < C && d < C) out[c * C + d] = in[c] * in[d];
        }
      } else if (EXB_S_SINGLE) {
        V s = in[0] * in[1];
#pragma unroll
        for (int d = 1; d < 3; ++d)
          if (d < D) s = vfma(in[0], in[1 + d], s);
        out[0] = s;
      } else {
#pragma unroll
        for (int c = 0; c < EXB_MAXC; ++c) {
          if (c < C) {
            V s = in[0] * in[C + c * D];
#pragma unroll
            for (int d = 1; d < 3; ++d)
              if (d < D) s = vfma(in[d], in[C + c * D + d], s);
            out[c] = s;
          }
        }
      }
      break;
    }
    case EXB_NL_GRADIENT_NORM: {
#pragma unroll
      for (int c = 0; c < EXB_MAXC; ++c) {
        if (c < C) {
          V s = in[c * D] * in[c * D];
#pragma unroll
          for (int d = 1; d < 3; ++d)
            if (d < D) s = vfma(in[c * D + d], in[c * D + d], s);
          out[c] = s;
        }
      }
      break;
    }
    case EXB_NL_POLYNOMIAL: {
#pragma unroll
      for (int c = 0; c < EXB_MAXC; ++c) {
        if (c < C) {
          V u = in[c], pw = V((T)1), acc = V((T)0);
          for (int k = 0; k < P.n_poly; ++k) {  // (_polynomial.py:69-73)
            acc = vfma(V(P.poly[k]), pw, acc);
            pw = pw * u;
          }
          out[c] = acc;
        }
      }
      break;
    }
    case EXB_NL_VORTICITY_2D:
      out[0] = vfma(in[1], in[3], in[0] * in[2]);
      break;
    case EXB_NL_GRAY_SCOTT: {  // feed = gen[0], kill = gen[1]   (_gray_scott.py:36-43)
      V uvv = in[0] * (in[1] * in[1]);
      out[0] = V(P.gen[0]) * (V((T)1) - in[0]) - uvv;
      out[1] = vfma(V(-(P.gen[0] + P.gen[1])), in[1], uvv);
      break;
    }
    case EXB_NL_CAHN_HILLIARD:  // (_cahn_hilliard.py:33)
      out[0] = in[0] * in[0] * in[0];
      break;
    case EXB_NL_PROJECTED_3D: {
      // convection = velocity x curl   (in[0..2] = curl, in[3..5] = velocity)
      out[0] = vfma(in[4], in[2], -(in[5] * in[1]));
      out[1] = vfma(in[5], in[0], -(in[3] * in[2]));
      out[2] = vfma(in[3], in[1], -(in[4] * in[0]));
      break;
    }
    case EXB_NL_GENERAL: {
#pragma unroll
      for (int c = 0; c < EXB_MAXC; ++c) {
        if (c < C) {
          out[c] = in[c] * in[c];
          V s = in[C + c * D] * in[C + c * D];
#pragma unroll
          for (int d = 1; d < 3; ++d)
            if (d < D) s = vfma(in[C + c * D + d], in[C + c * D + d], s);
          out[C + c] = s;
        }
      }
      break;
    }
    default:
      break;
  }
}

// ---- N(u)_c at one mode from the forward-transformed fields W[g] ---------------------------
// Post-dealiasing (nonlin_fun/_base.py:99-115) applied here; the Kolmogorov injection is
// added un-masked.
template <class T, class S = NlDyn>
__device__ __forceinline__ void nl_from_fwd(const NlParams<T>& P, const cpx<T>* W, const ModeK<T>& m,
                                            cpx<T>* out) {
  const cpx<T> zero((T)0, (T)0);
  const int C = EXB_S_C, D = EXB_S_D;
#pragma unroll
  for (int c = 0; c < EXB_MAXC; ++c) out[c] = zero;
  if (m.keep) {
    switch (EXB_S_KIND) {
      case EXB_NL_CONVECTION: {
        if (EXB_S_CONS) {
          if (EXB_S_SINGLE) {  // -s * 0.5 * (sum_d i kd) * F(u^2)  (_convection.py:192-205)
            T sd = (T)0;
#pragma unroll
            for (int d = 0; d < 3; ++d)
              if (d < D) sd += m.kd[d];
#pragma unroll
            for (int c = 0; c < EXB_MAXC; ++c)
              if (c < C) out[c] = (-P.scale) * ((T)0.5 * mul_i(sd * W[c]));
          } else {                 // -s * 0.5 * sum_d i kd * F(u_c u_d)  (_convection.py:140-163)
#pragma unroll
            for (int c = 0; c < EXB_MAXC; ++c) {
              if (c < C) {
                cpx<T> s = zero;
#pragma unroll
                for (int d = 0; d < EXB_MAXC; ++d)
                  if (d < C) s = s + mul_i(m.kd[d] * W[c * C + d]);
                out[c] = (-P.scale) * ((T)0.5 * s);
              }
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < EXB_MAXC; ++c)
            if (c < (EXB_S_SINGLE ? 1 : C)) out[c] = (-P.scale) * W[c];
        }
        break;
      }
      case EXB_NL_GRADIENT_NORM: {
#pragma unroll
        for (int c = 0; c < EXB_MAXC; ++c)
          if (c < C) out[c] = (P.zero_mode_fix && m.is_dc) ? zero : (-P.scale) * ((T)0.5 * W[c]);
        break;
      }
      case EXB_NL_POLYNOMIAL: {
#pragma unroll
        for (int c = 0; c < EXB_MAXC; ++c)
          if (c < C) out[c] = W[c];
        break;
      }
      case EXB_NL_VORTICITY_2D:
        out[0] = (-P.scale) * W[0];
        break;
      case EXB_NL_GRAY_SCOTT:
        out[0] = W[0];
        out[1] = W[1];
        break;
      case EXB_NL_CAHN_HILLIARD: {  // scale * laplace * F[u^3]   (_cahn_hilliard.py:34-36)
        T lap = (T)0;
#pragma unroll
        for (int d = 0; d < 3; ++d)
          if (d < D) lap -= m.kd[d] * m.kd[d];
        out[0] = P.scale * (lap * W[0]);
        break;
      }
      case EXB_NL_PROJECTED_3D: {  // Leray projection (_leray.py:114-136)
        cpx<T> div = mul_i(m.kd[0] * W[0] + m.kd[1] * W[1] + m.kd[2] * W[2]);
        T lap = -(m.kd[0] * m.kd[0]) - (m.kd[1] * m.kd[1]) - (m.kd[2] * m.kd[2]);
        T inv = (lap != (T)0) ? exb_rcp(lap) : (T)0;
        cpx<T> p = (-inv) * div;
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = W[c] + mul_i(m.kd[c] * p);
        break;
      }
      case EXB_NL_GENERAL: {
        T sd = (T)0;
#pragma unroll
        for (int d = 0; d < 3; ++d)
          if (d < D) sd += m.kd[d];
#pragma unroll
        for (int c = 0; c < EXB_MAXC; ++c) {
          if (c < C) {
            cpx<T> sq = W[c];
            cpx<T> gn = (P.zero_mode_fix && m.is_dc) ? zero : W[C + c];
            out[c] = P.gen[0] * sq + P.gen[1] * ((T)0.5 * mul_i(sd * sq)) + P.gen[2] * ((T)0.5 * gn);
          }
        }
        break;
      }
      default:
        break;
    }
  }
  if (m.is_inj) out[0].x += P.inj_val;
}

// ---- ETDRK stage updates ------------------------------------------------------------------
// Buffers (all spectral, same indexing `off` = element offset of (trajectory, channel, mode)):
//   U   : state at the start of the step (read-only during the step)
//   OUT : result of the step (may alias U: each mode is read before it is written by one thread)
//   S[0..3] : scratch states.  Roles per order:
//     order 2: S0 = a, S1 = n0
//     order 3: S0 = a then b, S1 = n0, S2 = n1
//     order 4: S0 = a, S1 = n0, S2 = b then c, S3 = n1 then n1+n2
// stage_input(order, s): which buffer feeds the nonlinear function of stage s
//   (-1 = U, otherwise index into S).
__host__ __device__ inline int etdrk_stage_input(int order, int s) {
  if (s == 0) return -1;
  if (order == 4 && s >= 2) return 2;
  return 0;
}
__host__ __device__ inline int etdrk_num_scratch(int order) {
  return order <= 1 ? 0 : order;  // 2, 3, 4
}

template <class T> struct StateBufs {
  const cpx<T>* U;
  cpx<T>* OUT;
  cpx<T>* S[4];
};

// ci = coefficient index = (E == 1 ? 0 : c) * M + mode
template <class T>
__device__ __forceinline__ void etdrk_update(const EtdrkCoefs<T>& K, int stage, long long ci,
                                             size_t off, cpx<T> n, const StateBufs<T>& B) {
  const cpx<T> u = B.U[off];
  switch (K.order) {
    case 1:  // (_etdrk_1.py:78-82)
      B.OUT[off] = K.exp_term[ci] * u + K.c[0][ci] * n;
      break;
    case 2:  // (_etdrk_2.py:91-102)
      if (stage == 0) {
        B.S[0][off] = K.exp_term[ci] * u + K.c[0][ci] * n;
        B.S[1][off] = n;
      } else {
        B.OUT[off] = B.S[0][off] + K.c[1][ci] * (n - B.S[1][off]);
      }
      break;
    case 3:  // (_etdrk_3.py:191-212)
      if (stage == 0) {
        B.S[0][off] = K.half_exp[ci] * u + K.c[0][ci] * n;
        B.S[1][off] = n;
      } else if (stage == 1) {
        B.S[0][off] = K.exp_term[ci] * u + K.c[1][ci] * ((T)2 * n - B.S[1][off]);
        B.S[2][off] = n;
      } else {
        B.OUT[off] = K.exp_term[ci] * u + K.c[2][ci] * B.S[1][off] + K.c[3][ci] * B.S[2][off] +
                     K.c[4][ci] * n;
      }
      break;
    case 4:  // (_etdrk_4.py:198-224)
      if (stage == 0) {
        B.S[0][off] = K.half_exp[ci] * u + K.c[0][ci] * n;
        B.S[1][off] = n;
      } else if (stage == 1) {
        B.S[2][off] = K.half_exp[ci] * u + K.c[1][ci] * n;
        B.S[3][off] = n;
      } else if (stage == 2) {
        B.S[2][off] = K.half_exp[ci] * B.S[0][off] + K.c[2][ci] * ((T)2 * n - B.S[1][off]);
        B.S[3][off] = B.S[3][off] + n;
      } else {
        B.OUT[off] = K.exp_term[ci] * u + K.c[3][ci] * B.S[1][off] +
                     K.c[4][ci] * ((T)2 * B.S[3][off]) + K.c[5][ci] * n;
      }
      break;
    default:
      break;
  }
}

// Modes outside the dealiasing mask (and not the injection mode) have N(u) == 0 in EVERY stage
// (nonlin_fun/_base.py:99-115 zeroes them after the forward transform), so all ETDRK orders reduce
// there to  u+ = exp(dt L) u  exactly (E u + c * 0, a + c2 (0 - 0)): the intermediate stage buffers
// of such a mode are never written and never read (the prologue of the next stage pre-masks its
// input, nonlin_fun/_base.py:117-137), the last stage writes E u.  With the 2/3 rule that is 56 %
// (2-D) / 70 % (3-D) of all modes: 8 -> 2.5 HBM words per mode and step (ETDRK2) for them.
template <class T>
__device__ __forceinline__ void etdrk_update_masked(const EtdrkCoefs<T>& K, int stage, long long ci, size_t off,
                                                    const StateBufs<T>& B) {
  if (stage == K.order - 1) B.OUT[off] = K.exp_term[ci] * B.U[off];
}

// L2 prefetch of everything etdrk_update(stage) will read at (ci, off): issued by the column-pass
// epilogue kernels BEFORE their transforms, so that the DRAM latency of the stage operands (2/3 of the
// bytes these kernels read) overlaps the FFT instead of being paid once per mode afterwards.
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
template <class T>
__device__ __forceinline__ void etdrk_prefetch(const EtdrkCoefs<T>& K, int stage, long long ci, size_t off,
                                               const StateBufs<T>& B, bool coefs) {
  auto pc = [&](int i) { if (coefs) prefetch_l2(K.c[i] + ci); };
  auto pe = [&](const cpx<T>* e) { if (coefs) prefetch_l2(e + ci); };
  const int o = K.order;
  const bool last = stage == o - 1;
  if (o == 1 || (o == 2 && stage == 0) || o == 3 || (o == 4 && stage != 2)) prefetch_l2(B.U + off);
  switch (o) {
    case 1: pe(K.exp_term); pc(0); break;
    case 2:
      if (stage == 0) { pe(K.exp_term); pc(0); }
      else { prefetch_l2(B.S[0] + off); prefetch_l2(B.S[1] + off); pc(1); }
      break;
    case 3:
      if (stage == 0) { pe(K.half_exp); pc(0); }
      else if (stage == 1) { pe(K.exp_term); pc(1); prefetch_l2(B.S[1] + off); }
      else { pe(K.exp_term); pc(2); pc(3); pc(4); prefetch_l2(B.S[1] + off); prefetch_l2(B.S[2] + off); }
      break;
    case 4:
      if (stage == 0) { pe(K.half_exp); pc(0); }
      else if (stage == 1) { pe(K.half_exp); pc(1); }
      else if (stage == 2) { pe(K.half_exp); pc(2); prefetch_l2(B.S[0] + off); prefetch_l2(B.S[1] + off); prefetch_l2(B.S[3] + off); }
      else { pe(K.exp_term); pc(3); pc(4); pc(5); prefetch_l2(B.S[1] + off); prefetch_l2(B.S[3] + off); }
      break;
    default: break;
  }
  (void)last;
}

}  // namespace exb
