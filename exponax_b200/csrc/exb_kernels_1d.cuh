// 1-D path: ONE persistent kernel per call.  A CTA owns a PAIR of trajectories; the pair shares
// every complex FFT through the two-for-one trick (z = x1 + i*x2), the spectral state and all
// ETDRK stage buffers stay in shared memory across every stage, sub-step and saved step, and
// HBM sees only the initial condition and the saved snapshots.
//
// Reference semantics (per trajectory):
//   rollout / repeat         exponax/_utils.py:92-254
//   step = fft -> step_fourier -> ifft   exponax/_base_stepper.py:201-220
//   RepeatedStepper sub-steps (spectral carry)   exponax/_repeated_stepper.py:56-102
//   ETDRK stages             exponax/etdrk/_etdrk_{0..4}.py
#pragma once
#include "exb_fft.cuh"
#include "exb_nl.cuh"

namespace exb {

enum Op1d { OP1_ROLLOUT = 0, OP1_STEP_FOURIER = 1, OP1_NL = 2, OP1_FFT = 3, OP1_IFFT = 4 };

template <class T> struct K1dParams {
  NlParams<T> P;
  EtdrkCoefs<T> K;
  FftDesc fd;
  const cpx<T>* tw;   // N-th roots of unity (device)
  long long batch;    // trajectories (FFT/IFFT ops: batch*channels single-channel rows)
  int op;
  int C;              // channels per trajectory for this launch (1 for FFT/IFFT ops)
  int nslots;         // FFT line slots per buffer
  const void* in;
  void* out;
  long long n_saved;
  int substeps;
  unsigned flags;
  // ForcedStepper / aux-taking rollouts (exponax/_forced_stepper.py:61-62: step(u + dt f)): before ETDRK step i the
  // Fourier transform of the forcing is added to the state, u_hat += fscale * f_hat[i * fstep + traj * fbatch + ...]
  // (the transform is linear, so this IS fft(u + dt f)); forcing == nullptr: none.  substeps must be 1.
  const cpx<T>* forcing;
  long long fstep, fbatch;   // element strides between steps (0: constant forcing) / trajectories (0: shared)
  T fscale;
};

template <class T> struct Ctx1d {
  const K1dParams<T>& p;
  cpx<T>* A;      // line buffer
  cpx<T>* B;      // line buffer
  StateBufs<T> sb;
  int N, Nh, C;
  long long t1, t2;
  bool has2;
  __device__ Ctx1d(const K1dParams<T>& q) : p(q) {}
};

// spectral state layout in shared memory: [(tr * C + c) * Nh + k], tr in {0,1}
template <class T> __device__ __forceinline__ size_t soff(const Ctx1d<T>& c, int tr, int ch, int k) {
  return (size_t)(tr * c.C + ch) * c.Nh + k;
}

// lines[c*N + x] = u1[c][x] + i*u2[c][x]
template <class T> __device__ void load_phys_pair(Ctx1d<T>& c, const T* src, cpx<T>* L) {
  const int N = c.N, C = c.C;
  for (int q = threadIdx.x; q < C * N; q += blockDim.x) {
    T a = src[(size_t)c.t1 * C * N + q];
    T b = c.has2 ? src[(size_t)c.t2 * C * N + q] : (T)0;
    L[q] = cpx<T>(a, b);
  }
}

// two-for-one split of `nf` forward-transformed lines Z into half-complex spectra:
// X1[k] = (Z[k] + conj(Z[N-k]))/2,  X2[k] = (Z[k] - conj(Z[N-k]))/(2i)
template <class T>
__device__ __forceinline__ void unpack2(const cpx<T>* Z, int N, int k, cpx<T>& X1, cpx<T>& X2) {
  cpx<T> z = Z[k];
  cpx<T> zp = Z[k == 0 ? 0 : N - k];
  X1 = cpx<T>((T)0.5 * (z.x + zp.x), (T)0.5 * (z.y - zp.y));
  X2 = cpx<T>((T)0.5 * (z.y + zp.y), (T)-0.5 * (z.x - zp.x));
}

// inverse of unpack2 with irfft semantics: imaginary parts of DC (and Nyquist for even N) are
// dropped (numpy/pocketfft/cuFFT C2R behaviour, SURVEY section 7 "hard parts")
template <class T>
__device__ __forceinline__ void pack2(cpx<T>* Z, int N, int k, cpx<T> F1, cpx<T> F2) {
  if (k == 0 || 2 * k == N) {
    Z[k] = cpx<T>(F1.x, F2.x);
  } else {
    Z[k] = cpx<T>(F1.x - F2.y, F1.y + F2.x);
    Z[N - k] = cpx<T>(F1.x + F2.y, F2.x - F1.y);
  }
}

// Evaluate the nonlinear function on the spectral state `src` (shared memory, pair layout).
// On return the forward-transformed fields are in the returned line buffer (synchronised).
template <class T> __device__ cpx<T>* eval_nl_1d(Ctx1d<T>& c, const cpx<T>* src) {
  const NlParams<T>& P = c.p.P;
  const int N = c.N, Nh = c.Nh, C = c.C;
  // prologue: build the n_inv packed inverse lines
  for (int k = threadIdx.x; k < Nh; k += blockDim.x) {
    ModeK<T> m = make_mode(P, k, 0, 0);
    cpx<T> u1[EXB_MAXC], u2[EXB_MAXC];
#pragma unroll
    for (int ch = 0; ch < EXB_MAXC; ++ch) {
      if (ch < C) {
        u1[ch] = src[soff(c, 0, ch, k)];
        u2[ch] = src[soff(c, 1, ch, k)];
      } else {
        u1[ch] = cpx<T>((T)0, (T)0);
        u2[ch] = u1[ch];
      }
    }
    for (int f = 0; f < P.n_inv; ++f)
      pack2(c.A + (size_t)f * N, N, k, nl_inv_field(P, f, u1, m), nl_inv_field(P, f, u2, m));
  }
  __syncthreads();
  cpx<T>* R = fft_lines<T, +1>(c.A, c.B, P.n_inv, N, 1, false, c.p.fd, c.p.tw);
  cpx<T>* O = (R == c.A) ? c.B : c.A;
  // pointwise products, both trajectories at once (re = trajectory 1, im = trajectory 2)
  for (int x = threadIdx.x; x < N; x += blockDim.x) {
    T i1[EXB_MAX_INV], i2[EXB_MAX_INV], o1[EXB_MAX_FWD], o2[EXB_MAX_FWD];
    for (int f = 0; f < P.n_inv; ++f) {
      cpx<T> v = R[(size_t)f * N + x];
      i1[f] = v.x * P.inv_norm;
      i2[f] = v.y * P.inv_norm;
    }
    nl_pointwise(P, i1, o1);
    nl_pointwise(P, i2, o2);
    for (int g = 0; g < P.n_fwd; ++g) R[(size_t)g * N + x] = cpx<T>(o1[g], o2[g]);
  }
  __syncthreads();
  return fft_lines<T, -1>(R, O, P.n_fwd, N, 1, false, c.p.fd, c.p.tw);
}

// N(u) of both trajectories at mode k from the forward lines W
template <class T>
__device__ __forceinline__ void nl_mode_pair(Ctx1d<T>& c, const cpx<T>* W, int k, cpx<T>* n1, cpx<T>* n2) {
  const NlParams<T>& P = c.p.P;
  ModeK<T> m = make_mode(P, k, 0, 0);
  cpx<T> w1[EXB_MAX_FWD], w2[EXB_MAX_FWD];
  for (int g = 0; g < P.n_fwd; ++g) unpack2(W + (size_t)g * c.N, c.N, k, w1[g], w2[g]);
  nl_from_fwd(P, w1, m, n1);
  nl_from_fwd(P, w2, m, n2);
}

// one ETDRK step on the shared-memory state (sb.U -> sb.OUT, both = U here)
template <class T> __device__ void etdrk_step_1d(Ctx1d<T>& c) {
  const EtdrkCoefs<T>& K = c.p.K;
  const int Nh = c.Nh, C = c.C;
  if (K.order == 0 && K.lin_matrix) {  // per-mode C x C matrix (Wave, _wave.py:175-197)
    for (int k = threadIdx.x; k < Nh; k += blockDim.x) {
#pragma unroll
      for (int tr = 0; tr < 2; ++tr) {
        cpx<T> u[EXB_MAXC], r[EXB_MAXC];
#pragma unroll
        for (int ch = 0; ch < EXB_MAXC; ++ch)
          if (ch < C) u[ch] = c.sb.U[soff(c, tr, ch, k)];
#pragma unroll
        for (int ci = 0; ci < EXB_MAXC; ++ci) {
          if (ci < C) {
            cpx<T> acc((T)0, (T)0);
#pragma unroll
            for (int cj = 0; cj < EXB_MAXC; ++cj)
              if (cj < C)
                acc = acc + K.exp_term[(long long)(ci * C + cj) * K.M + k + table_offset(K, tr ? c.t2 : c.t1)] * u[cj];
            r[ci] = acc;
          }
        }
#pragma unroll
        for (int ch = 0; ch < EXB_MAXC; ++ch)
          if (ch < C) c.sb.OUT[soff(c, tr, ch, k)] = r[ch];
      }
    }
    __syncthreads();
    return;
  }
  if (K.order == 0) {  // (_etdrk_0.py:30-34)
    for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
      int ch = q / Nh, k = q - ch * Nh;
      long long ci = (long long)(K.E == 1 ? 0 : ch) * K.M + k;
      size_t o1 = soff(c, 0, ch, k), o2 = soff(c, 1, ch, k);
      c.sb.OUT[o1] = K.exp_term[ci + table_offset(K, c.t1)] * c.sb.U[o1];
      c.sb.OUT[o2] = K.exp_term[ci + (c.has2 ? table_offset(K, c.t2) : 0)] * c.sb.U[o2];
    }
    __syncthreads();
    return;
  }
  for (int s = 0; s < K.order; ++s) {
    int si = etdrk_stage_input(K.order, s);
    const cpx<T>* src = si < 0 ? c.sb.U : c.sb.S[si];
    cpx<T>* W = eval_nl_1d(c, src);
    for (int k = threadIdx.x; k < Nh; k += blockDim.x) {
      cpx<T> n1[EXB_MAXC], n2[EXB_MAXC];
      nl_mode_pair(c, W, k, n1, n2);
#pragma unroll
      for (int ch = 0; ch < EXB_MAXC; ++ch) {
        if (ch < C) {
          long long ci = (long long)(K.E == 1 ? 0 : ch) * K.M + k;
          etdrk_update(K, s, ci + table_offset(K, c.t1), soff(c, 0, ch, k), n1[ch], c.sb);
          etdrk_update(K, s, ci + (c.has2 ? table_offset(K, c.t2) : 0), soff(c, 1, ch, k), n2[ch], c.sb);
        }
      }
    }
    __syncthreads();
  }
}

template <class T> __global__ void k1d_kernel(const K1dParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx1d<T> c(p);
  const int N = p.fd.N, Nh = N / 2 + 1, C = p.C;
  c.N = N;
  c.Nh = Nh;
  c.C = C;
  c.t1 = 2ll * blockIdx.x;
  c.t2 = c.t1 + 1;
  c.has2 = c.t2 < p.batch;
  cpx<T>* sm = reinterpret_cast<cpx<T>*>(smem_raw);
  c.A = sm;
  c.B = sm + (size_t)p.nslots * N;
  cpx<T>* st = c.B + (size_t)p.nslots * N;
  const size_t ssz = (size_t)2 * C * Nh;
  cpx<T>* U = st;
  c.sb.U = U;
  c.sb.OUT = U;
  for (int i = 0; i < 4; ++i) c.sb.S[i] = st + (size_t)(1 + i) * ssz;  // only etdrk_num_scratch used
  const T invN = p.P.inv_norm;

  if (p.op == OP1_FFT) {
    load_phys_pair(c, (const T*)p.in, c.A);
    __syncthreads();
    cpx<T>* R = fft_lines<T, -1>(c.A, c.B, C, N, 1, false, p.fd, p.tw);
    cpx<T>* out = (cpx<T>*)p.out;
    for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
      int ch = q / Nh, k = q - ch * Nh;
      cpx<T> X1, X2;
      unpack2(R + (size_t)ch * N, N, k, X1, X2);
      out[((size_t)c.t1 * C + ch) * Nh + k] = X1;
      if (c.has2) out[((size_t)c.t2 * C + ch) * Nh + k] = X2;
    }
    return;
  }
  if (p.op == OP1_IFFT) {
    const cpx<T>* in = (const cpx<T>*)p.in;
    for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
      int ch = q / Nh, k = q - ch * Nh;
      cpx<T> F1 = in[((size_t)c.t1 * C + ch) * Nh + k];
      cpx<T> F2 = c.has2 ? in[((size_t)c.t2 * C + ch) * Nh + k] : cpx<T>((T)0, (T)0);
      pack2(c.A + (size_t)ch * N, N, k, F1, F2);
    }
    __syncthreads();
    cpx<T>* R = fft_lines<T, +1>(c.A, c.B, C, N, 1, false, p.fd, p.tw);
    T* out = (T*)p.out;
    for (int q = threadIdx.x; q < C * N; q += blockDim.x) {
      cpx<T> v = R[q];
      out[(size_t)c.t1 * C * N + q] = v.x * invN;
      if (c.has2) out[(size_t)c.t2 * C * N + q] = v.y * invN;
    }
    return;
  }
  if (p.op == OP1_STEP_FOURIER || p.op == OP1_NL) {
    const cpx<T>* in = (const cpx<T>*)p.in;
    cpx<T>* out = (cpx<T>*)p.out;
    for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
      U[q] = in[(size_t)c.t1 * C * Nh + q];
      U[(size_t)C * Nh + q] = c.has2 ? in[(size_t)c.t2 * C * Nh + q] : cpx<T>((T)0, (T)0);
    }
    __syncthreads();
    if (p.op == OP1_NL) {
      cpx<T>* W = eval_nl_1d(c, U);
      for (int k = threadIdx.x; k < Nh; k += blockDim.x) {
        cpx<T> n1[EXB_MAXC], n2[EXB_MAXC];
        nl_mode_pair(c, W, k, n1, n2);
#pragma unroll
        for (int ch = 0; ch < EXB_MAXC; ++ch) {
          if (ch < C) {
            out[((size_t)c.t1 * C + ch) * Nh + k] = n1[ch];
            if (c.has2) out[((size_t)c.t2 * C + ch) * Nh + k] = n2[ch];
          }
        }
      }
      return;
    }
    for (int s = 0; s < p.substeps; ++s) etdrk_step_1d(c);
    for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
      out[(size_t)c.t1 * C * Nh + q] = U[q];
      if (c.has2) out[(size_t)c.t2 * C * Nh + q] = U[(size_t)C * Nh + q];
    }
    return;
  }

  // ---- OP1_ROLLOUT -------------------------------------------------------------------------
  const bool include_init = (p.flags & EXB_ROLLOUT_INCLUDE_INIT) != 0;
  const bool layout_tb = (p.flags & EXB_ROLLOUT_LAYOUT_TB) != 0;
  const bool final_only = (p.flags & EXB_ROLLOUT_FINAL_ONLY) != 0;
  const bool spectral_carry = (p.flags & EXB_ROLLOUT_SPECTRAL_CARRY) != 0;
  const long long Tn = final_only ? 1 : p.n_saved + (include_init ? 1 : 0);
  const size_t fsz = (size_t)C * N;
  T* out = (T*)p.out;
  auto slot_ptr = [&](long long traj, long long slot) -> T* {
    if (final_only) return out + (size_t)traj * fsz;
    return layout_tb ? out + ((size_t)slot * p.batch + traj) * fsz
                     : out + ((size_t)traj * Tn + slot) * fsz;
  };

  load_phys_pair(c, (const T*)p.in, c.A);
  __syncthreads();
  if (include_init && !final_only) {
    T* o1 = slot_ptr(c.t1, 0);
    T* o2 = c.has2 ? slot_ptr(c.t2, 0) : nullptr;
    for (int q = threadIdx.x; q < C * N; q += blockDim.x) {
      cpx<T> v = c.A[q];
      o1[q] = v.x;
      if (o2) o2[q] = v.y;
    }
  }
  {
    cpx<T>* R = fft_lines<T, -1>(c.A, c.B, C, N, 1, false, p.fd, p.tw);
    for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
      int ch = q / Nh, k = q - ch * Nh;
      cpx<T> X1, X2;
      unpack2(R + (size_t)ch * N, N, k, X1, X2);
      U[soff(c, 0, ch, k)] = X1;
      U[soff(c, 1, ch, k)] = X2;
    }
    __syncthreads();
  }
  for (long long s = 0; s < p.n_saved; ++s) {
    if (p.forcing) {
      const cpx<T>* f1 = p.forcing + s * p.fstep + c.t1 * p.fbatch;
      const cpx<T>* f2 = p.forcing + s * p.fstep + c.t2 * p.fbatch;
      for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
        int ch = q / Nh, k = q - ch * Nh;
        U[soff(c, 0, ch, k)] = axpy(p.fscale, f1[q], U[soff(c, 0, ch, k)]);
        if (c.has2) U[soff(c, 1, ch, k)] = axpy(p.fscale, f2[q], U[soff(c, 1, ch, k)]);
      }
      __syncthreads();
    }
    for (int sub = 0; sub < p.substeps; ++sub) etdrk_step_1d(c);
    const bool last = (s == p.n_saved - 1);
    const bool store = !final_only || last;
    if (!store && spectral_carry) {
      // irfft -> rfft round trip of the reference == Hermitian projection of the carry
      for (int q = threadIdx.x; q < 2 * C; q += blockDim.x) {
        U[(size_t)q * Nh].y = (T)0;
        if ((N & 1) == 0) U[(size_t)q * Nh + N / 2].y = (T)0;
      }
      __syncthreads();
      continue;
    }
    // ifft of the state (no dealiasing on the state itself, _base_stepper.py:213-220)
    for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
      int ch = q / Nh, k = q - ch * Nh;
      pack2(c.A + (size_t)ch * N, N, k, U[soff(c, 0, ch, k)], U[soff(c, 1, ch, k)]);
    }
    __syncthreads();
    cpx<T>* R = fft_lines<T, +1>(c.A, c.B, C, N, 1, false, p.fd, p.tw);
    cpx<T>* O = (R == c.A) ? c.B : c.A;
    if (store) {
      long long slot = final_only ? 0 : s + (include_init ? 1 : 0);
      T* o1 = slot_ptr(c.t1, slot);
      T* o2 = c.has2 ? slot_ptr(c.t2, slot) : nullptr;
      for (int q = threadIdx.x; q < C * N; q += blockDim.x) {
        cpx<T> v = R[q];
        v.x *= invN;
        v.y *= invN;
        R[q] = v;
        o1[q] = v.x;
        if (o2) o2[q] = v.y;
      }
    } else {
      for (int q = threadIdx.x; q < C * N; q += blockDim.x) {
        cpx<T> v = R[q];
        R[q] = cpx<T>(v.x * invN, v.y * invN);
      }
    }
    if (last) break;
    if (spectral_carry) {
      for (int q = threadIdx.x; q < 2 * C; q += blockDim.x) {
        U[(size_t)q * Nh].y = (T)0;
        if ((N & 1) == 0) U[(size_t)q * Nh + N / 2].y = (T)0;
      }
      __syncthreads();
    } else {
      // reference carry: the next step starts from fft(u_next)
      __syncthreads();
      cpx<T>* R2 = fft_lines<T, -1>(R, O, C, N, 1, false, p.fd, p.tw);
      for (int q = threadIdx.x; q < C * Nh; q += blockDim.x) {
        int ch = q / Nh, k = q - ch * Nh;
        cpx<T> X1, X2;
        unpack2(R2 + (size_t)ch * N, N, k, X1, X2);
        U[soff(c, 0, ch, k)] = X1;
        U[soff(c, 1, ch, k)] = X2;
      }
      __syncthreads();
    }
  }
}

}  // namespace exb
