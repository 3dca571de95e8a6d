// Shared device-side types for libexb (sm_100a).  Complex numbers are interleaved (re, im)
// pairs, 8 B (f32) / 16 B (f64) aligned so that one complex value is one vector load/store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EXB_MAX_STAGES 24
#define EXB_MAX_PEERS 8     // ranks of one slab-decomposed field (one NVSwitch domain)
#define EXB_MAXC 3        // channels handled in registers by the per-mode operators
#define EXB_MAX_INV 12    // inverse fields per nonlinear evaluation (3-D multi-channel convection)
#define EXB_MAX_FWD 9     // forward fields (3-D conservative multi-channel convection)

namespace exb {

template <class T> struct alignas(2 * sizeof(T)) cpx {
  T x, y;
  __host__ __device__ cpx() {}
  __host__ __device__ cpx(T a, T b) : x(a), y(b) {}
};

template <class T> __host__ __device__ inline cpx<T> operator+(cpx<T> a, cpx<T> b) { return cpx<T>(a.x + b.x, a.y + b.y); }
template <class T> __host__ __device__ inline cpx<T> operator-(cpx<T> a, cpx<T> b) { return cpx<T>(a.x - b.x, a.y - b.y); }
template <class T> __host__ __device__ inline cpx<T> operator-(cpx<T> a) { return cpx<T>(-a.x, -a.y); }
template <class T> __host__ __device__ inline cpx<T> operator*(cpx<T> a, cpx<T> b) {
  return cpx<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class T> __host__ __device__ inline cpx<T> operator*(T s, cpx<T> a) { return cpx<T>(s * a.x, s * a.y); }
template <class T> __host__ __device__ inline cpx<T> operator*(cpx<T> a, T s) { return cpx<T>(s * a.x, s * a.y); }
template <class T> __host__ __device__ inline cpx<T> conj(cpx<T> a) { return cpx<T>(a.x, -a.y); }
// multiply by +i / -i
template <class T> __host__ __device__ inline cpx<T> mul_i(cpx<T> a) { return cpx<T>(-a.y, a.x); }
template <class T> __host__ __device__ inline cpx<T> mul_mi(cpx<T> a) { return cpx<T>(a.y, -a.x); }
// a * conj(b)
template <class T> __host__ __device__ inline cpx<T> mul_conj(cpx<T> a, cpx<T> b) {
  return cpx<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// FFT factorisation of one axis (host-built, passed by value to kernels)
struct FftDesc {
  int N;
  int nst;
  int radix[EXB_MAX_STAGES];
};

// Parameters of the nonlinear function + spectral geometry, passed by value.
template <class T> struct NlParams {
  int kind;           // EXB_NL_*
  int D, N, Nh, C;    // spatial dims, points, N/2+1, channels
  int n_inv, n_fwd;   // single-field inverse / forward transforms per evaluation
  int kmax;           // dealias cutoff (inclusive); <0: none
  int single_channel, conservative, zero_mode_fix;
  int n_poly;
  int has_inj;
  int inj_idx[3];
  T inj_val;
  T dscale;           // (T)(2*pi/L)
  T scale;            // nl_scale
  T poly[8];
  T gen[3];
  T inv_norm;         // 1/N^D
  int i1_off;         // slab decomposition (3-D): global axis-1 index of local index 0
};

// ETDRK coefficient tables (device pointers); E = 1 or C leading extent, M modes each.
template <class T> struct EtdrkCoefs {
  int order;
  int E;
  long long M;
  const cpx<T>* exp_term;
  const cpx<T>* half_exp;
  const T* c[6];
};

__host__ __device__ inline int wavenumber_of(int idx, int N) {
  // fftfreq ordering on full axes: 0..ceil(N/2)-1, -floor(N/2)..-1  (_spectral.py:40-41)
  return (idx < (N + 1) / 2) ? idx : idx - N;
}

}  // namespace exb
