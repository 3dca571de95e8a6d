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

// ---- packed FP32 (sm_100a: FADD2 / FMUL2 / FFMA2) -------------------------------------------------------
// A cpx<float> is one 64-bit register pair, and Blackwell's packed-FP32 instructions take per-operand lane
// swaps and per-lane negations (SASS `.F32x2.LO_HI`, `.NP`), so complex adds, multiplications by +-i folded
// into an add, complex products (2 instructions) and real scalings each cost HALF the issue slots of the
// scalar sequences -- the register-FFT kernels are issue-bound, not FMA-pipe-bound (FFMA2 chains measure the
// same 74 TFLOP/s as FFMA chains, exb_peak_fp32).  Non-template overloads: preferred over the templates above
// for T = float in device code; identical IEEE results per lane (round-to-nearest, products fused as before).
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(EXB_NO_PACKED_F32)
#define EXB_PACKED_F32 1
__device__ __forceinline__ float2 exb_f2(cpx<float> a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ cpx<float> exb_c(float2 a) { return cpx<float>(a.x, a.y); }
__device__ __forceinline__ cpx<float> operator+(cpx<float> a, cpx<float> b) { return exb_c(__fadd2_rn(exb_f2(a), exb_f2(b))); }
__device__ __forceinline__ cpx<float> operator-(cpx<float> a, cpx<float> b) {
  return exb_c(__fadd2_rn(exb_f2(a), make_float2(-b.x, -b.y)));
}
__device__ __forceinline__ cpx<float> operator*(cpx<float> a, cpx<float> b) {
  const float2 t = __fmul2_rn(exb_f2(a), make_float2(b.x, b.x));
  return exb_c(__ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), t));
}
__device__ __forceinline__ cpx<float> operator*(float s, cpx<float> a) { return exb_c(__fmul2_rn(exb_f2(a), make_float2(s, s))); }
__device__ __forceinline__ cpx<float> operator*(cpx<float> a, float s) { return exb_c(__fmul2_rn(exb_f2(a), make_float2(s, s))); }
__device__ __forceinline__ cpx<float> mul_conj(cpx<float> a, cpx<float> b) {  // a * conj(b)
  const float2 t = __fmul2_rn(exb_f2(a), make_float2(b.x, b.x));
  return exb_c(__ffma2_rn(make_float2(a.y, a.x), make_float2(b.y, -b.y), t));
}
#else
#define EXB_PACKED_F32 0
#endif
// s * a + b (real scale, complex accumulate) and the lane-wise product-accumulate a (.) b + c used by the
// pointwise nonlinearities on (row 1, row 2) pairs
template <class T> __host__ __device__ inline cpx<T> axpy(T s, cpx<T> a, cpx<T> b) { return cpx<T>(s * a.x + b.x, s * a.y + b.y); }
template <class T> __host__ __device__ inline cpx<T> lane_fma(cpx<T> a, cpx<T> b, cpx<T> c) {
  return cpx<T>(a.x * b.x + c.x, a.y * b.y + c.y);
}
#if EXB_PACKED_F32
__device__ __forceinline__ cpx<float> axpy(float s, cpx<float> a, cpx<float> b) {
  return exb_c(__ffma2_rn(exb_f2(a), make_float2(s, s), exb_f2(b)));
}
__device__ __forceinline__ cpx<float> lane_fma(cpx<float> a, cpx<float> b, cpx<float> c) {
  return exb_c(__ffma2_rn(exb_f2(a), exb_f2(b), exb_f2(c)));
}
#endif

// ---- two independent real lanes (the two rows / trajectories of a two-for-one line at one grid point) ----
// The pointwise nonlinearities act on both lanes with the same operations: one packed instruction each.
struct alignas(8) f32x2 {
  float x, y;
  __host__ __device__ f32x2() {}
  __host__ __device__ f32x2(float a) : x(a), y(a) {}
  __host__ __device__ f32x2(float a, float b) : x(a), y(b) {}
};
#if EXB_PACKED_F32
__device__ __forceinline__ f32x2 operator+(f32x2 a, f32x2 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return f32x2(r.x, r.y); }
__device__ __forceinline__ f32x2 operator-(f32x2 a, f32x2 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y)); return f32x2(r.x, r.y); }
__device__ __forceinline__ f32x2 operator*(f32x2 a, f32x2 b) { float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return f32x2(r.x, r.y); }
__device__ __forceinline__ f32x2 vfma(f32x2 a, f32x2 b, f32x2 c) { float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y)); return f32x2(r.x, r.y); }
#else
__host__ __device__ inline f32x2 operator+(f32x2 a, f32x2 b) { return f32x2(a.x + b.x, a.y + b.y); }
__host__ __device__ inline f32x2 operator-(f32x2 a, f32x2 b) { return f32x2(a.x - b.x, a.y - b.y); }
__host__ __device__ inline f32x2 operator*(f32x2 a, f32x2 b) { return f32x2(a.x * b.x, a.y * b.y); }
__host__ __device__ inline f32x2 vfma(f32x2 a, f32x2 b, f32x2 c) { return f32x2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
__host__ __device__ inline f32x2 operator-(f32x2 a) { return f32x2(-a.x, -a.y); }
__host__ __device__ inline f32x2 operator*(float s, f32x2 a) { return f32x2(s) * a; }
__host__ __device__ inline float vfma(float a, float b, float c) { return a * b + c; }
__host__ __device__ inline double vfma(double a, double b, double c) { return a * b + c; }
// view of a packed line value (re = lane of row 1, im = lane of row 2) as two real lanes, and back
__host__ __device__ inline f32x2 lanes(cpx<float> a) { return f32x2(a.x, a.y); }
__host__ __device__ inline cpx<float> as_cpx(f32x2 a) { return cpx<float>(a.x, a.y); }

// FFT factorisation of one axis (host-built, passed by value to kernels)
struct FftDesc {
  int N;
  int nst;
  int radix[EXB_MAX_STAGES];
  // per stage, host-computed (exb_api.cu: factorize): the butterfly index arithmetic of the shared-memory Stockham
  // transform without run-time integer divisions.  m_* = floor(2^32 / d) + 1: q / d == __umulhi(q, m) for q * d < 2^32.
  int tunit[EXB_MAX_STAGES];        // N / (Ns * R): twiddle index step of the stage
  unsigned m_ns[EXB_MAX_STAGES];    // for d = Ns (product of the radices already applied)
  unsigned m_nr[EXB_MAX_STAGES];    // for d = N / R (butterflies per line)
};
__host__ __device__ inline unsigned fastdiv_magic(unsigned d) { return (unsigned)(0x100000000ull / d) + 1u; }
__device__ __forceinline__ int fastdiv(int q, int d, unsigned m) { return d == 1 ? q : (int)__umulhi((unsigned)q, m); }

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
  int i1_off;         // slab decomposition (3-D): global axis-1 index of local index l = l * i1_mul + i1_off
  int i1_mul;         //   block distribution: (1, rank * N/P);  cyclic: (P, rank);  single GPU: (1, 0)
};

// ETDRK coefficient tables (device pointers); E = 1 or C leading extent, M modes each.
template <class T> struct EtdrkCoefs {
  int order;
  int lin_matrix;   // order 0: exp_term holds a per-mode C x C matrix, entry [(i*C + j) * M + mode]
  // stepper ensembles: trajectory b reads its tables at offset (b / trep) * tstride (tstride = E * M elements per
  // table set, 0 = one shared set; trep = trajectories per set, filled in per launch)
  long long tstride;
  long long trep;
  int E;
  long long M;
  const cpx<T>* exp_term;
  const cpx<T>* half_exp;
  const T* c[6];
};

template <class T> __host__ __device__ inline long long table_offset(const EtdrkCoefs<T>& K, long long b) {
  return K.tstride ? (b / K.trep) * K.tstride : 0;
}

__host__ __device__ inline int wavenumber_of(int idx, int N) {
  // fftfreq ordering on full axes: 0..ceil(N/2)-1, -floor(N/2)..-1  (_spectral.py:40-41)
  return (idx < (N + 1) / 2) ? idx : idx - N;
}

}  // namespace exb
