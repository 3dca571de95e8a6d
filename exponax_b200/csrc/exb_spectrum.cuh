// Radially binned Fourier spectrum of half-complex fields (the consumer side of the rollout loop:
// exponax/_spectral.py:866-1030 `get_spectrum`, after its `fft`).  One pass over u_hat; every CTA
// accumulates its modes into shared-memory bins and flushes them with one atomic per bin.
//   amplitude: |u_hat| / recon          recon = N^(D-1) * (N at k_last in {0, N/2}, N/2 otherwise)
//   power:     0.5 * (|u_hat| / recon) * (|u_hat| / N^D)          (_spectral.py:986-1003)
//   bin b collects the modes with b - 1/2 <= |k| < b + 1/2, b = 0..N/2 (modes outside the Nyquist
//   sphere are dropped, _spectral.py:1011-1019); the half-integer test is done in exact integer
//   arithmetic (4|k|^2 against (2b -+ 1)^2), which is what the reference's f32 comparison resolves to for
//   every N <= 2048.
#pragma once
#include "exb_common.cuh"

namespace exb {

template <class T> struct SpectrumParams {
  const cpx<T>* uh;   // (nfields, M)
  T* out;             // (nfields, Nh), zeroed
  unsigned* counts;   // (Nh) or nullptr, zeroed: number of modes per bin (field 0 only)
  int D, N, Nh;
  long long M;
  int power;
  long long chunk;    // modes per CTA
};

__device__ __forceinline__ int spectrum_bin(long long n2) {
  int b = (int)(sqrt((double)n2) + 0.5);
  // exact: (2b-1)^2 <= 4 n2 < (2b+1)^2
  while ((long long)(2 * b + 1) * (2 * b + 1) <= 4 * n2) ++b;
  while (b > 0 && (long long)(2 * b - 1) * (2 * b - 1) > 4 * n2) --b;
  return b;
}

template <class T> __global__ void __launch_bounds__(256) spectrum_kernel(const SpectrumParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* bins = reinterpret_cast<T*>(smem_raw);
  unsigned* cnt = reinterpret_cast<unsigned*>(bins + p.Nh);
  const bool count = p.counts != nullptr && blockIdx.y == 0;
  for (int i = threadIdx.x; i < p.Nh; i += blockDim.x) {
    bins[i] = (T)0;
    cnt[i] = 0u;
  }
  __syncthreads();
  const cpx<T>* u = p.uh + (size_t)blockIdx.y * p.M;
  const long long m0 = (long long)blockIdx.x * p.chunk;
  const long long m1 = m0 + p.chunk < p.M ? m0 + p.chunk : p.M;
  const int N = p.N, Nh = p.Nh, half = N / 2;
  T lead = (T)1;  // N^(D-1)
  for (int d = 1; d < p.D; ++d) lead *= (T)N;
  const T norm = lead * (T)N;
  for (long long m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
    const int kl = (int)(m % Nh);
    long long rest = m / Nh;
    long long n2 = (long long)kl * kl;
    for (int d = 1; d < p.D; ++d) {
      const int k = wavenumber_of((int)(rest % N), N);
      rest /= N;
      n2 += (long long)k * k;
    }
    const int b = spectrum_bin(n2);
    if (b > half) continue;
    const cpx<T> v = u[m];
    const T a = sqrt(v.x * v.x + v.y * v.y);
    const T recon = lead * ((kl == 0 || (N % 2 == 0 && kl == half)) ? (T)N : (T)N / (T)2);
    const T mag = a / recon;
    const T q = p.power ? (T)0.5 * mag * (a / norm) : mag;
    atomicAdd(&bins[b], q);
    if (count) atomicAdd(&cnt[b], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.Nh; i += blockDim.x) {
    if (bins[i] != (T)0) atomicAdd(&p.out[(size_t)blockIdx.y * p.Nh + i], bins[i]);
    if (count && cnt[i]) atomicAdd(&p.counts[i], cnt[i]);
  }
}

// radial_binning="average": nanmean over the bin (an empty bin gives NaN, as jnp.nanmean does)
template <class T>
__global__ void spectrum_average_kernel(T* out, const unsigned* counts, int Nh, long long nfields) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfields * Nh) return;
  const unsigned c = counts[i % Nh];
  out[i] = c ? out[i] / (T)c : (T)NAN;
}

// exponax.derivative (exponax/_spectral.py:724-792) between its fft and ifft:
//   out[(f * D + d) * M + m] = u_hat[f * M + m] * (i * 2 pi k_d / L)^order        (integer order >= 0)
// The Nyquist entries are left as the reference leaves them (no "fix"); the inverse transform drops the
// imaginary part an odd order produces on the rfft axis, exactly as irfftn does.
template <class T>
__global__ void derivative_kernel(const cpx<T>* __restrict__ uh, cpx<T>* __restrict__ out, int D, int N, int Nh,
                                  long long M, long long total, int order, T two_pi_over_L) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long f = i / M, m = i - f * M;
  int k[3] = {0, 0, 0};
  k[D - 1] = (int)(m % Nh);
  long long rest = m / Nh;
  for (int d = D - 2; d >= 0; --d) {
    k[d] = wavenumber_of((int)(rest % N), N);
    rest /= N;
  }
  const cpx<T> v = uh[i];
  // v * i^order
  cpx<T> r;
  switch (order & 3) {
    case 0: r = v; break;
    case 1: r = cpx<T>(-v.y, v.x); break;
    case 2: r = cpx<T>(-v.x, -v.y); break;
    default: r = cpx<T>(v.y, -v.x); break;
  }
  for (int d = 0; d < D; ++d) {
    const T kc = (T)k[d] * two_pi_over_L;
    T w = (T)1;
    for (int e = 0; e < order; ++e) w *= kc;
    out[((size_t)f * D + d) * M + m] = cpx<T>(r.x * w, r.y * w);
  }
}

// Leray projection of a D-channel spectral field (exponax/nonlin_fun/_leray.py:114-136, the default order 2):
//   div = sum_d (i kd_d) u_d,   p = -inv_lap * div  (inv_lap = 1 / sum_d (i kd_d)^2, := 0 at k = 0),   out_c = u_c + i kd_c p
// one thread per mode, all D channels; `out` may alias `uh`.
template <class T>
__global__ void leray_kernel(const cpx<T>* uh, cpx<T>* out, int D, int N, int Nh, long long M, long long total,
                             T two_pi_over_L) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long f = i / M, m = i - f * M;
  T kd[3] = {(T)0, (T)0, (T)0};
  kd[D - 1] = two_pi_over_L * (T)(int)(m % Nh);
  long long rest = m / Nh;
  for (int d = D - 2; d >= 0; --d) {
    kd[d] = two_pi_over_L * (T)wavenumber_of((int)(rest % N), N);
    rest /= N;
  }
  cpx<T> u[3];
  cpx<T> s((T)0, (T)0);
  T lap = (T)0;
  for (int d = 0; d < D; ++d) {
    u[d] = uh[((size_t)f * D + d) * M + m];
    s = s + kd[d] * u[d];
    lap -= kd[d] * kd[d];
  }
  const cpx<T> div = mul_i(s);
  const T inv = lap != (T)0 ? (T)1 / lap : (T)0;
  const cpx<T> p = (-inv) * div;
  for (int d = 0; d < D; ++d) out[((size_t)f * D + d) * M + m] = u[d] + mul_i(kd[d] * p);
}

}  // namespace exb
