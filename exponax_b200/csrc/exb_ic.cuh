// Producers in front of the rollout loop: the random initial-condition generators of
// exponax/ic are  white noise -> fft -> per-mode real factor -> ifft -> normalise.  The transforms are the
// plan's own exb_fft / exb_ifft; this file holds the two pieces in between and after:
//   ic_shape_kernel      the per-mode factor, in place on u_hat
//       kind 0  RandomTruncatedFourierSeries (ic/_truncated_fourier_series.py:65-100): keep |k_d| <= cutoff on
//               every axis, then the DC entry := dc_value (an UNNORMALISED coefficient, as in the reference)
//       kind 1  GaussianRandomField (ic/_gaussian_random_field.py:64-93): |2 pi k / L|^(-exponent / 2), DC factor 1
//   ic_stats / ic_maxabs / ic_apply   normalize_ic (ic/_base_ic.py:16-33): x -= mean, x /= std, x /= max|x|
#pragma once
#include "exb_common.cuh"

namespace exb {

template <class T>
__global__ void ic_shape_kernel(cpx<T>* uh, int D, int N, int Nh, long long M, long long total, int kind, T param,
                                T two_pi_over_L, T dc_value) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long m = i % M;
  const int kl = (int)(m % Nh);
  long long rest = m / Nh;
  int kabs_max = kl;
  T n2 = (T)kl * (T)kl;
  for (int d = 1; d < D; ++d) {
    int k = wavenumber_of((int)(rest % N), N);
    rest /= N;
    k = k < 0 ? -k : k;
    kabs_max = k > kabs_max ? k : kabs_max;
    n2 += (T)k * (T)k;
  }
  cpx<T> v = uh[i];
  if (kind == 0) {
    if ((T)kabs_max > param) v = cpx<T>((T)0, (T)0);
    if (m == 0) v = cpx<T>(dc_value, (T)0);
  } else {
    const T amp = m == 0 ? (T)1 : pow(sqrt(n2) * two_pi_over_L, -param / (T)2);
    v = cpx<T>(v.x * amp, v.y * amp);
  }
  uh[i] = v;
}

// normalize_ic statistics, per field f, in double:  stats[4f + 3] = shift (the first element; makes the
// one-pass variance stable), stats[4f + 0] = sum (x - shift), stats[4f + 1] = sum (x - shift)^2,
// stats[4f + 2] = max |x - mean|  (mean taken as 0 unless zero_mean)
template <class T> __global__ void ic_shift_kernel(const T* u, long long npoints, long long nfields, double* stats) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f < nfields) stats[4 * f + 3] = (double)u[(size_t)f * npoints];
}

template <class T>
__global__ void __launch_bounds__(256) ic_stats_kernel(const T* __restrict__ u, long long npoints, long long chunk,
                                                       double* __restrict__ stats) {
  const long long f = blockIdx.y;
  const T* x = u + (size_t)f * npoints;
  const double shift = stats[4 * f + 3];
  const long long i0 = (long long)blockIdx.x * chunk;
  const long long i1 = i0 + chunk < npoints ? i0 + chunk : npoints;
  double s = 0, s2 = 0;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const double d = (double)x[i] - shift;
    s += d;
    s2 += d * d;
  }
  __shared__ double red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(&stats[4 * f + threadIdx.x], t);
  }
}

// non-negative doubles order like their bit patterns, so the maximum is an integer atomicMax
template <class T>
__global__ void __launch_bounds__(256) ic_maxabs_kernel(const T* __restrict__ u, long long npoints, long long chunk,
                                                        int zero_mean, double* __restrict__ stats) {
  const long long f = blockIdx.y;
  const T* x = u + (size_t)f * npoints;
  const double mean = zero_mean ? stats[4 * f + 3] + stats[4 * f] / (double)npoints : 0.0;
  const long long i0 = (long long)blockIdx.x * chunk;
  const long long i1 = i0 + chunk < npoints ? i0 + chunk : npoints;
  double m = 0;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) m = fmax(m, fabs((double)x[i] - mean));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0)
    atomicMax(reinterpret_cast<unsigned long long*>(&stats[4 * f + 2]), (unsigned long long)__double_as_longlong(m));
}

template <class T>
__global__ void ic_apply_kernel(T* u, long long npoints, long long total, int zero_mean, int std_one, int max_one,
                                const double* __restrict__ stats) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long f = i / npoints;
  const double n = (double)npoints;
  const double dmean = stats[4 * f] / n;
  const double mean = stats[4 * f + 3] + dmean;
  double x = (double)u[i];
  if (zero_mean) x -= mean;
  double scale = 1.0;
  if (max_one) {
    scale = 1.0 / stats[4 * f + 2];   // (x - mean) / std / max|(x - mean) / std| == (x - mean) / max|x - mean|
  } else if (std_one) {
    const double var = stats[4 * f + 1] / n - dmean * dmean;   // population variance (jnp.std, ddof = 0)
    scale = 1.0 / sqrt(var > 0 ? var : 0.0);
  }
  u[i] = (T)(x * scale);
}

}  // namespace exb
