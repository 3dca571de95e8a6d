// Fused reductions behind exponax.metrics (consumers of saved snapshots):
// ONE pass over a prediction / reference pair gives, per field (one channel of one sample),
//   s[0] = sum |a - b|^p     s[1] = sum |b|^p     s[2] = sum |a|^p     s[3] = sum a * b
// from which the host forms every spatial metric of the reference (metrics/_spatial.py:8-196:
// MAE / MSE / RMSE and their normalised / symmetric variants) and the correlation
// (metrics/_correlation.py:6-60) without further passes.  Per-thread partial sums in the working
// precision, block tree reduction and the cross-block atomics in double.
#pragma once
#include "exb_common.cuh"

namespace exb {

template <class T> __device__ __forceinline__ T abs_pow(T x, T p, int pi) {
  const T a = fabs(x);
  if (pi == 1) return a;
  if (pi == 2) return a * a;
  return pow(a, p);
}

template <class T>
__global__ void __launch_bounds__(256) metric_sums_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                          long long npoints, long long chunk, T p, int pi,
                                                          double* __restrict__ out) {
  const long long f = blockIdx.y;
  const T* af = a + (size_t)f * npoints;
  const T* bf = b ? b + (size_t)f * npoints : nullptr;
  const long long i0 = (long long)blockIdx.x * chunk;
  const long long i1 = i0 + chunk < npoints ? i0 + chunk : npoints;
  T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const T x = af[i];
    const T y = bf ? bf[i] : (T)0;
    s0 += abs_pow<T>(x - y, p, pi);
    s1 += abs_pow<T>(y, p, pi);
    s2 += abs_pow<T>(x, p, pi);
    s3 += x * y;
  }
  double d[4] = {(double)s0, (double)s1, (double)s2, (double)s3};
  __shared__ double red[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = d[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(&out[f * 4 + threadIdx.x], t);
  }
}

// Fourier-space aggregation behind exponax.metrics.fourier_* / H1_* (metrics/_fourier.py:15-140): one pass
// over x_hat = exb_fft(x) gives, per field and per derivative component d (one component without a
// derivative),   sum_modes  band(k) * ( |x_hat| * |2 pi k_d / L|^order )^p / recon(k)
// with the reference's clean-up of rounding noise (|x_hat| < 1e-5 -> 0), its band-pass
// (not all |k_d| <= low - 1) and (all |k_d| <= high), and the "reconstruction" scaling of the rfft layout.
template <class T> struct FourierSumParams {
  const cpx<T>* xh;   // (nfields, M)
  double* out;        // (nfields, ncomp), zeroed
  int D, N, Nh, ncomp;
  long long M, chunk;
  T p;
  int pi;             // 1 / 2: integer fast paths of |.|^p
  int filter, low, high;
  T order;            // < 0: no derivative
  T two_pi_over_L;
};

template <class T> __global__ void __launch_bounds__(256) fourier_sums_kernel(const FourierSumParams<T> q) {
  const long long f = blockIdx.y;
  const cpx<T>* x = q.xh + (size_t)f * q.M;
  const long long m0 = (long long)blockIdx.x * q.chunk;
  const long long m1 = m0 + q.chunk < q.M ? m0 + q.chunk : q.M;
  const int N = q.N, Nh = q.Nh, half = N / 2;
  T lead = (T)1;
  for (int d = 1; d < q.D; ++d) lead *= (T)N;
  double acc[3] = {0.0, 0.0, 0.0};
  for (long long m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
    int k[3] = {0, 0, 0};  // |k_d| per axis, axis order as in the state (last axis = rfft axis)
    const int kl = (int)(m % Nh);
    long long rest = m / Nh;
    k[q.D - 1] = kl;
    for (int d = q.D - 2; d >= 0; --d) {
      const int kk = wavenumber_of((int)(rest % N), N);
      rest /= N;
      k[d] = kk < 0 ? -kk : kk;
    }
    int kmax = 0;
    for (int d = 0; d < q.D; ++d) kmax = k[d] > kmax ? k[d] : kmax;
    if (q.filter && (kmax <= q.low - 1 || kmax > q.high)) continue;
    const cpx<T> v = x[m];
    T a = sqrt(v.x * v.x + v.y * v.y);
    if (a < (T)1e-5) continue;
    const T recon = lead * ((kl == 0 || (N % 2 == 0 && kl == half)) ? (T)N : (T)N / (T)2);
    for (int c = 0; c < q.ncomp; ++c) {
      T w = (T)1;
      if (q.order >= (T)0) w = q.order == (T)0 ? (T)1 : pow((T)k[c] * q.two_pi_over_L, q.order);
      acc[c] += (double)(abs_pow<T>(a * w, q.p, q.pi) / recon);
    }
  }
  __shared__ double red[3][8];
  for (int c = 0; c < q.ncomp; ++c) {
    double d = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((threadIdx.x & 31) == 0) red[c][threadIdx.x >> 5] = d;
  }
  __syncthreads();
  if (threadIdx.x < q.ncomp) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    if (t != 0.0) atomicAdd(&q.out[f * q.ncomp + threadIdx.x], t);
  }
}

}  // namespace exb
