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

}  // namespace exb
