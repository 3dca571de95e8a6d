// Measured peaks of the two on-chip resources that bound the register-FFT kernels ("measure it"):
//   * FP32 throughput: dependent-chain-free FFMA loops, scalar (FFMA) and packed (fma.rn.f32x2 -> FFMA2)
//   * shared-memory bandwidth: conflict-free 8-byte (the kernels' complex<float> accesses) and 16-byte loads
// Measurement utilities behind the C ABI (exb_peak_fp32 / exb_peak_smem); bench.py reports c2's fraction of them.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/exb.h"

namespace {

template <int PACKED> __global__ void __launch_bounds__(512) fp32_peak_kernel(float* out, int iters, float b, float c) {
  // 16 independent accumulators per thread (8 float2): latency 4 x 2 issue cycles is covered many times over
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
  const float2 bb = make_float2(b, b), cc = make_float2(c, c);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (PACKED) {
          unsigned long long ra, rb, rc;
          ra = *reinterpret_cast<unsigned long long*>(&a[i]);
          rb = *reinterpret_cast<const unsigned long long*>(&bb);
          rc = *reinterpret_cast<const unsigned long long*>(&cc);
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(ra) : "l"(rb), "l"(rc));
          *reinterpret_cast<unsigned long long*>(&a[i]) = ra;
        } else {
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(b), "f"(c));
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].y) : "f"(b), "f"(c));
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  if (s == 123.456f) out[0] = s;  // never true; keeps the chain alive
}

template <int VEC> __global__ void __launch_bounds__(512) smem_peak_kernel(float* out, int iters) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* sm = reinterpret_cast<float*>(smem);
  constexpr int WORDS = 16384;  // 64 KB window
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  float acc = 0.f;
  // lane-contiguous accesses (conflict-free), a new row every iteration
  int idx = threadIdx.x * VEC;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int o = (idx + r * 512 * VEC) & (WORDS - 1);
      if (VEC == 2) {
        float2 v = *reinterpret_cast<const float2*>(sm + o);
        acc += v.x + v.y;
      } else {
        float4 v = *reinterpret_cast<const float4*>(sm + o);
        acc += (v.x + v.y) + (v.z + v.w);
      }
    }
    idx += 8 * 512 * VEC + VEC;  // keeps the compiler from hoisting the loads; stays VEC-aligned
  }
  if (acc == 123.456f) out[0] = acc;
}

int time_launches(cudaStream_t st, void (*launch)(cudaStream_t, float*, int), float* scratch, int iters, float* ms) {
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return EXB_ECUDA;
  launch(st, scratch, iters);  // warm-up
  launch(st, scratch, iters);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, st);
    launch(st, scratch, iters);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) return EXB_ECUDA;
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    if (t < best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms = best;
  return cudaGetLastError() == cudaSuccess ? EXB_OK : EXB_ECUDA;
}

constexpr int kBlocks = 148 * 8, kThreads = 512;
void launch_fp32_scalar(cudaStream_t st, float* s, int it) { fp32_peak_kernel<0><<<kBlocks, kThreads, 0, st>>>(s, it, 1.0001f, 1e-7f); }
void launch_fp32_packed(cudaStream_t st, float* s, int it) { fp32_peak_kernel<1><<<kBlocks, kThreads, 0, st>>>(s, it, 1.0001f, 1e-7f); }
void launch_smem8(cudaStream_t st, float* s, int it) { smem_peak_kernel<2><<<kBlocks, kThreads, 65536, st>>>(s, it); }
void launch_smem16(cudaStream_t st, float* s, int it) { smem_peak_kernel<4><<<kBlocks, kThreads, 65536, st>>>(s, it); }

}  // namespace

extern "C" {

// FP32 FMA throughput of the device in TFLOP/s: [0] scalar FFMA, [1] packed FFMA2 (fma.rn.f32x2).
int exb_peak_fp32(void* stream, double* tflops2) {
  if (!tflops2) return EXB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  float* scratch = nullptr;
  if (cudaMalloc(&scratch, 256) != cudaSuccess) return EXB_ECUDA;
  const int iters = 4096;
  const double flops = 2.0 * 16 * 4 * (double)iters * kBlocks * kThreads;
  float ms = 0.f;
  int rc = time_launches(st, launch_fp32_scalar, scratch, iters, &ms);
  if (rc == EXB_OK) {
    tflops2[0] = flops / (ms * 1e-3) / 1e12;
    rc = time_launches(st, launch_fp32_packed, scratch, iters, &ms);
    if (rc == EXB_OK) tflops2[1] = flops / (ms * 1e-3) / 1e12;
  }
  cudaFree(scratch);
  return rc;
}

// Shared-memory load bandwidth of the device in GB/s: [0] 8-byte loads (LDS.64), [1] 16-byte loads (LDS.128).
int exb_peak_smem(void* stream, double* gbs2) {
  if (!gbs2) return EXB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  float* scratch = nullptr;
  if (cudaMalloc(&scratch, 256) != cudaSuccess) return EXB_ECUDA;
  cudaFuncSetAttribute(smem_peak_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(smem_peak_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int iters = 2048;
  float ms = 0.f;
  int rc = time_launches(st, launch_smem8, scratch, iters, &ms);
  if (rc == EXB_OK) {
    gbs2[0] = 8.0 * 8 * (double)iters * kBlocks * kThreads / (ms * 1e-3) / 1e9;
    rc = time_launches(st, launch_smem16, scratch, iters, &ms);
    if (rc == EXB_OK) gbs2[1] = 16.0 * 8 * (double)iters * kBlocks * kThreads / (ms * 1e-3) / 1e9;
  }
  cudaFree(scratch);
  return rc;
}

}  // extern "C"
