// TMA-staged column pass (sm_100a: cp.async.bulk.tensor + mbarrier) for the strided middle axis of the fast 3-D
// path -- COL_PLAIN over axis 1 of the pitch-padded field buffers (n_inv / n_fwd fields between the passes of one
// N(u) evaluation; exb_api.cu: Nhp).  A PERSISTENT CTA walks over [N x TW] tiles (TW last-axis wavenumbers of one
// x-plane of one field); thread 0 drives a two-deep ring:
//     TMA load of tile i+1 (global -> shared, completion on an mbarrier)   || register FFT of tile i  ||
//     TMA store of tile i-1 (shared -> global, bulk async group)
// so the threads never compute a global address, never predicate a load or a store (the tensor map's extent is the
// dealiasing cutoff: columns beyond it are zero-filled on load and dropped on store; masked rows are not transferred
// at all -- two boxes cover |k| <= kmax), and HBM latency is off the critical path.  The tile doubles as the exchange
// buffer of the radix passes.  Same arithmetic as col_fast_kernel<.., COL_PLAIN, ..>: bit-identical results.
#pragma once
#include <cuda.h>

#include "exb_kernels_nd_fast.cuh"

namespace exb {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct ColTmaParams {
  const cpx<float>* tw;   // N roots followed by the arranged register-FFT table (ColParams::tw)
  unsigned nblk;          // tiles in total = ntx * planes
  unsigned ntx;           // tiles along the last axis (kept wavenumbers only)
  int kmax;               // dealiasing cutoff (>= 0)
  int prune_in_rows;      // 1: only the line entries |k| <= kmax are loaded (the rest are zeros by contract)
  int prune_out_rows;     // 1: only the line entries |k| <= kmax are stored
};

// in_map / out_map: rank-3 maps {last axis (extent = kept columns), line axis (N), planes}; box {TW, BH, 1} with
// BH = N (all entries of a line) or kmax + 1 (pruned: two boxes, rows [0, kmax] and [N - kmax - 1, N - 1])
template <int N, int TW, int DIR>
__global__ void __launch_bounds__((N / 8) * TW, 2)
col_plain_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                     const ColTmaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  static_assert(TW == 16 && N <= 256, "one 128-byte row per line entry, box height <= 256");
  constexpr int P = N / 8;
  constexpr unsigned TILE_BYTES = N * TW * sizeof(cpx<float>);
  // TMA destinations must be 128-byte aligned: round the dynamic shared-memory base up (the launcher adds the slack)
  unsigned char* sbase = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  cpx<float>* tiles = reinterpret_cast<cpx<float>*>(sbase);                       // 2 x [N][TW]
  cpx<float>* tw = tiles + 2 * N * TW;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tw + ((Fft8Tw<N>::SIZE + 1) & ~1));
  Fft8Tw<N>::fill(tw, p.tw);
  const int w = threadIdx.x % TW, j = threadIdx.x / TW;
  const cpx<float> zero(0.f, 0.f);
  const int kmax = p.kmax;
  unsigned rowmask = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int ak = q < 4 ? j + P * q : N - (j + P * q);
    rowmask |= (ak <= kmax ? 1u : 0u) << q;
  }
  const unsigned in_mask = p.prune_in_rows ? rowmask : 0xffu;
  const int bh = kmax + 1;                                   // box height of a pruned transfer
  const unsigned in_bytes = p.prune_in_rows ? 2u * bh * TW * sizeof(cpx<float>) : TILE_BYTES;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_load = [&](unsigned blk, int s) {                // thread 0 only
    const int x = (int)(blk % p.ntx) * (2 * TW), z = (int)(blk / p.ntx);
    cpx<float>* dst = tiles + s * N * TW;
    mbar_expect_tx(&bars[s], in_bytes);
    if (p.prune_in_rows) {
      tma_load_3d(dst, &in_map, &bars[s], x, 0, z);
      tma_load_3d(dst + (N - bh) * TW, &in_map, &bars[s], x, N - bh, z);
    } else {
      tma_load_3d(dst, &in_map, &bars[s], x, 0, z);
    }
  };
  if (threadIdx.x == 0 && blockIdx.x < p.nblk) issue_load(blockIdx.x, 0);
  unsigned phase0 = 0, phase1 = 0;
  int s = 0;
  for (unsigned blk = blockIdx.x; blk < p.nblk; blk += gridDim.x, s ^= 1) {
    if (threadIdx.x == 0 && blk + gridDim.x < p.nblk) {
      tma_wait_read_all();                                    // the store that last read buffer s^1 has left shared memory
      issue_load(blk + gridDim.x, s ^ 1);
    }
    mbar_wait(&bars[s], s ? phase1 : phase0);
    if (s) phase1 ^= 1; else phase0 ^= 1;
    cpx<float>* tile = tiles + s * N * TW;
    ExTile<TW> ex{tile + w};
    cpx<float> v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = ((in_mask >> q) & 1u) ? tile[(j + P * q) * TW + w] : zero;
    fft8_run<N, DIR>(v, ex, j, tw);
    __syncthreads();                                          // every exchange read of the last pass is done
#pragma unroll
    for (int q = 0; q < 8; ++q) tile[(j + P * q) * TW + w] = v[q];
    fence_async_smem();                                       // generic-proxy writes -> visible to the TMA engine
    __syncthreads();
    if (threadIdx.x == 0) {
      const int x = (int)(blk % p.ntx) * (2 * TW), z = (int)(blk / p.ntx);
      if (p.prune_out_rows) {
        tma_store_3d(&out_map, tile, x, 0, z);
        tma_store_3d(&out_map, tile + (N - bh) * TW, x, N - bh, z);
      } else {
        tma_store_3d(&out_map, tile, x, 0, z);
      }
      tma_commit();
    }
  }
  if (threadIdx.x == 0) tma_wait_all();
}

}  // namespace exb
