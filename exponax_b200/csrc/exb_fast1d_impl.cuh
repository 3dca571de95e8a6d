// Shared body of exb_fast1d_r{8,16}.cu: instantiations + launcher for one radix EXB_FAST_R.
#include "exb_fast1d.h"
#include "exb_kernels_1d_fast.cuh"

using namespace exb;

namespace {

template <int R, class S, int NI, int NF>
int launch_fast_t(cudaStream_t st, const K1dParams<float>& p, int nscr, int max_smem, const char** err) {
  constexpr int NN = R * R, NNh = NN / 2 + 1;
  // CTA shape: 2 CTAs per SM, sized so that the pairs of one SM split into equal rounds
  // (all CTAs take the same time; a partially filled last round is pure loss).
  const int gpw = 32 / R;  // pair groups per warp
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static const int nstate_of_order[5] = {1, 1, 2, 4, 5};
  const int nstate = nstate_of_order[p.K.order];
  const long long npairs_all = (p.batch + 1) / 2;
  const int pair_bytes_est = (nstate * (NN + 2) + (R + 1) * R) * 8;   // packed state arrays + exchange buffer
  // shared memory per SM: 228 KB minus 2 x (coefficient/twiddle tables + 1 KB system reserve)
  int max_pairs_sm = (228 * 1024 - 2 * (R * R * 8 + NNh * 48 + 256 + 1024)) / pair_bytes_est;
  max_pairs_sm -= max_pairs_sm % (2 * gpw);
  if (max_pairs_sm > 16 * gpw) max_pairs_sm = 16 * gpw;                 // register bound: 16 warps/SM
  if (max_pairs_sm < 2 * gpw) max_pairs_sm = 2 * gpw;
  long long pairs_per_sm = (npairs_all + sms - 1) / sms;
  long long rounds = (pairs_per_sm + max_pairs_sm - 1) / max_pairs_sm;
  long long pairs_per_round = (pairs_per_sm + rounds - 1) / rounds;     // per SM
  int warps = (int)((pairs_per_round + 2 * gpw - 1) / (2 * gpw));       // 2 CTAs per SM
  if (warps < 1) warps = 1;
  if (warps > 8) warps = 8;
  const int groups = warps * gpw;
  FastLayout lay;
  int off = 0;
  auto take = [&](int bytes) {
    int o = off;
    off += (bytes + 15) / 16 * 16;
    return o;
  };
  lay.off_tw2 = take(R * R * 8);
  lay.off_exp = take(NNh * 8);
  lay.off_hexp = take(NNh * 8);
  for (int i = 0; i < 6; ++i) lay.off_c[i] = take(NNh * 4);
  lay.off_mk = take(NNh * 8);
  lay.nstate = nstate;
  (void)nscr;
  lay.pair_bytes = (lay.nstate * (NN + 2) + (R + 1) * R) * 8;
  lay.off_pairs = take(0);
  size_t smem = (size_t)off + (size_t)groups * lay.pair_bytes;
  if ((long long)smem > max_smem) {
    *err = "fast 1-D kernel: not enough shared memory";
    return EXB_EUNSUPPORTED;
  }
  long long npairs = (p.batch + 1) / 2;
  long long grid = (npairs + groups - 1) / groups;
  // ETDRK2 (the reference's default order) has its own instance: no order dispatch, state slots handed over in registers
  auto launch = [&](auto kernel) -> cudaError_t {
    // per device attribute: set on every launch (a process may drive several devices)
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)grid, warps * 32, smem, st>>>(p, lay);
    return cudaSuccess;
  };
  {
    cudaError_t e = p.K.order == 2 ? launch(k1d_fast_kernel<R, S, NI, NF, 2>) : launch(k1d_fast_kernel<R, S, NI, NF, 0>);
    if (e != cudaSuccess) {
      *err = cudaGetErrorString(e);
      return EXB_ECUDA;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return EXB_ECUDA;
  }
  return EXB_OK;
}

template <int R>
int launch_fast_r(cudaStream_t st, const K1dParams<float>& p, int nscr, int max_smem, const char** err) {
  const NlParams<float>& P = p.P;
  const int kind = p.K.order == 0 ? EXB_NL_POLYNOMIAL : P.kind;  // order 0 never evaluates N
  switch (kind) {
    case EXB_NL_CONVECTION:
      if (P.conservative) return launch_fast_t<R, NlS<EXB_NL_CONVECTION, 3, 1, 1>, 1, 1>(st, p, nscr, max_smem, err);
      return launch_fast_t<R, NlS<EXB_NL_CONVECTION, 0, 1, 1>, 2, 1>(st, p, nscr, max_smem, err);
    case EXB_NL_GRADIENT_NORM:
      return launch_fast_t<R, NlS<EXB_NL_GRADIENT_NORM, -1, 1, 1>, 1, 1>(st, p, nscr, max_smem, err);
    case EXB_NL_POLYNOMIAL:
    case EXB_NL_ZERO:
      return launch_fast_t<R, NlS<EXB_NL_POLYNOMIAL, -1, 1, 1>, 1, 1>(st, p, nscr, max_smem, err);
    case EXB_NL_GENERAL:
      return launch_fast_t<R, NlS<EXB_NL_GENERAL, -1, 1, 1>, 2, 2>(st, p, nscr, max_smem, err);
    default:
      *err = "fast 1-D kernel: unsupported nonlinear function";
      return EXB_EUNSUPPORTED;
  }
}

}  // namespace
