// N = 256 (R = 16) instantiations of the fast 1-D persistent rollout kernel.
#include "exb_fast1d_impl.cuh"

int exb_launch_fast1d_r16(cudaStream_t st, const K1dParams<float>& p, int nscr, int max_smem, const char** err) {
  return launch_fast_r<16>(st, p, nscr, max_smem, err);
}

bool exb_fast1d_supported(int N, const NlParams<float>& P, int order) {
  if (N != 256 && N != 64) return false;
  if (P.D != 1 || P.C != 1) return false;
  if (P.has_inj) return false;   // (no 1-D stepper injects; the generic kernel handles it)
  if (order == 0) return true;
  switch (P.kind) {
    case EXB_NL_CONVECTION:
    case EXB_NL_GRADIENT_NORM:
    case EXB_NL_POLYNOMIAL:
    case EXB_NL_GENERAL:
      return true;
    default:
      return false;
  }
}
