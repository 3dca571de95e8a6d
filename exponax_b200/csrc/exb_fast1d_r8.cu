// N = 64 (R = 8) instantiations of the fast 1-D persistent rollout kernel.
#include "exb_fast1d_impl.cuh"

int exb_launch_fast1d_r8(cudaStream_t st, const K1dParams<float>& p, int nscr, int max_smem, const char** err) {
  return launch_fast_r<8>(st, p, nscr, max_smem, err);
}
