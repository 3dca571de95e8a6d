// N = 2048 instantiations of the fast 2-D / 3-D pass kernels.
#include "exb_fastnd_impl.cuh"

int exb_fastnd_col_n2048(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  return col_n<2048, 2, K_PROJ>(st, p, dir, grid, err);
}
int exb_fastnd_row_n2048(cudaStream_t st, const RowParams<float>& p, const char** err) {
  return row_n<2048, K_PROJ>(st, p, err);
}
