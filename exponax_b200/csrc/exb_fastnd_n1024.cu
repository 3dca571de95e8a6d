// N = 1024 instantiations of the fast 2-D / 3-D pass kernels.
#include "exb_fastnd_impl.cuh"

int exb_fastnd_col_n1024(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  return col_n<1024, 4, K_PROJ>(st, p, dir, grid, err);
}
int exb_fastnd_row_n1024(cudaStream_t st, const RowParams<float>& p, const char** err) {
  return row_n<1024, K_PROJ>(st, p, err);
}
