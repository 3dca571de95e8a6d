// N = 128 instantiations of the fast 2-D / 3-D pass kernels.
#include "exb_fastnd_impl.cuh"

int exb_fastnd_col_n128(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  return col_n<128, 16, K_VORT | K_PROJ | K_GRAD2 | K_POLY2 | K_CONV2>(st, p, dir, grid, err);
}
int exb_fastnd_row_n128(cudaStream_t st, const RowParams<float>& p, const char** err) {
  return row_n<128, K_VORT | K_PROJ | K_GRAD2 | K_POLY2 | K_CONV2>(st, p, err);
}
