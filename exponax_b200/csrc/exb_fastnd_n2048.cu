// N = 2048 instantiations (3-D projected convection + plain passes: BASELINE config c5).
#include "exb_fastnd_impl.cuh"

int exb_fastnd_col_n2048(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  return col_n<2048, 2, false, true>(st, p, dir, grid, err);
}
int exb_fastnd_row_n2048(cudaStream_t st, const RowParams<float>& p, const char** err) {
  return row_n<2048, false, true>(st, p, err);
}
