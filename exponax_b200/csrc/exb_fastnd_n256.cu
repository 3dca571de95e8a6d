// N = 256 instantiations of the fast 2-D / 3-D pass kernels + the size dispatcher.
#include "exb_fastnd_impl.cuh"

int exb_fastnd_col_n512(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err);
int exb_fastnd_row_n512(cudaStream_t st, const RowParams<float>& p, const char** err);
int exb_fastnd_col_n1024(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err);
int exb_fastnd_row_n1024(cudaStream_t st, const RowParams<float>& p, const char** err);
int exb_fastnd_col_n2048(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err);
int exb_fastnd_row_n2048(cudaStream_t st, const RowParams<float>& p, const char** err);

bool exb_fastnd_supported(int D, int N, const NlParams<float>& P) {
  if (D == 2) return (N == 256 || N == 512) && P.kind == EXB_NL_VORTICITY_2D;
  if (D == 3) return (N == 256 || N == 512 || N == 1024 || N == 2048) && P.kind == EXB_NL_PROJECTED_3D;
  return false;
}

int exb_fastnd_col(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  if (p.fd.N == 512) return exb_fastnd_col_n512(st, p, dir, grid, err);
  if (p.fd.N == 1024) return exb_fastnd_col_n1024(st, p, dir, grid, err);
  if (p.fd.N == 2048) return exb_fastnd_col_n2048(st, p, dir, grid, err);
  return col_n<256, 16>(st, p, dir, grid, err);
}

int exb_fastnd_row(cudaStream_t st, const RowParams<float>& p, const char** err) {
  if (p.fd.N == 512) return exb_fastnd_row_n512(st, p, err);
  if (p.fd.N == 1024) return exb_fastnd_row_n1024(st, p, err);
  if (p.fd.N == 2048) return exb_fastnd_row_n2048(st, p, err);
  return row_n<256>(st, p, err);
}
