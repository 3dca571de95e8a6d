// N = 256 instantiations of the fast 2-D / 3-D pass kernels.
#include "exb_fastnd_impl.cuh"

int exb_fastnd_col_n256(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  return col_n<256, 16, K_VORT | K_PROJ | K_GRAD2 | K_POLY2 | K_CONV2>(st, p, dir, grid, err);
}
int exb_fastnd_row_n256(cudaStream_t st, const RowParams<float>& p, const char** err) {
  return row_n<256, K_VORT | K_PROJ | K_GRAD2 | K_POLY2 | K_CONV2>(st, p, err);
}

int exb_fastnd_col_n128(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err);
int exb_fastnd_row_n128(cudaStream_t st, const RowParams<float>& p, const char** err);
int exb_fastnd_col_n512(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err);
int exb_fastnd_row_n512(cudaStream_t st, const RowParams<float>& p, const char** err);
int exb_fastnd_col_n1024(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err);
int exb_fastnd_row_n1024(cudaStream_t st, const RowParams<float>& p, const char** err);
int exb_fastnd_col_n2048(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err);
int exb_fastnd_row_n2048(cudaStream_t st, const RowParams<float>& p, const char** err);

template <int N> static void tw_arrange_t(const cpx<float>* roots, cpx<float>* out) {
  for (int q = 0; q < Fft8Tw<N>::SIZE; ++q) out[q] = roots[Fft8Tw<N>::root_index(q)];
}
int exb_fastnd_tw_size(int N) {
  switch (N) {
    case 128: return Fft8Tw<128>::SIZE;
    case 256: return Fft8Tw<256>::SIZE;
    case 512: return Fft8Tw<512>::SIZE;
    case 1024: return Fft8Tw<1024>::SIZE;
    case 2048: return Fft8Tw<2048>::SIZE;
  }
  return 0;
}
void exb_fastnd_tw_arrange(int N, const cpx<float>* roots, cpx<float>* out) {
  switch (N) {
    case 128: return tw_arrange_t<128>(roots, out);
    case 256: return tw_arrange_t<256>(roots, out);
    case 512: return tw_arrange_t<512>(roots, out);
    case 1024: return tw_arrange_t<1024>(roots, out);
    case 2048: return tw_arrange_t<2048>(roots, out);
  }
}

bool exb_fastnd_supported(int D, int N, const NlParams<float>& P) {
  const int fk = fast_kind_of(P);
  if (fk == 0) return false;
  if (fk == K_PROJ) return N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048;
  return N == 128 || N == 256 || N == 512;
}

int exb_fastnd_col(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  switch (p.fd.N) {
    case 128: return exb_fastnd_col_n128(st, p, dir, grid, err);
    case 256: return exb_fastnd_col_n256(st, p, dir, grid, err);
    case 512: return exb_fastnd_col_n512(st, p, dir, grid, err);
    case 1024: return exb_fastnd_col_n1024(st, p, dir, grid, err);
    case 2048: return exb_fastnd_col_n2048(st, p, dir, grid, err);
  }
  *err = "fast N-D column pass: unsupported N";
  return EXB_EUNSUPPORTED;
}

int exb_fastnd_row(cudaStream_t st, const RowParams<float>& p, const char** err) {
  switch (p.fd.N) {
    case 128: return exb_fastnd_row_n128(st, p, err);
    case 256: return exb_fastnd_row_n256(st, p, err);
    case 512: return exb_fastnd_row_n512(st, p, err);
    case 1024: return exb_fastnd_row_n1024(st, p, err);
    case 2048: return exb_fastnd_row_n2048(st, p, err);
  }
  *err = "fast N-D row pass: unsupported N";
  return EXB_EUNSUPPORTED;
}
