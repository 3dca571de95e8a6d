// Shared body of exb_fastnd_n{256,512}.cu: instantiations + launchers for one line length.
#include "exb_fastnd.h"
#include "exb_kernels_nd_fast.cuh"

using namespace exb;

namespace {

using SVort = NlS<EXB_NL_VORTICITY_2D, -1, 2, 1>;
using SProj = NlS<EXB_NL_PROJECTED_3D, -1, 3, 3>;
using SPlain = NlS<EXB_NL_POLYNOMIAL, -1, 1, 1>;

template <class K> int set_smem(K kernel, size_t smem, const char** err) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return EXB_ECUDA;
  }
  return EXB_OK;
}

template <int N, int TW, class S, int NFWD, int MODE, int DIR>
int launch_col(cudaStream_t st, const ColParams<float>& p, long long grid, const char** err) {
  const size_t smem = (size_t)(Fft8Tw<N>::SIZE + ExTile<TW>::rows(N) * TW) * sizeof(cpx<float>);
  // (set on every launch: the attribute is per device, and a process may drive several devices)
  if (int rc = set_smem(col_fast_kernel<N, TW, S, NFWD, MODE, DIR>, smem, err)) return rc;
  ColParams<float> q = p;
  q.TW = TW;
  col_fast_kernel<N, TW, S, NFWD, MODE, DIR><<<(unsigned)grid, (N / 8) * TW, smem, st>>>(q);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return EXB_ECUDA;
  }
  return EXB_OK;
}

template <int N, class S, int NINV, int NFWD, int MODE>
int launch_row(cudaStream_t st, const RowParams<float>& p, const char** err) {
  constexpr int P = N / 8, GROUPS = 256 / P;
  const bool stash = false;  // see row_fast_kernel: STASH is compiled out
  const size_t smem = (size_t)(Fft8Tw<N>::SIZE + GROUPS * (N + N / 8 + (stash ? NINV * N : 0))) * sizeof(cpx<float>);
  if (int rc = set_smem(row_fast_kernel<N, S, NINV, NFWD, MODE, GROUPS>, smem, err)) return rc;
  const long long npairs = (p.rows + 1) / 2 * p.batch;
  const long long grid = (npairs + GROUPS - 1) / GROUPS;
  row_fast_kernel<N, S, NINV, NFWD, MODE, GROUPS><<<(unsigned)grid, P * GROUPS, smem, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return EXB_ECUDA;
  }
  return EXB_OK;
}

template <int N, int TW, bool VORT = true, bool PROJ = true> int col_n(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  const int kind = p.P.kind;
  switch (p.mode) {
    case COL_PLAIN:
      return dir < 0 ? launch_col<N, TW, SPlain, 1, COL_PLAIN, -1>(st, p, grid, err)
                     : launch_col<N, TW, SPlain, 1, COL_PLAIN, +1>(st, p, grid, err);
    case COL_INV_PRO:
      if constexpr (VORT) {
        if (kind == EXB_NL_VORTICITY_2D) return launch_col<N, TW, SVort, 1, COL_INV_PRO, +1>(st, p, grid, err);
      }
      if constexpr (PROJ) {
        if (kind == EXB_NL_PROJECTED_3D) return launch_col<N, TW, SProj, 3, COL_INV_PRO, +1>(st, p, grid, err);
      }
      break;
    case COL_FWD_EPI:
      if constexpr (VORT) {
        if (kind == EXB_NL_VORTICITY_2D) return launch_col<N, TW, SVort, 1, COL_FWD_EPI, -1>(st, p, grid, err);
      }
      if constexpr (PROJ) {
        if (kind == EXB_NL_PROJECTED_3D) return launch_col<N, TW, SProj, 3, COL_FWD_EPI, -1>(st, p, grid, err);
      }
      break;
    case COL_FWD_NL:
      if constexpr (VORT) {
        if (kind == EXB_NL_VORTICITY_2D) return launch_col<N, TW, SVort, 1, COL_FWD_NL, -1>(st, p, grid, err);
      }
      if constexpr (PROJ) {
        if (kind == EXB_NL_PROJECTED_3D) return launch_col<N, TW, SProj, 3, COL_FWD_NL, -1>(st, p, grid, err);
      }
      break;
  }
  *err = "fast N-D column pass: unsupported configuration";
  return EXB_EUNSUPPORTED;
}

template <int N, bool VORT = true, bool PROJ = true> int row_n(cudaStream_t st, const RowParams<float>& p, const char** err) {
  switch (p.mode) {
    case ROW_R2C:
      return launch_row<N, SPlain, 1, 1, ROW_R2C>(st, p, err);
    case ROW_C2R:
      return launch_row<N, SPlain, 1, 1, ROW_C2R>(st, p, err);
    case ROW_NL:
      if constexpr (VORT) {
        if (p.P.kind == EXB_NL_VORTICITY_2D) return launch_row<N, SVort, 4, 1, ROW_NL>(st, p, err);
      }
      if constexpr (PROJ) {
        if (p.P.kind == EXB_NL_PROJECTED_3D) return launch_row<N, SProj, 6, 3, ROW_NL>(st, p, err);
      }
      break;
  }
  *err = "fast N-D row pass: unsupported configuration";
  return EXB_EUNSUPPORTED;
}

}  // namespace
