// Shared body of exb_fastnd_n{256,512}.cu: instantiations + launchers for one line length.
#include <cstdlib>

#include "exb_fastnd.h"
#include "exb_kernels_nd_fast.cuh"
#include "exb_kernels_nd_tma.cuh"
#include "exb_tma.h"
// 16-points-per-thread row pass for N = 256 (one shared-memory exchange per transform, exb_row16.cuh): the row pass
// is bound by the shared-memory pipe; measured on c4 +4.5 % (round-1 code, r02_first) / +1.6 % (final code, r02s)
#ifndef EXB_ROW16
#define EXB_ROW16 1
#endif
#if EXB_ROW16
#include "exb_row16.cuh"
#endif

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

template <int N, int TW, class S, int NFWD, int MODE, int DIR, int STG = 0, int FUSE = 0>
int launch_col_stg(cudaStream_t st, const ColParams<float>& p, long long grid, const char** err) {
  constexpr bool usm = MODE == COL_INV_PRO && S::C == 1 && EXB_INVPRO_SMEM_U != 0;  // parked stage input
  constexpr int nstash = FUSE ? 1 : (usm ? (S::kind == EXB_NL_VORTICITY_2D ? 2 : 1) : 0);  // (+ the stream function)
  const size_t smem = (size_t)(Fft8Tw<N>::SIZE + ExTile<TW>::rows(N) * TW + nstash * N * TW) * sizeof(cpx<float>);
  // (set on every launch: the attribute is per device, and a process may drive several devices)
  if (int rc = set_smem(col_fast_kernel<N, TW, S, NFWD, MODE, DIR, STG, FUSE>, smem, err)) return rc;
  ColParams<float> q = p;
  q.TW = TW;
  col_fast_kernel<N, TW, S, NFWD, MODE, DIR, STG, FUSE><<<(unsigned)grid, (N / 8) * TW, smem, st>>>(q);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return EXB_ECUDA;
  }
  return EXB_OK;
}
#ifndef EXB_PLAIN_TMA
#define EXB_PLAIN_TMA 1
#endif
// TMA-staged persistent COL_PLAIN pass over axis 1 of a pitch-padded 3-D field buffer (exb_kernels_nd_tma.cuh).
// Returns EXB_OK / an error, or 1 when the configuration does not qualify (caller falls back to col_fast_kernel).
template <int N, int TW, int DIR>
int try_launch_plain_tma(cudaStream_t st, const ColParams<float>& p, const char** err) {
  if constexpr (N > 256 || TW != 16) {
    return 1;
  } else {
    static const bool off = getenv("EXB_PLAIN_TMA") && atoi(getenv("EXB_PLAIN_TMA")) == 0;
    const int kmax = p.P.kmax;
    const long long pitch = p.line_stride;
    if (off || !EXB_PLAIN_TMA || p.seg_len > 0 || p.peer || p.in != p.out || !(p.prune & PRUNE_COLS) || kmax < 0 ||
        2 * (kmax + 1) > N || pitch % 16 != 0 || p.outer_stride != (long long)N * pitch ||
        p.M != (long long)N * p.outer_stride || p.inner != p.P.Nh || p.n_outer != N)
      return 1;
    const bool pin = (p.prune & PRUNE_IN_ROWS) != 0, pout = (p.prune & PRUNE_OUT_ROWS) != 0;
    const uint64_t planes = (uint64_t)N * (uint64_t)p.batch * (uint64_t)p.nfields;
    const uint64_t dims[3] = {(uint64_t)kmax + 1, (uint64_t)N, planes};
    const uint64_t strides[3] = {0, (uint64_t)pitch * sizeof(cpx<float>), (uint64_t)p.outer_stride * sizeof(cpx<float>)};
    const uint32_t box_in[3] = {(uint32_t)TW, (uint32_t)(pin ? kmax + 1 : N), 1};
    const uint32_t box_out[3] = {(uint32_t)TW, (uint32_t)(pout ? kmax + 1 : N), 1};
    CUtensorMap in_map, out_map;
    if (exb_tma_encode(&in_map, p.in, 3, dims, strides, box_in, err)) return EXB_ECUDA;
    if (exb_tma_encode(&out_map, p.out, 3, dims, strides, box_out, err)) return EXB_ECUDA;
    ColTmaParams q;
    q.tw = p.tw;
    q.ntx = (unsigned)((kmax + 1 + TW - 1) / TW);
    const uint64_t nblk = (uint64_t)q.ntx * planes;
    if (nblk >= (1ull << 32)) return 1;
    q.nblk = (unsigned)nblk;
    q.kmax = kmax;
    q.prune_in_rows = pin;
    q.prune_out_rows = pout;
    const size_t smem = (size_t)(2 * N * TW + ((Fft8Tw<N>::SIZE + 1) & ~1)) * sizeof(cpx<float>) + 16 + 128;
    if (int rc = set_smem(col_plain_tma_kernel<N, TW, DIR>, smem, err)) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(nblk < 2ull * sms ? nblk : 2ull * sms);
    col_plain_tma_kernel<N, TW, DIR><<<grid, (N / 8) * TW, smem, st>>>(in_map, out_map, q);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      *err = cudaGetErrorString(e);
      return EXB_ECUDA;
    }
    return EXB_OK;
  }
}

#ifndef EXB_INVPRO_PERSISTENT
#define EXB_INVPRO_PERSISTENT 1
#endif
// persistent prologue pass of the one-channel 2-D kinds: only the tiles inside the dealiasing mask are enumerated
template <int N, int TW, class S, int DIR>
int launch_invpro_persistent(cudaStream_t st, const ColParams<float>& p, const char** err) {
  const size_t smem = (size_t)(Fft8Tw<N>::SIZE + ExTile<TW>::rows(N) * TW + 2 * N * TW) * sizeof(cpx<float>);
  if (int rc = set_smem(col_invpro_persistent_kernel<N, TW, S, DIR>, smem, err)) return rc;
  const long long cols = p.P.kmax >= 0 ? (long long)p.P.kmax + 1 : p.inner;      // kept last-axis wavenumbers
  const unsigned ntiles = (unsigned)((cols + TW - 1) / TW);
  const long long nblk = (long long)ntiles * p.batch;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long grid = nblk < 2LL * sms ? nblk : 2LL * sms;
  ColParams<float> q = p;
  q.TW = TW;
  col_invpro_persistent_kernel<N, TW, S, DIR><<<(unsigned)grid, (N / 8) * TW, smem, st>>>(q, (unsigned)nblk, ntiles);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return EXB_ECUDA;
  }
  return EXB_OK;
}

template <int N, int TW, class S, int NFWD, int MODE, int DIR>
int launch_col(cudaStream_t st, const ColParams<float>& p, long long grid, const char** err) {
  if constexpr (MODE == COL_INV_PRO && S::C == 1 && S::D == 2 && EXB_INVPRO_PERSISTENT != 0) {
    static const bool off = getenv("EXB_INVPRO_PERSISTENT") && atoi(getenv("EXB_INVPRO_PERSISTENT")) == 0;
    if (!off && !p.peer && (p.prune & PRUNE_COLS) && p.inner == p.P.Nh && p.M < (1LL << 31))
      return launch_invpro_persistent<N, TW, S, DIR>(st, p, err);
  }
  if constexpr (MODE == COL_PLAIN) {
    const int rc = try_launch_plain_tma<N, TW, DIR>(st, p, err);
    if (rc != 1) return rc;
  }
  if constexpr (MODE == COL_FWD_EPI) {
    // ETDRK2 (the default order): stage updates with compile-time operand sets
    if constexpr (S::C == 1 && S::D == 2) {
      // + the next evaluation's prologue pass fused in (p.fuse_next, set by the host when a step follows directly)
      if (p.fuse_next && p.K.order == 2 && p.stage == 0) return launch_col_stg<N, TW, S, NFWD, MODE, DIR, 1, 1>(st, p, grid, err);
      if (p.fuse_next && p.K.order == 2 && p.stage == 1) return launch_col_stg<N, TW, S, NFWD, MODE, DIR, 2, 1>(st, p, grid, err);
    }
    if (p.K.order == 2 && p.stage == 0) return launch_col_stg<N, TW, S, NFWD, MODE, DIR, 1>(st, p, grid, err);
    if (p.K.order == 2 && p.stage == 1) return launch_col_stg<N, TW, S, NFWD, MODE, DIR, 2>(st, p, grid, err);
  }
  return launch_col_stg<N, TW, S, NFWD, MODE, DIR, 0>(st, p, grid, err);
}

template <int N, class S, int NINV, int NFWD, int MODE>
int launch_row(cudaStream_t st, const RowParams<float>& p, const char** err) {
#if EXB_ROW16
  // experimental 16-points-per-thread row pass (one shared-memory exchange per transform), see exb_row16.cuh
  if constexpr (N == 256 && MODE == ROW_NL && row_streams<S, NINV, NFWD, MODE>()) {
    const size_t smem16 = row16_smem<S, NINV, NFWD>();
    if (int rc = set_smem(row16_kernel<S, NINV, NFWD>, smem16, err)) return rc;
    const long long np16 = (p.rows + 1) / 2 * p.batch;
    row16_kernel<S, NINV, NFWD><<<(unsigned)((np16 + 15) / 16), 256, smem16, st>>>(p);
    cudaError_t e16 = cudaGetLastError();
    if (e16 != cudaSuccess) {
      *err = cudaGetErrorString(e16);
      return EXB_ECUDA;
    }
    return EXB_OK;
  }
#endif
  constexpr int P = N / 8, GROUPS = 256 / P;
  const bool prefetch = MODE == ROW_NL && (EXB_ROW_PREFETCH != 0);  // staging rows, see row_fast_kernel
  const int nhp = (N / 2 + 1 + 7) / 8 * 8;
  constexpr int kst = row_streams<S, NINV, NFWD, MODE>() ? row_stream_stash<S>() : 0;  // parked fields
  const size_t smem =
      (size_t)(Fft8Tw<N>::SIZE + GROUPS * (N + N / 8 + kst * N + (prefetch ? 2 * nhp : 0))) * sizeof(cpx<float>);
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

// kinds with fast instantiations (bit mask per translation unit)
enum : int { K_VORT = 1, K_PROJ = 2, K_GRAD2 = 4, K_POLY2 = 8, K_CONV2 = 16 };
using SGrad2 = NlS<EXB_NL_GRADIENT_NORM, -1, 2, 1>;   // 2-D gradient norm, 1 channel (KS 2-D): 2 inverse, 1 forward
using SPoly2 = NlS<EXB_NL_POLYNOMIAL, -1, 2, 1>;      // 2-D polynomial, 1 channel (reaction-diffusion): 1 / 1
using SConv2 = NlS<EXB_NL_CONVECTION, 0, 2, 2>;       // 2-D convection, 2 channels, non-conservative (Burgers): 6 / 2

// which fast kind (0: none) does this nonlinear function map to
inline int fast_kind_of(const NlParams<float>& P) {
  if (P.D == 2) {
    if (P.kind == EXB_NL_VORTICITY_2D) return K_VORT;
    if (P.kind == EXB_NL_GRADIENT_NORM && P.C == 1) return K_GRAD2;
    if (P.kind == EXB_NL_POLYNOMIAL && P.C == 1) return K_POLY2;
    if (P.kind == EXB_NL_CONVECTION && !P.single_channel && !P.conservative && P.C == 2) return K_CONV2;
  }
  if (P.D == 3 && P.kind == EXB_NL_PROJECTED_3D) return K_PROJ;
  return 0;
}

template <int N, int TW, int KINDS, int MODE, int DIR>
int col_mode_dispatch(cudaStream_t st, const ColParams<float>& p, long long grid, const char** err) {
  const int fk = fast_kind_of(p.P);
  if constexpr ((KINDS & K_VORT) != 0) {
    if (fk == K_VORT) return launch_col<N, TW, SVort, 1, MODE, DIR>(st, p, grid, err);
  }
  if constexpr ((KINDS & K_PROJ) != 0) {
    if (fk == K_PROJ) return launch_col<N, TW, SProj, 3, MODE, DIR>(st, p, grid, err);
  }
  if constexpr ((KINDS & K_GRAD2) != 0) {
    if (fk == K_GRAD2) return launch_col<N, TW, SGrad2, 1, MODE, DIR>(st, p, grid, err);
  }
  if constexpr ((KINDS & K_POLY2) != 0) {
    if (fk == K_POLY2) return launch_col<N, TW, SPoly2, 1, MODE, DIR>(st, p, grid, err);
  }
  if constexpr ((KINDS & K_CONV2) != 0) {
    if (fk == K_CONV2) return launch_col<N, TW, SConv2, 2, MODE, DIR>(st, p, grid, err);
  }
  *err = "fast N-D column pass: unsupported configuration";
  return EXB_EUNSUPPORTED;
}

template <int N, int TW, int KINDS> int col_n(cudaStream_t st, const ColParams<float>& p, int dir, long long grid, const char** err) {
  switch (p.mode) {
    case COL_PLAIN:
      return dir < 0 ? launch_col<N, TW, SPlain, 1, COL_PLAIN, -1>(st, p, grid, err)
                     : launch_col<N, TW, SPlain, 1, COL_PLAIN, +1>(st, p, grid, err);
    case COL_INV_PRO:
      return col_mode_dispatch<N, TW, KINDS, COL_INV_PRO, +1>(st, p, grid, err);
    case COL_FWD_EPI:
      return col_mode_dispatch<N, TW, KINDS, COL_FWD_EPI, -1>(st, p, grid, err);
    case COL_FWD_NL:
      return col_mode_dispatch<N, TW, KINDS, COL_FWD_NL, -1>(st, p, grid, err);
  }
  *err = "fast N-D column pass: unsupported configuration";
  return EXB_EUNSUPPORTED;
}

template <int N, int KINDS> int row_n(cudaStream_t st, const RowParams<float>& p, const char** err) {
  switch (p.mode) {
    case ROW_R2C:
      return launch_row<N, SPlain, 1, 1, ROW_R2C>(st, p, err);
    case ROW_C2R:
      return launch_row<N, SPlain, 1, 1, ROW_C2R>(st, p, err);
    case ROW_NL: {
      const int fk = fast_kind_of(p.P);
      if constexpr ((KINDS & K_VORT) != 0) {
        if (fk == K_VORT) return launch_row<N, SVort, 4, 1, ROW_NL>(st, p, err);
      }
      if constexpr ((KINDS & K_PROJ) != 0) {
        if (fk == K_PROJ) return launch_row<N, SProj, 6, 3, ROW_NL>(st, p, err);
      }
      if constexpr ((KINDS & K_GRAD2) != 0) {
        if (fk == K_GRAD2) return launch_row<N, SGrad2, 2, 1, ROW_NL>(st, p, err);
      }
      if constexpr ((KINDS & K_POLY2) != 0) {
        if (fk == K_POLY2) return launch_row<N, SPoly2, 1, 1, ROW_NL>(st, p, err);
      }
      if constexpr ((KINDS & K_CONV2) != 0) {
        if (fk == K_CONV2) return launch_row<N, SConv2, 6, 2, ROW_NL>(st, p, err);
      }
      break;
    }
  }
  *err = "fast N-D row pass: unsupported configuration";
  return EXB_EUNSUPPORTED;
}

}  // namespace
