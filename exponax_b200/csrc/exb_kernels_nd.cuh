// 2-D / 3-D path.  A D-dimensional real transform is (D-1) strided "column" passes over the
// outer axes (complex FFTs on [N x TW] tiles staged in shared memory) plus one "row" pass over
// the contiguous last axis.  Everything that is not a butterfly is fused into those passes:
//   * inverse column pass over axis 0  : prologue = dealias mask, i*k, inverse Laplacian, curl
//   * row pass                         : c2r -> pointwise nonlinearity -> r2c, two rows per
//                                        complex FFT (two-for-one), physical fields never in HBM
//   * forward column pass over axis 0  : epilogue = dealias mask, scale, Leray projection,
//                                        injection and the ETDRK stage update
// Reference semantics: exponax/nonlin_fun/_base.py:99-137 (mask placement), exponax/_spectral.py
// :614-721 (rfftn / irfftn), exponax/etdrk/*.py (stage updates).
#pragma once
#include "exb_fft.cuh"
#include "exb_kernels_1d.cuh"  // pack2 / unpack2
#include "exb_nl.cuh"

namespace exb {

enum ColMode { COL_PLAIN = 0, COL_INV_PRO = 1, COL_FWD_EPI = 2, COL_FWD_NL = 3 };
enum RowMode { ROW_NL = 0, ROW_R2C = 1, ROW_C2R = 2 };
// Dealiasing-aware pruning (2/3 rule: about 1/3 of the modes per masked axis are identically
// zero before an inverse / discarded after a forward transform):
enum Prune {
  PRUNE_COLS = 1,      // skip lines whose fixed wavenumbers lie outside the mask
  PRUNE_IN_ROWS = 2,   // entries of a line outside the mask are read as zero (no load)
  PRUNE_OUT_ROWS = 4   // entries of a line outside the mask are not stored
};

template <class T> struct ColParams {
  NlParams<T> P;
  EtdrkCoefs<T> K;
  FftDesc fd;
  const cpx<T>* tw;
  int mode;
  int TW;                  // tile width (lines per CTA)
  int nfields;             // fields per batch element (PLAIN); n_inv / n_fwd otherwise
  int stage;               // ETDRK stage (FWD_EPI)
  int prune;               // fast kernels only: PRUNE_* bits (dealiased modes are never touched)
  int f0, fcount;          // COL_INV_PRO: inverse fields [f0, f0 + fcount) (fcount <= 0: all)
  int batch_fastest;       // fast COL_FWD_EPI: block index = tile * batch + trajectory (tables larger than the L2)
  // Fast kernels: the FIELD buffers between the passes of one N(u) evaluation (n_inv inverse / n_fwd forward fields)
  // may have a padded last-axis pitch (fpitch >= N/2+1 complex elements, a multiple of 16 = 128 B: every tile row
  // is line-aligned and a tensor map can describe the buffer) and their own field stride fM = N^(D-1) * fpitch.
  // State buffers and coefficient tables are always dense (pitch N/2+1, stride M).  0 = dense (generic kernels, slabs).
  int fpitch;
  long long fM;
  int masked_external;     // fast COL_FWD_EPI: the dealiased modes are advanced by etdrk_masked_linear_kernel, skip them
  int fuse_next;           // fast COL_FWD_EPI (ETDRK2, one-channel 2-D kinds): also run the NEXT evaluation's prologue
                           //   pass on the value this pass produces; `out` = inverse-field buffer
  int seg_len;             // COL_PLAIN, slab transposes without pack/unpack: line entry i lives at
  long long seg_stride;    //   (i / seg_len) * seg_stride + (i % seg_len) * line_stride   (seg_len = 0: off)
  int seg_cyclic;          //   P > 0: cyclic axis-1 distribution, entry i at (i % P) * seg_stride + (i / P) * line_stride
  // peer output (fast kernels, COL_INV_PRO / COL_PLAIN of a slab plan): entry i of an output line is stored
  // into the buffer of rank i / seg_len -- peer_out[r] is that rank's destination buffer mapped into this
  // process (NVLink peer memory) -- at peer_off + (i % seg_len) * line_stride; the pass IS the transpose.
  int peer;
  long long peer_off;
  cpx<T>* peer_out[EXB_MAX_PEERS];
  long long line_stride;   // elements between successive points of a line
  long long inner;         // contiguous positions across which lines are tiled
  long long n_outer;       // independent slabs per field (3-D axis-1 pass: N)
  long long outer_stride;  // elements between slabs
  long long M;             // elements per field
  long long batch;
  const cpx<T>* in;        // PLAIN / FWD: fields; INV_PRO: stage input state (batch, C, M)
  cpx<T>* out;             // PLAIN / INV_PRO: fields; FWD_NL: N(u) (batch, C, M)
  StateBufs<T> sb;         // FWD_EPI
};

// decode the spatial indices of tile element (i along the line, iw along the inner direction);
// only used by the passes over axis 0 (prologue / epilogue), where inner = everything else.
template <class T>
__device__ __forceinline__ ModeK<T> col_mode(const NlParams<T>& P, int i, long long iw) {
  if (P.D == 2) return make_mode(P, i, (int)iw, 0);
  int i1 = (int)(iw / P.Nh);
  int i2 = (int)(iw - (long long)i1 * P.Nh);
  return make_mode(P, i, i1 * P.i1_mul + P.i1_off, i2);
}

template <class T, int DIR> __global__ void col_pass_kernel(const ColParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<T>* sm = reinterpret_cast<cpx<T>*>(smem_raw);
  const int N = p.fd.N, TW = p.TW;
  const size_t tile = (size_t)N * TW;
  const long long ntiles = (p.inner + TW - 1) / TW;
  long long bid = blockIdx.x;
  const long long t = bid % ntiles;
  bid /= ntiles;
  const long long w0 = t * TW;
  const int wv = (int)((p.inner - w0) < TW ? (p.inner - w0) : TW);
  const cpx<T> zero((T)0, (T)0);

  if (p.mode == COL_PLAIN) {
    const long long o = bid % p.n_outer;
    bid /= p.n_outer;
    const long long fb = bid;  // batch * nfields + field
    const size_t base = (size_t)fb * p.M + (size_t)o * p.outer_stride + w0;
    cpx<T>* A = sm;
    cpx<T>* B = sm + tile;
    auto line_off = [&](int i) -> size_t {
      if (p.seg_len > 0)
        return p.seg_cyclic ? (size_t)(i % p.seg_cyclic) * p.seg_stride + (size_t)(i / p.seg_cyclic) * p.line_stride
                            : (size_t)(i / p.seg_len) * p.seg_stride + (size_t)(i % p.seg_len) * p.line_stride;
      return (size_t)i * p.line_stride;
    };
    for (int q = threadIdx.x; q < N * TW; q += blockDim.x) {
      int i = q / TW, w = q - i * TW;
      A[q] = w < wv ? p.in[base + line_off(i) + w] : zero;
    }
    __syncthreads();
    cpx<T>* R = fft_lines<T, DIR>(A, B, TW, 1, TW, true, p.fd, p.tw);
    for (int q = threadIdx.x; q < N * TW; q += blockDim.x) {
      int i = q / TW, w = q - i * TW;
      if (w < wv) p.out[base + line_off(i) + w] = R[q];
    }
    return;
  }

  const NlParams<T>& P = p.P;
  const int C = P.C;
  const long long b = bid;

  if (p.mode == COL_INV_PRO) {
    cpx<T>* Ust = sm;                       // C tiles of the stage input
    cpx<T>* A = sm + (size_t)C * tile;
    cpx<T>* B = A + tile;
    for (int q = threadIdx.x; q < C * N * TW; q += blockDim.x) {
      int ch = q / (N * TW);
      int r = q - ch * (N * TW);
      int i = r / TW, w = r - i * TW;
      Ust[q] = w < wv ? p.in[((size_t)b * C + ch) * p.M + (size_t)i * p.line_stride + w0 + w] : zero;
    }
    __syncthreads();
    const int f_begin = p.fcount > 0 ? p.f0 : 0, f_end = p.fcount > 0 ? p.f0 + p.fcount : P.n_inv;
    for (int f = f_begin; f < f_end; ++f) {
      for (int q = threadIdx.x; q < N * TW; q += blockDim.x) {
        int i = q / TW, w = q - i * TW;
        cpx<T> v = zero;
        if (w < wv) {
          ModeK<T> m = col_mode(P, i, w0 + w);
          cpx<T> uh[EXB_MAXC];
#pragma unroll
          for (int ch = 0; ch < EXB_MAXC; ++ch) uh[ch] = ch < C ? Ust[(size_t)ch * tile + q] : zero;
          v = nl_inv_field(P, f, uh, m);
        }
        A[q] = v;
      }
      __syncthreads();
      cpx<T>* R = fft_lines<T, DIR>(A, B, TW, 1, TW, true, p.fd, p.tw);
      const size_t obase = ((size_t)b * P.n_inv + f) * p.M + w0;
      for (int q = threadIdx.x; q < N * TW; q += blockDim.x) {
        int i = q / TW, w = q - i * TW;
        if (w < wv) p.out[obase + (size_t)i * p.line_stride + w] = R[q];
      }
      __syncthreads();
    }
    return;
  }

  // COL_FWD_EPI / COL_FWD_NL: transform all n_fwd fields of this tile, then the per-mode epilogue
  cpx<T>* res[EXB_MAX_FWD];
  cpx<T>* scratch = sm + (size_t)P.n_fwd * tile;
  for (int g = 0; g < P.n_fwd; ++g) {
    cpx<T>* A = sm + (size_t)g * tile;
    const size_t ibase = ((size_t)b * P.n_fwd + g) * p.M + w0;
    for (int q = threadIdx.x; q < N * TW; q += blockDim.x) {
      int i = q / TW, w = q - i * TW;
      A[q] = w < wv ? p.in[ibase + (size_t)i * p.line_stride + w] : zero;
    }
  }
  __syncthreads();
  for (int g = 0; g < P.n_fwd; ++g) {
    cpx<T>* A = sm + (size_t)g * tile;
    // the scratch of field g is whatever buffer is currently free
    cpx<T>* R = fft_lines<T, DIR>(A, scratch, TW, 1, TW, true, p.fd, p.tw);
    if (R != A) {
      // result landed in the scratch: that buffer now belongs to field g, A becomes the scratch.
      // Keep the slot order simple by remembering the pointer.
      res[g] = R;
      scratch = A;
    } else {
      res[g] = A;
    }
  }
  for (int q = threadIdx.x; q < N * TW; q += blockDim.x) {
    int i = q / TW, w = q - i * TW;
    if (w >= wv) continue;
    ModeK<T> m = col_mode(P, i, w0 + w);
    cpx<T> W[EXB_MAX_FWD], n[EXB_MAXC];
    for (int g = 0; g < P.n_fwd; ++g) W[g] = res[g][q];
    nl_from_fwd(P, W, m, n);
    const long long mode = (long long)i * p.line_stride + w0 + w;
#pragma unroll
    for (int ch = 0; ch < EXB_MAXC; ++ch) {
      if (ch < C) {
        const size_t off = ((size_t)b * C + ch) * p.M + mode;
        if (p.mode == COL_FWD_NL) {
          p.out[off] = n[ch];
        } else {
          const long long ci = (long long)(p.K.E == 1 ? 0 : ch) * p.K.M + mode + table_offset(p.K, b);
          if (!m.keep && !m.is_inj) etdrk_update_masked(p.K, p.stage, ci, off, p.sb);  // N(u) == 0 there
          else etdrk_update(p.K, p.stage, ci, off, n[ch], p.sb);
        }
      }
    }
  }
}

template <class T> struct RowParams {
  NlParams<T> P;
  FftDesc fd;
  const cpx<T>* tw;
  int mode;
  int nin, nout;             // fields per batch element read / written
  long long rows;            // rows per field = N^(D-1)
  int prune;                 // fast kernels only: PRUNE_IN_ROWS / PRUNE_OUT_ROWS along the last axis
  int in_pitch, out_pitch;   // fast kernels only: elements between consecutive half-complex rows (0 = dense N/2+1)
  long long batch;
  long long in_batch_stride;   // elements (of the in type) between batch elements
  long long out_batch_stride;  // elements (of the out type) between batch elements
  const void* in;
  void* out;
};

// one CTA = one pair of rows (r1, r1+1) of one batch element, all fields
template <class T> __global__ void row_pass_kernel(const RowParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<T>* sm = reinterpret_cast<cpx<T>*>(smem_raw);
  const NlParams<T>& P = p.P;
  const int N = p.fd.N, Nh = N / 2 + 1;
  const long long npairs = (p.rows + 1) / 2;
  const long long b = blockIdx.x / npairs;
  const long long rp = blockIdx.x - b * npairs;
  const long long r1 = 2 * rp, r2 = r1 + 1;
  const bool has2 = r2 < p.rows;
  const int nslots = p.nin > p.nout ? p.nin : p.nout;
  cpx<T>* A = sm;
  cpx<T>* B = sm + (size_t)nslots * N;
  const cpx<T> zero((T)0, (T)0);

  if (p.mode == ROW_R2C) {
    const T* in = (const T*)p.in + (size_t)b * p.in_batch_stride;
    for (int q = threadIdx.x; q < p.nin * N; q += blockDim.x) {
      int f = q / N, x = q - f * N;
      T a = in[((size_t)f * p.rows + r1) * N + x];
      T c = has2 ? in[((size_t)f * p.rows + r2) * N + x] : (T)0;
      A[q] = cpx<T>(a, c);
    }
    __syncthreads();
    cpx<T>* R = fft_lines<T, -1>(A, B, p.nin, N, 1, false, p.fd, p.tw);
    cpx<T>* out = (cpx<T>*)p.out + (size_t)b * p.out_batch_stride;
    for (int q = threadIdx.x; q < p.nin * Nh; q += blockDim.x) {
      int f = q / Nh, k = q - f * Nh;
      cpx<T> X1, X2;
      unpack2(R + (size_t)f * N, N, k, X1, X2);
      out[((size_t)f * p.rows + r1) * Nh + k] = X1;
      if (has2) out[((size_t)f * p.rows + r2) * Nh + k] = X2;
    }
    return;
  }

  // ROW_NL / ROW_C2R start from half-complex rows
  {
    const cpx<T>* in = (const cpx<T>*)p.in + (size_t)b * p.in_batch_stride;
    for (int q = threadIdx.x; q < p.nin * Nh; q += blockDim.x) {
      int f = q / Nh, k = q - f * Nh;
      cpx<T> F1 = in[((size_t)f * p.rows + r1) * Nh + k];
      cpx<T> F2 = has2 ? in[((size_t)f * p.rows + r2) * Nh + k] : zero;
      pack2(A + (size_t)f * N, N, k, F1, F2);
    }
  }
  __syncthreads();
  cpx<T>* R = fft_lines<T, +1>(A, B, p.nin, N, 1, false, p.fd, p.tw);
  cpx<T>* O = (R == A) ? B : A;

  if (p.mode == ROW_C2R) {
    T* out = (T*)p.out + (size_t)b * p.out_batch_stride;
    for (int q = threadIdx.x; q < p.nin * N; q += blockDim.x) {
      int f = q / N, x = q - f * N;
      cpx<T> v = R[q];
      out[((size_t)f * p.rows + r1) * N + x] = v.x * P.inv_norm;
      if (has2) out[((size_t)f * p.rows + r2) * N + x] = v.y * P.inv_norm;
    }
    return;
  }

  for (int x = threadIdx.x; x < N; x += blockDim.x) {
    T i1[EXB_MAX_INV], i2[EXB_MAX_INV], o1[EXB_MAX_FWD], o2[EXB_MAX_FWD];
    for (int f = 0; f < p.nin; ++f) {
      cpx<T> v = R[(size_t)f * N + x];
      i1[f] = v.x * P.inv_norm;
      i2[f] = v.y * P.inv_norm;
    }
    nl_pointwise(P, i1, o1);
    nl_pointwise(P, i2, o2);
    for (int g = 0; g < p.nout; ++g) R[(size_t)g * N + x] = cpx<T>(o1[g], o2[g]);
  }
  __syncthreads();
  cpx<T>* W = fft_lines<T, -1>(R, O, p.nout, N, 1, false, p.fd, p.tw);
  cpx<T>* out = (cpx<T>*)p.out + (size_t)b * p.out_batch_stride;
  for (int q = threadIdx.x; q < p.nout * Nh; q += blockDim.x) {
    int g = q / Nh, k = q - g * Nh;
    cpx<T> X1, X2;
    unpack2(W + (size_t)g * N, N, k, X1, X2);
    out[((size_t)g * p.rows + r1) * Nh + k] = X1;
    if (has2) out[((size_t)g * p.rows + r2) * Nh + k] = X2;
  }
}

// Last ETDRK stage, modes OUTSIDE the dealiasing mask: u+ = exp(dt L) u (etdrk_update_masked) as a streaming pass.
// With the 2/3 rule these are 56 % (2-D) / 70 % (3-D) of all modes; inside the tiled epilogue pass they were 128-byte
// pieces at a line stride with the loads serialised behind 64 registers (c4 stage 1: 3.3 ms vs 1.65 ms for stage 0,
// long_scoreboard 27 per issue -- ncu r02g).  Here a warp owns one last-axis row: the masked part of the row is
// contiguous (the whole row when a leading wavenumber is masked, else the entries k > kmax), exp(dt L) is loaded once
// and reused for every trajectory and channel, all loads of a thread are independent.
template <class T>
__global__ void __launch_bounds__(256) etdrk_masked_linear_kernel(const EtdrkCoefs<T> K, const cpx<T>* U,   // (OUT may alias U)
                                                                 cpx<T>* OUT, int C, int D, int N, int Nh,
                                                                 int kmax, int n1, int i1_off, int i1_mul, long long rows,
                                                                 long long batch, int bchunk) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  int i0 = (int)r, i1 = 0;
  if (D == 3) {
    i0 = (int)(r / n1);
    i1 = (int)(r - (long long)i0 * n1) * i1_mul + i1_off;
  }
  int k0 = wavenumber_of(i0, N), k1 = D == 3 ? wavenumber_of(i1, N) : 0;
  k0 = k0 < 0 ? -k0 : k0;
  k1 = k1 < 0 ? -k1 : k1;
  const int start = (k0 > kmax || k1 > kmax) ? 0 : kmax + 1;
  const long long b0 = (long long)blockIdx.y * bchunk;
  const long long b1 = b0 + bchunk < batch ? b0 + bchunk : batch;
  for (int e = start + lane; e < Nh; e += 32) {
    const long long mode = r * Nh + e;
    for (int c = 0; c < C; ++c) {
      const long long ci = (long long)(K.E == 1 ? 0 : c) * K.M + mode;
      if (K.tstride) {  // stepper ensemble: one table set per group of trajectories
        for (long long b = b0; b < b1; ++b)
          OUT[((size_t)b * C + c) * K.M + mode] = K.exp_term[ci + table_offset(K, b)] * U[((size_t)b * C + c) * K.M + mode];
        continue;
      }
      const cpx<T> Ev = K.exp_term[ci];
      long long b = b0;
      for (; b + 4 <= b1; b += 4) {
        cpx<T> u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = U[((size_t)(b + q) * C + c) * K.M + mode];
#pragma unroll
        for (int q = 0; q < 4; ++q) OUT[((size_t)(b + q) * C + c) * K.M + mode] = Ev * u[q];
      }
      for (; b < b1; ++b) OUT[((size_t)b * C + c) * K.M + mode] = Ev * U[((size_t)b * C + c) * K.M + mode];
    }
  }
}

// order-0 step / generic per-mode scale: out = exp_term * in
template <class T>
__global__ void etdrk0_kernel(const EtdrkCoefs<T> K, int C, long long total, const cpx<T>* in, cpx<T>* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    long long mode = i % K.M;
    int ch = (int)((i / K.M) % C);
    long long ci = (long long)(K.E == 1 ? 0 : ch) * K.M + mode + table_offset(K, i / (K.M * C));
    out[i] = K.exp_term[ci] * in[i];
  }
}

// order-0 step with a per-mode C x C matrix (Wave: exponax/stepper/_wave.py:175-197 composed into one 2 x 2 map):
// out_i = sum_j A[i][j](mode) * in_j.  One thread per (trajectory, mode); `in` may alias `out`.
template <class T>
__global__ void etdrk0_matrix_kernel(const EtdrkCoefs<T> K, int C, long long total, const cpx<T>* in, cpx<T>* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const long long b = i / K.M, mode = i - b * K.M;
    cpx<T> u[EXB_MAXC], r[EXB_MAXC];
#pragma unroll
    for (int c = 0; c < EXB_MAXC; ++c)
      if (c < C) u[c] = in[((size_t)b * C + c) * K.M + mode];
#pragma unroll
    for (int ci = 0; ci < EXB_MAXC; ++ci) {
      if (ci < C) {
        cpx<T> acc((T)0, (T)0);
#pragma unroll
        for (int cj = 0; cj < EXB_MAXC; ++cj)
          if (cj < C) acc = acc + K.exp_term[(long long)(ci * C + cj) * K.M + mode + table_offset(K, b)] * u[cj];
        r[ci] = acc;
      }
    }
#pragma unroll
    for (int c = 0; c < EXB_MAXC; ++c)
      if (c < C) out[((size_t)b * C + c) * K.M + mode] = r[c];
  }
}

// ForcedStepper (exponax/_forced_stepper.py:61-62) in Fourier space: u_hat[b] += scale * f_hat[b * fbatch + ...]
template <class T>
__global__ void add_forcing_kernel(cpx<T>* u, const cpx<T>* f, long long per_traj, long long fbatch, long long batch, T scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < per_traj * batch; i += stride) {
    const long long b = i / per_traj, r = i - b * per_traj;
    u[i] = axpy(scale, f[b * fbatch + r], u[i]);
  }
}

// strided copy of `batch` contiguous chunks of `n` elements
template <class T>
__global__ void copy_batched_kernel(const T* in, long long in_stride, T* out, long long out_stride, long long n,
                                    long long batch) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n * batch; i += stride) {
    long long b = i / n, j = i - b * n;
    out[b * out_stride + j] = in[b * in_stride + j];
  }
}

}  // namespace exb
