// libexb.so -- C ABI (include/exb.h) over the sm_100a kernels.  Host side only: plan creation
// (factorisation, twiddles, coefficient upload) and launch sequencing.  No device allocation,
// synchronisation or foreign stream use outside exb_plan_create / exb_plan_destroy.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/exb.h"
#include "exb_kernels_1d.cuh"
#include "exb_fast1d.h"
#include "exb_fastnd.h"
#include "exb_kernels_nd.cuh"
#include "exb_spectrum.cuh"
#include "exb_metrics.cuh"
#include "exb_ic.cuh"

using namespace exb;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_OK(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess) return fail(EXB_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

struct exb_plan {
  virtual int fused_ok() const { return 1; }
  exb_desc d;
  int64_t launches = 0;
  virtual ~exb_plan() {}
  virtual size_t workspace_bytes(int64_t batch) const = 0;
  virtual int fft(cudaStream_t st, int64_t batch, int channels, const void* u, void* uh, void* ws) = 0;
  virtual int ifft(cudaStream_t st, int64_t batch, int channels, const void* uh, void* u, void* ws) = 0;
  virtual int nonlinear(cudaStream_t st, int64_t batch, const void* uh, void* out, void* ws) = 0;
  virtual int step_fourier(cudaStream_t st, int64_t batch, const void* in, void* out, void* ws) = 0;
  virtual int step(cudaStream_t st, int64_t batch, const void* in, void* out, void* ws) = 0;
  virtual int rollout(cudaStream_t st, int64_t batch, int64_t n_saved, int substeps, unsigned flags,
                      const void* u0, void* out, void* ws) = 0;
  virtual int rollout_forced(cudaStream_t st, int64_t batch, int64_t n_saved, int substeps, unsigned flags,
                             const void* u0, void* out, void* ws, const void* fhat, int64_t fstep, int64_t fbatch,
                             double scale) = 0;
  virtual int slab_pass(cudaStream_t st, int pass, int nfields, int stage, const void* in, void* out,
                        const void* U, void* OUT, void* const* S) = 0;
  virtual void nl_fields(int* ni, int* nf) const = 0;
  virtual int field_pitch() const = 0;
  virtual int slab_inv_pro_fields(cudaStream_t st, int f0, int nf, const void* in, void* out) = 0;
  virtual int slab_pass_peer(cudaStream_t st, int pass, int f0, int nf, const void* in, void* const* peers) = 0;
  virtual int fourier_sums(cudaStream_t st, int64_t nfields, const void* xh, double p, int low, int high,
                           double order, double domain_extent, double* out) = 0;
  virtual int derivative(cudaStream_t st, int64_t nfields, const void* uh, void* out, int order,
                         double domain_extent) = 0;
  virtual int leray(cudaStream_t st, int64_t nfields, const void* uh, void* out, double domain_extent) = 0;
  virtual int spectrum(cudaStream_t st, int64_t nfields, const void* uh, void* out, int power, int average,
                       void* counts) = 0;
  virtual int ic_shape(cudaStream_t st, int64_t nfields, void* uh, int kind, double param, double domain_extent,
                       double dc_value) = 0;
};

static void factorize(int N, FftDesc& fd) {
  fd.N = N;
  fd.nst = 0;
  int n = N, e = 0;
  while (n % 2 == 0) {
    n /= 2;
    ++e;
  }
  int n8 = e / 3, rem = e % 3;
  if (rem == 1 && n8 >= 1) {  // 8^(n8-1) * 4 * 4
    for (int i = 0; i < n8 - 1; ++i) fd.radix[fd.nst++] = 8;
    fd.radix[fd.nst++] = 4;
    fd.radix[fd.nst++] = 4;
  } else {
    for (int i = 0; i < n8; ++i) fd.radix[fd.nst++] = 8;
    if (rem == 1) fd.radix[fd.nst++] = 2;
    if (rem == 2) fd.radix[fd.nst++] = 4;
  }
  while (n % 5 == 0) {
    n /= 5;
    fd.radix[fd.nst++] = 5;
  }
  while (n % 3 == 0) {
    n /= 3;
    fd.radix[fd.nst++] = 3;
  }
  for (int p = 7; (long long)p * p <= n; p += 2)
    while (n % p == 0) {
      n /= p;
      fd.radix[fd.nst++] = p;
    }
  if (n > 1) fd.radix[fd.nst++] = n;
  int Ns = 1;
  for (int s = 0; s < fd.nst; ++s) {
    const int R = fd.radix[s];
    fd.tunit[s] = N / (Ns * R);
    fd.m_ns[s] = fastdiv_magic((unsigned)Ns);
    fd.m_nr[s] = fastdiv_magic((unsigned)(N / R));
    Ns *= R;
  }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <class T> struct PlanImpl : exb_plan {
  NlParams<T> P;
  EtdrkCoefs<T> K;
  FftDesc fd;
  cpx<T>* d_tw = nullptr;
  std::vector<void*> d_allocs;
  int D, N, Nh, C;
  long long M;     // modes per channel
  long long G;     // grid points per channel
  int Nhp = 0;     // last-axis pitch of the internal FIELD buffers (n_inv / n_fwd fields between the passes of one
  long long Mf = 0;  // N(u) evaluation) and their field stride: padded to a multiple of 16 complex elements (128 B) on the
                     // fast N-D path so that every tile row is line-aligned and a tensor map can describe the buffers;
                     // == Nh / M otherwise (generic kernels, slab plans whose buffers the caller lays out)
  int nscr;        // ETDRK scratch states
  int nranks = 1;  // slab decomposition (3-D): number of ranks, local extent of the split axis
  int nloc = 0;
  int max_smem = 0;
  int sm_count = 148;
  const void* twiddle_host_tmp = nullptr;

  ~PlanImpl() override {
    for (void* p : d_allocs) cudaFree(p);
  }

  bool tables_on_device = false;
  int upload(const void* host, size_t bytes, void** dev) {
    if (tables_on_device && host != (const void*)twiddle_host_tmp) {  // caller-owned device table
      *dev = const_cast<void*>(host);
      return EXB_OK;
    }
    CUDA_OK(cudaMalloc(dev, bytes));
    d_allocs.push_back(*dev);
    CUDA_OK(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
    return EXB_OK;
  }

  int init(const exb_desc& desc) {
    d = desc;
    tables_on_device = desc.tables_on_device != 0;
    D = desc.num_spatial_dims;
    N = desc.num_points;
    C = desc.num_channels;
    Nh = N / 2 + 1;
    if (D < 1 || D > 3) return fail(EXB_EINVAL, "num_spatial_dims must be 1, 2 or 3 (got %d)", D);
    if (N < 2) return fail(EXB_EINVAL, "num_points must be >= 2 (got %d)", N);
    if (C < 1 || C > EXB_MAXC) return fail(EXB_EUNSUPPORTED, "num_channels must be in 1..%d (got %d)", EXB_MAXC, C);
    if (desc.lin_matrix) {
      if (desc.lin_channels != C * C) return fail(EXB_EINVAL, "lin_matrix needs lin_channels = num_channels^2");
      if (desc.order != 0 || desc.nl_kind != EXB_NL_ZERO)
        return fail(EXB_EINVAL, "lin_matrix is an order-0 feature (no nonlinear function)");
      if (desc.slab_nranks > 1) return fail(EXB_EUNSUPPORTED, "lin_matrix is not available for slab plans");
    } else if (desc.lin_channels != 1 && desc.lin_channels != C)
      return fail(EXB_EINVAL, "lin_channels must be 1 or num_channels");
    if (desc.order < 0 || desc.order > 4) return fail(EXB_EINVAL, "order %d not implemented", desc.order);
    M = Nh;
    G = N;
    for (int i = 1; i < D; ++i) {
      M *= N;
      G *= N;
    }
    nranks = desc.slab_nranks > 1 ? desc.slab_nranks : 1;
    nloc = N;
    if (nranks > 1) {
      if (D != 3) return fail(EXB_EINVAL, "slab decomposition needs num_spatial_dims = 3");
      if (N % nranks != 0) return fail(EXB_EINVAL, "num_points must be divisible by slab_nranks");
      if (desc.slab_rank < 0 || desc.slab_rank >= nranks) return fail(EXB_EINVAL, "slab_rank out of range");
      nloc = N / nranks;
      M = (long long)N * nloc * Nh;  // local spectral slab (N, N/P, Nh) == local half-complex physical slab
      G = (long long)nloc * N * N;   // local physical slab (N/P, N, N)
    }
    factorize(N, fd);
    if (fd.nst > EXB_MAX_STAGES) return fail(EXB_EUNSUPPORTED, "too many FFT stages");

    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));

    // ---- nonlinear function descriptor ----
    memset(&P, 0, sizeof(P));
    P.kind = desc.nl_kind;
    P.D = D;
    P.N = N;
    P.Nh = Nh;
    P.C = C;
    P.kmax = desc.dealias_kmax;
    P.single_channel = desc.single_channel;
    P.conservative = desc.conservative;
    P.zero_mode_fix = desc.zero_mode_fix;
    P.n_poly = desc.n_poly;
    P.has_inj = desc.has_injection;
    for (int i = 0; i < 3; ++i) P.inj_idx[i] = desc.injection_index[i];
    P.inj_val = (T)desc.injection_value;
    P.dscale = (T)(2.0 * M_PI / desc.domain_extent);
    P.scale = (T)desc.nl_scale;
    slab_cyclic = nranks > 1 && desc.slab_cyclic;
    P.i1_mul = slab_cyclic ? nranks : 1;
    P.i1_off = nranks > 1 ? (slab_cyclic ? desc.slab_rank : desc.slab_rank * (N / nranks)) : 0;
    for (int i = 0; i < EXB_MAX_POLY; ++i) P.poly[i] = (T)desc.poly[i];
    for (int i = 0; i < 3; ++i) P.gen[i] = (T)desc.general_scales[i];
    {
      double g = 1.0;
      for (int i = 0; i < D; ++i) g *= N;
      P.inv_norm = (T)(1.0 / g);
    }
    int order = desc.order;
    switch (P.kind) {
      case EXB_NL_ZERO:
        P.n_inv = P.n_fwd = 0;
        order = 0;  // every ETDRK order degenerates to exp(dt L) u when N == 0
        break;
      case EXB_NL_CONVECTION:
        if (P.single_channel) {
          if (P.conservative) {
            P.n_inv = C;
            P.n_fwd = C;
          } else {
            if (C != 1) return fail(EXB_EINVAL, "single-channel non-conservative convection needs 1 channel");
            P.n_inv = 1 + D;
            P.n_fwd = 1;
          }
        } else {
          if (C != D)
            return fail(EXB_EINVAL, "Number of channels in u_hat should match number of spatial dimensions");
          if (P.conservative) {
            P.n_inv = C;
            P.n_fwd = C * C;
          } else {
            P.n_inv = C + C * D;
            P.n_fwd = C;
          }
        }
        break;
      case EXB_NL_GRADIENT_NORM:
        P.n_inv = C * D;
        P.n_fwd = C;
        break;
      case EXB_NL_POLYNOMIAL:
        if (P.n_poly < 0 || P.n_poly > EXB_MAX_POLY) return fail(EXB_EUNSUPPORTED, "at most %d polynomial coefficients", EXB_MAX_POLY);
        P.n_inv = C;
        P.n_fwd = C;
        break;
      case EXB_NL_VORTICITY_2D:
        if (D != 2) return fail(EXB_EINVAL, "Expected num_spatial_dims = 2, got %d.", D);
        if (C != 1) return fail(EXB_EINVAL, "vorticity convection needs 1 channel");
        P.n_inv = 4;
        P.n_fwd = 1;
        break;
      case EXB_NL_PROJECTED_3D:
        if (D != 3) return fail(EXB_EINVAL, "ProjectedConvection3d only supports 3 spatial dimensions.");
        if (C != 3) return fail(EXB_EINVAL, "projected convection needs 3 channels");
        P.n_inv = 6;
        P.n_fwd = 3;
        break;
      case EXB_NL_GENERAL:
        P.n_inv = C * (1 + D);
        P.n_fwd = 2 * C;
        break;
      case EXB_NL_GRAY_SCOTT:
        if (C != 2) return fail(EXB_EINVAL, "num_channels must be 2");
        P.n_inv = 2;
        P.n_fwd = 2;
        break;
      case EXB_NL_CAHN_HILLIARD:
        if (C != 1) return fail(EXB_EINVAL, "Cahn-Hilliard needs 1 channel");
        P.n_inv = 1;
        P.n_fwd = 1;
        break;
      default:
        return fail(EXB_EINVAL, "unknown nl_kind %d", P.kind);
    }
    if (P.n_inv > EXB_MAX_INV || P.n_fwd > EXB_MAX_FWD) return fail(EXB_EUNSUPPORTED, "too many fields");

    // ---- twiddles: N-th roots of unity in double, rounded once ----
    {
      std::vector<cpx<T>> tw(N);
      for (int j = 0; j < N; ++j) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)N;
        tw[j] = cpx<T>((T)cosl(a), (T)sinl(a));
      }
      if constexpr (std::is_same<T, float>::value) {
        // register-FFT kernels: their per-pass table, already arranged, right behind the roots (Fft8Tw<N>::fill)
        const int extra = D >= 2 ? exb_fastnd_tw_size(N) : 0;
        if (extra > 0) {
          tw.resize(N + extra);
          exb_fastnd_tw_arrange(N, tw.data(), tw.data() + N);
        }
      }
      twiddle_host_tmp = tw.data();
      int rc = upload(tw.data(), sizeof(cpx<T>) * tw.size(), (void**)&d_tw);
      twiddle_host_tmp = nullptr;
      if (rc) return rc;
    }
    // ---- ETDRK coefficients ----
    memset(&K, 0, sizeof(K));
    K.order = order;
    K.E = desc.lin_channels;
    K.lin_matrix = desc.lin_matrix ? 1 : 0;
    K.M = M;
    table_sets = desc.table_sets > 1 ? desc.table_sets : 1;
    if (table_sets > 1 && nranks > 1) return fail(EXB_EUNSUPPORTED, "stepper ensembles are not available for slab plans");
    K.tstride = table_sets > 1 ? (long long)K.E * M : 0;
    K.trep = 1;
    const size_t ne = (size_t)K.E * M * table_sets;
    if (!desc.exp_term) return fail(EXB_EINVAL, "exp_term is required");
    int rc = upload(desc.exp_term, ne * sizeof(cpx<T>), (void**)&K.exp_term);
    if (rc) return rc;
    if (order >= 3) {
      if (!desc.half_exp_term) return fail(EXB_EINVAL, "half_exp_term is required for order >= 3");
      rc = upload(desc.half_exp_term, ne * sizeof(cpx<T>), (void**)&K.half_exp);
      if (rc) return rc;
    }
    static const int ncoef[5] = {0, 1, 2, 5, 6};
    for (int i = 0; i < ncoef[order]; ++i) {
      if (!desc.coef[i]) return fail(EXB_EINVAL, "coef[%d] is required for order %d", i, order);
      rc = upload(desc.coef[i], ne * sizeof(T), (void**)&K.c[i]);
      if (rc) return rc;
    }
    nscr = etdrk_num_scratch(order);
    // Trajectory-major chunking (EXB_ND_CHUNK) is OFF by default: measured on B200 (c3, 512^2, B=512) the
    // pass kernels are issue-bound, not DRAM-bound, so keeping a chunk's intermediates in L2 does not pay and the
    // smaller grids lose occupancy (chunk 11: 0.31, chunk 48: 0.38, whole batch: 0.43 of the HBM roofline).
    nd_chunk = 0;
    if constexpr (std::is_same<T, float>::value) {
      fast_nd = D >= 2 && !getenv("EXB_DISABLE_FAST_ND") && exb_fastnd_supported(D, N, P);
    }
    Nhp = Nh;
    Mf = M;
    if (fast_nd && nranks == 1 && !getenv("EXB_DENSE_FIELDS")) {
      Nhp = (Nh + 15) / 16 * 16;
      Mf = M / Nh * Nhp;
    }
    // Slab plans: the field buffers ARE the all-to-all payload.  Their last axis only ever holds the wavenumbers
    // inside the dealiasing mask (pruned passes never touch k > kmax), so they are laid out with the COMPACT pitch
    // kmax + 1 (rounded up to 16): with the 2/3 rule every transpose ships 2/3 of the bytes ("prune to
    // the mask").  The caller allocates them with exb_plan_field_pitch().
    if (fast_nd && nranks > 1 && P.kmax >= 0 && !getenv("EXB_DENSE_FIELDS")) {
      const int kp = (P.kmax + 1 + 15) / 16 * 16;
      if (kp < Nh) {
        Nhp = kp;
        Mf = M / Nh * Nhp;
      }
    }

    if (D == 1) {
      // The persistent kernel keeps the state, the ETDRK stage buffers and the transform lines of a trajectory
      // pair in shared memory.  Grids too large for that (f32 ETDRK2 Burgers: N > ~4000) keep their standalone
      // transforms (exb_fft / exb_ifft need only 2 N complex values); the fused entry points then answer
      // EXB_EUNSUPPORTED and the host runs the stage formulas around those transforms (exb_plan_fused_ok).
      size_t need = smem_1d(C, nslots_1d());
      fused_1d_ok = (long long)need <= max_smem;
      fused_1d_need = need;
      if ((size_t)2 * N * sizeof(cpx<T>) > (size_t)max_smem)
        return fail(EXB_EUNSUPPORTED, "1-D transforms need %zu B of shared memory for N=%d (limit %d)",
                    (size_t)2 * N * sizeof(cpx<T>), N, max_smem);
      CUDA_OK(cudaFuncSetAttribute(k1d_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    } else {
      CUDA_OK(cudaFuncSetAttribute(col_pass_kernel<T, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      CUDA_OK(cudaFuncSetAttribute(col_pass_kernel<T, +1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      CUDA_OK(cudaFuncSetAttribute(row_pass_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      // every pass must fit at tile width >= 1
      int maxf = P.n_fwd + 1 > C + 2 ? P.n_fwd + 1 : C + 2;
      if ((size_t)maxf * N * sizeof(cpx<T>) > (size_t)max_smem)
        return fail(EXB_EUNSUPPORTED, "N=%d too large for the shared-memory column pass", N);
      int rs = P.n_inv > P.n_fwd ? P.n_inv : P.n_fwd;
      if (rs < C) rs = C;
      if ((size_t)2 * rs * N * sizeof(cpx<T>) > (size_t)max_smem)
        return fail(EXB_EUNSUPPORTED, "N=%d too large for the shared-memory row pass", N);
    }
    return EXB_OK;
  }

  // ------------------------------------------------------------------ 1-D
  int nslots_1d() const {
    int s = P.n_inv > P.n_fwd ? P.n_inv : P.n_fwd;
    if (s < C) s = C;
    return s < 1 ? 1 : s;
  }
  size_t smem_1d(int ch, int nslots) const {
    return ((size_t)2 * nslots * N + (size_t)(1 + nscr) * 2 * ch * Nh) * sizeof(cpx<T>);
  }
  int threads_1d(int lines) const {
    int work = (N / 4 > 1 ? N / 4 : 1) * (lines < 1 ? 1 : (lines > 2 ? 2 : lines));
    int t = (work + 31) / 32 * 32;
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    return t;
  }

  // ---- fast path: N = R*R, f32, one channel (exb_kernels_1d_fast.cuh) ----
  bool slab_cyclic = false;   // slab plans: cyclic distribution of the axis-1 indices over the ranks
  int table_sets = 1;   // stepper ensemble: number of coefficient table sets (exb_desc.table_sets)
  // trajectories per table set for a call over `batch` trajectories (EXB_EINVAL if the batch does not divide)
  int set_table_rep(int64_t batch) {
    if (table_sets <= 1) return EXB_OK;
    if (batch % table_sets) return fail(EXB_EINVAL, "batch %lld is not a multiple of the %d table sets of this ensemble plan",
                                        (long long)batch, table_sets);
    K.trep = batch / table_sets;
    return EXB_OK;
  }
  bool fast_1d_ok() const {
    if constexpr (std::is_same<T, float>::value) {
      if (D != 1 || C != 1 || K.E != 1 || table_sets > 1) return false;   // (the fast kernel shares its tables per CTA)
      if (getenv("EXB_DISABLE_FAST_1D")) return false;
      return exb_fast1d_supported(N, P, K.order);
    }
    return false;
  }
  int launch_fast(cudaStream_t st, const K1dParams<T>& pt) {
    if constexpr (std::is_same<T, float>::value) {
      const char* err = nullptr;
      int rc = N == 256 ? exb_launch_fast1d_r16(st, pt, nscr, max_smem, &err)
                        : exb_launch_fast1d_r8(st, pt, nscr, max_smem, &err);
      if (rc) return fail(rc, "%s", err ? err : "fast 1-D launch failed");
      ++launches;
      return EXB_OK;
    }
    return fail(EXB_EUNSUPPORTED, "fast 1-D path is f32 only");
  }

  int launch_1d(cudaStream_t st, int op, long long batch, int ch, const void* in, void* out,
                long long n_saved, int substeps, unsigned flags) {
    if (batch <= 0) return EXB_OK;
    K1dParams<T> p;
    p.P = P;
    p.K = K;
    p.fd = fd;
    p.tw = d_tw;
    p.batch = batch;
    p.op = op;
    p.C = ch;
    p.in = in;
    p.out = out;
    p.n_saved = n_saved;
    p.substeps = substeps;
    p.flags = flags;
    p.forcing = op == OP1_ROLLOUT ? frc.f : nullptr;
    p.fstep = frc.step;
    p.fbatch = frc.batch;
    p.fscale = frc.scale;
    if (op == OP1_ROLLOUT && fast_1d_ok()) return launch_fast(st, p);
    bool plain = (op == OP1_FFT || op == OP1_IFFT);
    if (!plain && !fused_1d_ok)
      return fail(EXB_EUNSUPPORTED,
                  "the fused 1-D kernel needs %zu B of shared memory for N=%d (limit %d): use the standalone "
                  "transforms (exb_fft / exb_ifft) with the stage formulas on the host side",
                  fused_1d_need, N, max_smem);
    p.nslots = plain ? ch : nslots_1d();
    size_t smem = plain ? (size_t)2 * ch * N * sizeof(cpx<T>) : smem_1d(ch, p.nslots);
    long long grid = (batch + 1) / 2;
    int nthr = threads_1d(plain ? ch : p.nslots);
    k1d_kernel<T><<<(unsigned)grid, nthr, smem, st>>>(p);
    ++launches;
    CUDA_OK(cudaGetLastError());
    return EXB_OK;
  }

  // ------------------------------------------------------------------ N-D
  bool fast_nd = false;  // set in init(): register-FFT kernels available for this (D, N, N(u))
  struct Forcing {            // set for the duration of one exb_rollout_forced call
    const cpx<T>* f = nullptr;
    long long step = 0, batch = 0;
    T scale = (T)0;
  } frc;
  bool fused_1d_ok = true;   // 1-D: state + stage buffers + transform lines fit shared memory
  size_t fused_1d_need = 0;
  int fused_ok() const override { return D != 1 || fused_1d_ok; }
  int fast_tw() const { return N == 512 ? 8 : (N == 1024 ? 4 : (N == 2048 ? 2 : 16)); }  // == TW of exb_fastnd_n*.cu
  int launch_col_fast(cudaStream_t st, ColParams<T>& p, int dir, long long units) {
    if constexpr (std::is_same<T, float>::value) {
      p.TW = fast_tw();
      long long ntiles = (p.inner + p.TW - 1) / p.TW;
      const char* err = nullptr;
      int rc = exb_fastnd_col(st, p, dir, ntiles * units, &err);
      if (rc) return fail(rc, "%s", err ? err : "fast column pass failed");
      ++launches;
      return EXB_OK;
    }
    return fail(EXB_EUNSUPPORTED, "fast N-D path is f32 only");
  }
  int launch_row_fast(cudaStream_t st, RowParams<T>& p) {
    if constexpr (std::is_same<T, float>::value) {
      const char* err = nullptr;
      int rc = exb_fastnd_row(st, p, &err);
      if (rc) return fail(rc, "%s", err ? err : "fast row pass failed");
      ++launches;
      return EXB_OK;
    }
    return fail(EXB_EUNSUPPORTED, "fast N-D path is f32 only");
  }

  int pick_tw(int ntiles_resident) const {
    // widest tile (<= 16 lines) such that `ntiles_resident` tiles fit in shared memory
    int tw = 16;
    while (tw > 1 && (size_t)ntiles_resident * N * tw * sizeof(cpx<T>) > (size_t)max_smem - 1024) tw /= 2;
    return tw;
  }

  // axis: 0 or 1 (3-D only for axis 1).  Geometry of a strided pass over that axis.
  void col_geom(ColParams<T>& p, int axis) const {
    if (D == 2 || axis == 0) {
      p.line_stride = M / N;  // Nh (2-D) or N*Nh (3-D)
      p.inner = M / N;
      p.n_outer = 1;
      p.outer_stride = 0;
    } else {  // 3-D axis 1 (slab mode: the local physical slab has N/P planes)
      p.line_stride = Nh;
      p.inner = Nh;
      p.n_outer = nranks > 1 ? nloc : N;
      p.outer_stride = (long long)N * Nh;
    }
  }

  // fields: in / out are internal field buffers (pitch Nhp, field stride Mf); otherwise dense spectral arrays
  template <int DIR>
  int col_plain(cudaStream_t st, int axis, long long batch, int nfields, const cpx<T>* in, cpx<T>* out,
                int prune = 0, bool segmented = false, void* const* peers = nullptr, long long peer_field_off = 0,
                bool fields = false) {
    ColParams<T> p;
    memset(&p, 0, sizeof(p));
    p.prune = prune;
    p.P = P;
    p.K = K;
    p.fd = fd;
    p.tw = d_tw;
    p.mode = COL_PLAIN;
    p.TW = pick_tw(2);
    p.nfields = nfields;
    p.M = M;
    p.batch = batch;
    p.in = in;
    p.out = out;
    col_geom(p, axis);
    if (fields && Nhp != Nh) {  // (3-D axis 1 only: the lines of a padded field buffer)
      p.M = Mf;
      p.line_stride = Nhp;
      p.outer_stride = (long long)N * Nhp;
    }
    if (segmented) {  // slab layout A as received from / sent to the peers: [peer][x][n][pitch]
      const long long pitch = fields ? Nhp : Nh;
      p.seg_len = nloc;
      p.seg_cyclic = slab_cyclic ? nranks : 0;
      p.seg_stride = (long long)nloc * nloc * pitch;
      p.outer_stride = (long long)nloc * pitch;
    }
    if (peers) {
      if (!fast_nd || !segmented) return fail(EXB_EUNSUPPORTED, "peer stores need the fast N-D kernels");
      p.peer = 1;
      p.peer_off = (long long)this->d.slab_rank * p.seg_stride + peer_field_off;
      for (int r = 0; r < nranks; ++r) p.peer_out[r] = (cpx<T>*)peers[r];
    }
    if (fast_nd) return launch_col_fast(st, p, DIR, p.n_outer * batch * nfields);
    long long ntiles = (p.inner + p.TW - 1) / p.TW;
    long long grid = ntiles * p.n_outer * batch * nfields;
    size_t smem = (size_t)2 * N * p.TW * sizeof(cpx<T>);
    col_pass_kernel<T, DIR><<<(unsigned)grid, 256, smem, st>>>(p);
    ++launches;
    CUDA_OK(cudaGetLastError());
    return EXB_OK;
  }

  int col_inv_pro(cudaStream_t st, long long batch, const cpx<T>* state, cpx<T>* winv, int f0 = 0, int fcount = 0,
                  void* const* peers = nullptr) {
    ColParams<T> p;
    memset(&p, 0, sizeof(p));
    p.f0 = f0;
    p.fcount = fcount;
    p.P = P;
    p.K = K;
    p.fd = fd;
    p.tw = d_tw;
    p.mode = COL_INV_PRO;
    p.prune = PRUNE_COLS;
    p.TW = pick_tw(C + 2);
    p.nfields = P.n_inv;
    p.M = M;
    p.batch = batch;
    p.in = state;
    p.out = winv;
    col_geom(p, 0);
    p.fpitch = Nhp;
    p.fM = Mf;
    if (peers) {
      if (!fast_nd || nranks <= 1) return fail(EXB_EUNSUPPORTED, "peer stores need the fast N-D kernels");
      p.peer = 1;
      p.seg_len = nloc;                                      // x-planes per rank
      p.peer_off = (long long)this->d.slab_rank * nloc * nloc * Nhp;      // block [me] of the peer's [src][x][k1][pitch] buffer
      for (int r = 0; r < nranks; ++r) p.peer_out[r] = (cpx<T>*)peers[r];
    }
    if (fast_nd) return launch_col_fast(st, p, +1, batch);
    long long ntiles = (p.inner + p.TW - 1) / p.TW;
    long long grid = ntiles * batch;
    size_t smem = (size_t)(C + 2) * N * p.TW * sizeof(cpx<T>);
    col_pass_kernel<T, +1><<<(unsigned)grid, 256, smem, st>>>(p);
    ++launches;
    CUDA_OK(cudaGetLastError());
    return EXB_OK;
  }

  // the dealiased modes all have N(u) == 0 unless the (single) injection mode lies outside the mask
  bool masked_stream_ok() const {
    if (P.kmax < 0 || getenv("EXB_NO_MASKED_STREAM")) return false;
    if (P.has_inj) {
      for (int d = 0; d < D; ++d) {
        int k = d == D - 1 ? P.inj_idx[d] : wavenumber_of(P.inj_idx[d], N);
        if ((k < 0 ? -k : k) > P.kmax) return false;
      }
    }
    return true;
  }
  // can the ETDRK2 epilogue of this plan run the next evaluation's prologue pass in the same launch?
  bool can_fuse_prologue() const {
    return fast_nd && D == 2 && C == 1 && K.order == 2 && nranks == 1 && masked_stream_ok() && !getenv("EXB_NO_FUSE");
  }
  int col_fwd(cudaStream_t st, long long batch, int mode, int stage, const cpx<T>* wfwd, cpx<T>* nl_out,
              const StateBufs<T>& sb, cpx<T>* fuse_winv = nullptr) {
    ColParams<T> p;
    memset(&p, 0, sizeof(p));
    p.P = P;
    p.K = K;
    p.fd = fd;
    p.tw = d_tw;
    p.mode = mode;
    p.prune = PRUNE_COLS;
    p.TW = pick_tw(P.n_fwd + 1);
    p.nfields = P.n_fwd;
    p.stage = stage;
    p.M = M;
    p.batch = batch;
    p.in = wfwd;
    p.out = nl_out;
    p.sb = sb;
    col_geom(p, 0);
    p.fpitch = Nhp;
    p.fM = Mf;
    if (fuse_winv && mode == COL_FWD_EPI) {
      p.fuse_next = 1;
      p.out = fuse_winv;
    }
    // last stage: the dealiased modes (u+ = exp(dt L) u) go through a streaming pass instead of the tiled epilogue
    const bool masked_stream = fast_nd && mode == COL_FWD_EPI && stage == K.order - 1 && masked_stream_ok();
    p.masked_external = masked_stream ? 1 : 0;
    if (masked_stream) {
      const long long rows = M / Nh;
      const int n1 = nranks > 1 ? nloc : N;
      long long bchunk = batch;
      while (rows / 8 * ((batch + bchunk - 1) / bchunk) < 4LL * sm_count && bchunk > 4) bchunk = (bchunk + 1) / 2;
      dim3 grid((unsigned)((rows + 7) / 8), (unsigned)((batch + bchunk - 1) / bchunk));
      etdrk_masked_linear_kernel<T><<<grid, 256, 0, st>>>(K, sb.U, sb.OUT, C, D, N, Nh, P.kmax, n1, P.i1_off, P.i1_mul, rows, batch,
                                                         (int)bchunk);
      ++launches;
      CUDA_OK(cudaGetLastError());
    }
    {  // coefficient tables this stage reads (E or E/2 complex + one or more real tables) vs the L2 (126 MB)
      static const char* env = getenv("EXB_EPI_BATCH_FASTEST");
      const size_t table_bytes = (size_t)K.E * M * (sizeof(cpx<T>) + sizeof(T));
      p.batch_fastest = env ? atoi(env) : (batch > 1 && table_bytes > ((size_t)48 << 20));
    }
    if (fast_nd) return launch_col_fast(st, p, -1, batch);
    long long ntiles = (p.inner + p.TW - 1) / p.TW;
    long long grid = ntiles * batch;
    size_t smem = (size_t)(P.n_fwd + 1) * N * p.TW * sizeof(cpx<T>);
    col_pass_kernel<T, -1><<<(unsigned)grid, 256, smem, st>>>(p);
    ++launches;
    CUDA_OK(cudaGetLastError());
    return EXB_OK;
  }

  int row_pass(cudaStream_t st, int mode, long long batch, int nin, int nout, const void* in,
               long long in_bs, void* out, long long out_bs, int in_pitch = 0, int out_pitch = 0) {
    RowParams<T> p;
    memset(&p, 0, sizeof(p));
    p.in_pitch = in_pitch;
    p.out_pitch = out_pitch;
    p.prune = mode == ROW_NL ? (PRUNE_IN_ROWS | PRUNE_OUT_ROWS) : 0;
    p.P = P;
    p.fd = fd;
    p.tw = d_tw;
    p.mode = mode;
    p.nin = nin;
    p.nout = nout;
    p.rows = G / N;
    p.batch = batch;
    p.in_batch_stride = in_bs;
    p.out_batch_stride = out_bs;
    p.in = in;
    p.out = out;
    if (fast_nd) return launch_row_fast(st, p);
    int nslots = nin > nout ? nin : nout;
    size_t smem = (size_t)2 * nslots * N * sizeof(cpx<T>);
    long long grid = ((p.rows + 1) / 2) * batch;
    int nthr = threads_1d(nslots);
    row_pass_kernel<T><<<(unsigned)grid, nthr, smem, st>>>(p);
    ++launches;
    CUDA_OK(cudaGetLastError());
    return EXB_OK;
  }

  // physical (batch stride in_bs reals, `ch` fields) -> spectral dst (dense, ch fields)
  int fft_nd(cudaStream_t st, long long batch, int ch, const T* src, long long in_bs, cpx<T>* dst) {
    int rc = row_pass(st, ROW_R2C, batch, ch, ch, src, in_bs, dst, (long long)ch * M);
    if (rc) return rc;
    if (D == 3) {
      rc = col_plain<-1>(st, 1, batch, ch, dst, dst);
      if (rc) return rc;
    }
    return col_plain<-1>(st, 0, batch, ch, dst, dst);
  }

  // spectral src (dense; untouched) -> physical dst; tmp: dense spectral scratch of the same size
  int ifft_nd(cudaStream_t st, long long batch, int ch, const cpx<T>* src, cpx<T>* tmp, T* dst, long long out_bs) {
    int rc = col_plain<+1>(st, 0, batch, ch, src, tmp);
    if (rc) return rc;
    if (D == 3) {
      rc = col_plain<+1>(st, 1, batch, ch, tmp, tmp);
      if (rc) return rc;
    }
    return row_pass(st, ROW_C2R, batch, ch, ch, tmp, (long long)ch * M, dst, out_bs);
  }

  // workspace carving (N-D)
  struct Ws {
    cpx<T>* S[4];
    cpx<T>* Uh;
    cpx<T>* Winv;
    cpx<T>* Wfwd;
  };
  size_t state_bytes(long long batch) const { return align_up((size_t)batch * C * M * sizeof(cpx<T>), 256); }
  int winv_fields() const { return P.n_inv > C ? P.n_inv : C; }
  int wfwd_fields() const { return P.n_fwd > C ? P.n_fwd : C; }
  size_t workspace_bytes(int64_t batch) const override {
    if (D == 1 || batch <= 0) return 0;
    size_t sb = state_bytes(batch);
    size_t fb = align_up((size_t)batch * Mf * sizeof(cpx<T>), 256);
    return sb * (size_t)(nscr + 1) + fb * (size_t)(winv_fields() + wfwd_fields());
  }
  Ws carve(void* ws, long long batch) const {
    Ws w;
    unsigned char* p = (unsigned char*)ws;
    size_t sb = state_bytes(batch);
    size_t fb = align_up((size_t)batch * Mf * sizeof(cpx<T>), 256);
    for (int i = 0; i < 4; ++i) w.S[i] = nullptr;
    for (int i = 0; i < nscr; ++i) {
      w.S[i] = (cpx<T>*)p;
      p += sb;
    }
    w.Uh = (cpx<T>*)p;
    p += sb;
    w.Winv = (cpx<T>*)p;
    p += fb * winv_fields();
    w.Wfwd = (cpx<T>*)p;
    return w;
  }

  // N(state) -> forward fields in w.Wfwd (all passes except the final forward column pass)
  // have_winv: the inverse fields of `state` were already produced by the previous (fused) epilogue pass
  int nl_front_nd(cudaStream_t st, long long batch, const cpx<T>* state, const Ws& w, bool have_winv = false) {
    int rc = have_winv ? EXB_OK : col_inv_pro(st, batch, state, w.Winv);
    if (rc) return rc;
    if (D == 3) {
      rc = col_plain<+1>(st, 1, batch, P.n_inv, w.Winv, w.Winv, PRUNE_COLS | PRUNE_IN_ROWS, false, nullptr, 0, true);
      if (rc) return rc;
    }
    rc = row_pass(st, ROW_NL, batch, P.n_inv, P.n_fwd, w.Winv, (long long)P.n_inv * Mf, w.Wfwd,
                  (long long)P.n_fwd * Mf, Nhp, Nhp);
    if (rc) return rc;
    if (D == 3) {
      rc = col_plain<-1>(st, 1, batch, P.n_fwd, w.Wfwd, w.Wfwd, PRUNE_COLS | PRUNE_OUT_ROWS, false, nullptr, 0, true);
      if (rc) return rc;
    }
    return EXB_OK;
  }

  // have_winv: the previous step's last epilogue already ran this step's first prologue pass (fused);
  // want_next: another step follows directly on `out` (no physical-space round trip in between) -> fuse its prologue
  int step_fourier_nd(cudaStream_t st, long long batch, const cpx<T>* in, cpx<T>* out, const Ws& w,
                      bool have_winv = false, bool want_next = false) {
    if (K.order == 0 && K.lin_matrix) {
      long long total = batch * M;
      int grid = (int)((total + 255) / 256 < (long long)sm_count * 16 ? (total + 255) / 256 : (long long)sm_count * 16);
      etdrk0_matrix_kernel<T><<<grid, 256, 0, st>>>(K, C, total, in, out);
      ++launches;
      CUDA_OK(cudaGetLastError());
      return EXB_OK;
    }
    if (K.order == 0) {
      long long total = batch * C * M;
      int grid = (int)((total + 255) / 256 < (long long)sm_count * 16 ? (total + 255) / 256 : (long long)sm_count * 16);
      etdrk0_kernel<T><<<grid, 256, 0, st>>>(K, C, total, in, out);
      ++launches;
      CUDA_OK(cudaGetLastError());
      return EXB_OK;
    }
    StateBufs<T> sb;
    sb.U = in;
    sb.OUT = out;
    for (int i = 0; i < 4; ++i) sb.S[i] = w.S[i];
    const bool fuse = can_fuse_prologue();
    bool fused_prev = have_winv && fuse;
    for (int s = 0; s < K.order; ++s) {
      int si = etdrk_stage_input(K.order, s);
      const cpx<T>* src = si < 0 ? in : w.S[si];
      int rc = nl_front_nd(st, batch, src, w, fused_prev);
      if (rc) return rc;
      // (the row pass has consumed Winv by the time the epilogue of the same stage overwrites it: stream order)
      const bool fuse_this = fuse && (s < K.order - 1 || want_next);
      rc = col_fwd(st, batch, COL_FWD_EPI, s, w.Wfwd, nullptr, sb, fuse_this ? w.Winv : nullptr);
      if (rc) return rc;
      fused_prev = fuse_this;
    }
    return EXB_OK;
  }

  // ------------------------------------------------------------------ slab passes
  int slab_inv_pro_fields(cudaStream_t st, int f0, int nf, const void* in, void* out) override {
    if (D != 3) return fail(EXB_EINVAL, "exb_slab_inv_pro_fields needs a 3-D plan");
    if (f0 < 0 || nf < 1 || f0 + nf > P.n_inv) return fail(EXB_EINVAL, "field range out of bounds");
    return col_inv_pro(st, 1, (const cpx<T>*)in, (cpx<T>*)out, f0, nf);
  }
  int slab_pass_peer(cudaStream_t st, int pass, int f0, int nf, const void* in, void* const* peers) override {
    if (D != 3 || nranks <= 1) return fail(EXB_EINVAL, "exb_slab_pass_peer needs a multi-rank 3-D slab plan");
    if (nranks > EXB_MAX_PEERS) return fail(EXB_EUNSUPPORTED, "at most %d ranks", EXB_MAX_PEERS);
    if (!peers) return fail(EXB_EINVAL, "null peer table");
    if (pass == EXB_SLAB_COL0_INV_PRO) {
      if (f0 < 0 || nf < 1 || f0 + nf > P.n_inv) return fail(EXB_EINVAL, "field range out of bounds");
      return col_inv_pro(st, 1, (const cpx<T>*)in, nullptr, f0, nf, peers);
    }
    if (pass == (EXB_SLAB_COL1_FWD_NL | EXB_SLAB_SEGMENTED) || pass == (EXB_SLAB_COL1_FWD | EXB_SLAB_SEGMENTED)) {
      if (f0 < 0 || nf < 1) return fail(EXB_EINVAL, "field range out of bounds");
      const int prune = pass == (EXB_SLAB_COL1_FWD_NL | EXB_SLAB_SEGMENTED) ? (PRUNE_COLS | PRUNE_OUT_ROWS) : 0;
      return col_plain<-1>(st, 1, 1, nf, (const cpx<T>*)in, nullptr, prune, true, peers, (long long)f0 * (prune ? Mf : M),
                           prune != 0);
    }
    return fail(EXB_EINVAL, "exb_slab_pass_peer: pass must be COL0_INV_PRO or a segmented COL1_FWD pass");
  }
  int field_pitch() const override { return Nhp; }
  void nl_fields(int* ni, int* nf) const override {
    *ni = P.n_inv;
    *nf = P.n_fwd;
  }
  int slab_pass(cudaStream_t st, int pass, int nfields, int stage, const void* in, void* out, const void* U,
                void* OUT, void* const* S) override {
    if (D != 3) return fail(EXB_EINVAL, "exb_slab_pass needs a 3-D plan");
    const long long rows_stride_c = M;  // per-field stride, half-complex
    const bool seg = (pass & EXB_SLAB_SEGMENTED) != 0;
    pass &= ~EXB_SLAB_SEGMENTED;
    if (seg && (nranks <= 1 || (pass != EXB_SLAB_COL1_FWD && pass != EXB_SLAB_COL1_INV &&
                                pass != EXB_SLAB_COL1_FWD_NL && pass != EXB_SLAB_COL1_INV_NL)))
      return fail(EXB_EINVAL, "EXB_SLAB_SEGMENTED applies to the axis-1 passes of a multi-rank slab plan");
    switch (pass) {
      case EXB_SLAB_ROW_R2C:
        return row_pass(st, ROW_R2C, 1, nfields, nfields, in, (long long)nfields * G, out, (long long)nfields * rows_stride_c);
      case EXB_SLAB_ROW_C2R:
        return row_pass(st, ROW_C2R, 1, nfields, nfields, in, (long long)nfields * rows_stride_c, out, (long long)nfields * G);
      case EXB_SLAB_COL1_FWD:
        return col_plain<-1>(st, 1, 1, nfields, (const cpx<T>*)in, (cpx<T>*)out, 0, seg);
      case EXB_SLAB_COL1_INV:
        return col_plain<+1>(st, 1, 1, nfields, (const cpx<T>*)in, (cpx<T>*)out, 0, seg);
      case EXB_SLAB_COL1_FWD_NL:   // (field buffers: pitch exb_plan_field_pitch)
        return col_plain<-1>(st, 1, 1, nfields, (const cpx<T>*)in, (cpx<T>*)out, PRUNE_COLS | PRUNE_OUT_ROWS, seg, nullptr,
                             0, true);
      case EXB_SLAB_COL1_INV_NL:
        return col_plain<+1>(st, 1, 1, nfields, (const cpx<T>*)in, (cpx<T>*)out, PRUNE_COLS | PRUNE_IN_ROWS, seg, nullptr,
                             0, true);
      case EXB_SLAB_COL0_FWD:
        return col_plain<-1>(st, 0, 1, nfields, (const cpx<T>*)in, (cpx<T>*)out);
      case EXB_SLAB_COL0_INV:
        return col_plain<+1>(st, 0, 1, nfields, (const cpx<T>*)in, (cpx<T>*)out);
      case EXB_SLAB_COL0_INV_PRO:
        return col_inv_pro(st, 1, (const cpx<T>*)in, (cpx<T>*)out);
      case EXB_SLAB_ROW_NL:
        return row_pass(st, ROW_NL, 1, P.n_inv, P.n_fwd, in, (long long)P.n_inv * Mf, out, (long long)P.n_fwd * Mf,
                        Nhp, Nhp);
      case EXB_SLAB_COL0_FWD_EPI: {
        if (K.order == 0) {
          long long total = (long long)C * M;
          int grid = (int)((total + 255) / 256 < (long long)sm_count * 16 ? (total + 255) / 256 : (long long)sm_count * 16);
          etdrk0_kernel<T><<<grid, 256, 0, st>>>(K, C, total, (const cpx<T>*)U, (cpx<T>*)OUT);
          ++launches;
          CUDA_OK(cudaGetLastError());
          return EXB_OK;
        }
        StateBufs<T> sb;
        sb.U = (const cpx<T>*)U;
        sb.OUT = (cpx<T>*)OUT;
        for (int i = 0; i < 4; ++i) sb.S[i] = S ? (cpx<T>*)S[i] : nullptr;
        return col_fwd(st, 1, COL_FWD_EPI, stage, (const cpx<T>*)in, nullptr, sb);
      }
      default:
        return fail(EXB_EINVAL, "unknown slab pass %d", pass);
    }
  }

  // ------------------------------------------------------------------ entry points
  int need_ws(void* ws, int64_t batch) {
    if (workspace_bytes(batch) > 0 && !ws) return fail(EXB_EINVAL, "workspace is NULL but %zu bytes are required", workspace_bytes(batch));
    return EXB_OK;
  }

  int fft(cudaStream_t st, int64_t batch, int channels, const void* u, void* uh, void* ws) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (channels < 1) return fail(EXB_EINVAL, "channels must be >= 1");
    if (D == 1) return launch_1d(st, OP1_FFT, batch * channels, 1, u, uh, 0, 0, 0);
    return fft_nd(st, batch, channels, (const T*)u, (long long)channels * G, (cpx<T>*)uh);
  }

  int ifft(cudaStream_t st, int64_t batch, int channels, const void* uh, void* u, void* ws) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (channels < 1) return fail(EXB_EINVAL, "channels must be >= 1");
    if (D == 1) return launch_1d(st, OP1_IFFT, batch * channels, 1, uh, u, 0, 0, 0);
    if (channels > winv_fields()) return fail(EXB_EUNSUPPORTED, "ifft supports at most %d channels with this plan", winv_fields());
    int rc = need_ws(ws, batch);
    if (rc) return rc;
    Ws w = carve(ws, batch);
    return ifft_nd(st, batch, channels, (const cpx<T>*)uh, w.Winv, (T*)u, (long long)channels * G);
  }

  int ic_shape(cudaStream_t st, int64_t nfields, void* uh, int kind, double param, double domain_extent,
               double dc_value) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (nfields < 1 || (kind != 0 && kind != 1)) return fail(EXB_EINVAL, "exb_ic_shape: bad arguments");
    if (kind == 1 && !(domain_extent > 0)) return fail(EXB_EINVAL, "exb_ic_shape: domain_extent must be > 0");
    const long long total = (long long)nfields * M;
    const double two_pi = 6.283185307179586476925286766559;
    ic_shape_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (cpx<T>*)uh, D, N, Nh, M, total, kind, (T)param, (T)(kind == 1 ? two_pi / domain_extent : 0.0), (T)dc_value);
    CUDA_OK(cudaGetLastError());
    ++launches;
    return EXB_OK;
  }

  int fourier_sums(cudaStream_t st, int64_t nfields, const void* xh, double p, int low, int high, double order,
                   double domain_extent, double* out) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (nfields < 1 || nfields > 65535 || !xh || !out) return fail(EXB_EINVAL, "exb_fourier_sums: bad arguments");
    if (!(domain_extent > 0)) return fail(EXB_EINVAL, "exb_fourier_sums: domain_extent must be > 0");
    FourierSumParams<T> q;
    q.xh = (const cpx<T>*)xh;
    q.out = out;
    q.D = D;
    q.N = N;
    q.Nh = Nh;
    q.ncomp = order >= 0 ? D : 1;
    q.M = M;
    long long nchunk = (M + 4095) / 4096;
    const long long target = std::max<long long>(1, (8ll * sm_count + nfields - 1) / nfields);
    if (nchunk > target) nchunk = target;
    q.chunk = (M + nchunk - 1) / nchunk;
    q.p = (T)p;
    q.pi = p == 1.0 ? 1 : (p == 2.0 ? 2 : 0);
    q.filter = (low >= 0 || high >= 0) ? 1 : 0;
    q.low = low >= 0 ? low : 0;
    q.high = high >= 0 ? high : N / 2 + 1;
    q.order = (T)order;
    q.two_pi_over_L = (T)(6.283185307179586476925286766559 / domain_extent);
    CUDA_OK(cudaMemsetAsync(out, 0, (size_t)nfields * q.ncomp * sizeof(double), st));
    dim3 grid((unsigned)nchunk, (unsigned)nfields);
    fourier_sums_kernel<T><<<grid, 256, 0, st>>>(q);
    CUDA_OK(cudaGetLastError());
    ++launches;
    return EXB_OK;
  }

  int leray(cudaStream_t st, int64_t nfields, const void* uh, void* out, double domain_extent) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (nfields < 1 || !uh || !out) return fail(EXB_EINVAL, "exb_leray: bad arguments");
    if (!(domain_extent > 0)) return fail(EXB_EINVAL, "exb_leray: domain_extent must be > 0");
    const long long total = (long long)nfields * M;
    leray_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const cpx<T>*)uh, (cpx<T>*)out, D, N, Nh, M, total, (T)(6.283185307179586476925286766559 / domain_extent));
    CUDA_OK(cudaGetLastError());
    ++launches;
    return EXB_OK;
  }

  int derivative(cudaStream_t st, int64_t nfields, const void* uh, void* out, int order,
                 double domain_extent) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (nfields < 1 || order < 0 || !uh || !out || uh == out) return fail(EXB_EINVAL, "exb_derivative: bad arguments");
    if (!(domain_extent > 0)) return fail(EXB_EINVAL, "exb_derivative: domain_extent must be > 0");
    const long long total = (long long)nfields * M;
    derivative_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const cpx<T>*)uh, (cpx<T>*)out, D, N, Nh, M, total, order, (T)(6.283185307179586476925286766559 / domain_extent));
    CUDA_OK(cudaGetLastError());
    ++launches;
    return EXB_OK;
  }

  int spectrum(cudaStream_t st, int64_t nfields, const void* uh, void* out, int power, int average,
               void* counts) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (nfields < 1) return fail(EXB_EINVAL, "nfields must be >= 1");
    if (average && !counts) return fail(EXB_EINVAL, "radial average needs the counts scratch buffer");
    SpectrumParams<T> sp;
    sp.uh = (const cpx<T>*)uh;
    sp.out = (T*)out;
    sp.counts = average ? (unsigned*)counts : nullptr;
    sp.D = D;
    sp.N = N;
    sp.Nh = Nh;
    sp.M = M;
    sp.power = power;
    // enough CTAs to fill the GPU, at least 4096 modes each
    long long nchunk = (M + 4095) / 4096;
    const long long target = std::max<long long>(1, (8ll * sm_count + nfields - 1) / nfields);
    if (nchunk > target) nchunk = target;
    sp.chunk = (M + nchunk - 1) / nchunk;
    CUDA_OK(cudaMemsetAsync(out, 0, (size_t)nfields * Nh * sizeof(T), st));
    if (average) CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)Nh * sizeof(unsigned), st));
    dim3 grid((unsigned)nchunk, (unsigned)nfields);
    const size_t smem = (size_t)Nh * (sizeof(T) + sizeof(unsigned));
    spectrum_kernel<T><<<grid, 256, smem, st>>>(sp);
    CUDA_OK(cudaGetLastError());
    ++launches;
    if (average) {
      const long long tot = nfields * Nh;
      spectrum_average_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((T*)out, (const unsigned*)counts, Nh, nfields);
      CUDA_OK(cudaGetLastError());
      ++launches;
    }
    return EXB_OK;
  }

  int nonlinear(cudaStream_t st, int64_t batch, const void* uh, void* out, void* ws) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (P.kind == EXB_NL_ZERO) {
      CUDA_OK(cudaMemsetAsync(out, 0, (size_t)batch * C * M * sizeof(cpx<T>), st));
      return EXB_OK;
    }
    if (D == 1) return launch_1d(st, OP1_NL, batch, C, uh, out, 0, 0, 0);
    int rc = need_ws(ws, batch);
    if (rc) return rc;
    Ws w = carve(ws, batch);
    rc = nl_front_nd(st, batch, (const cpx<T>*)uh, w);
    if (rc) return rc;
    StateBufs<T> sb;
    memset(&sb, 0, sizeof(sb));
    return col_fwd(st, batch, COL_FWD_NL, 0, w.Wfwd, (cpx<T>*)out, sb);
  }

  int step_fourier(cudaStream_t st, int64_t batch, const void* in, void* out, void* ws) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (int rct = set_table_rep(batch)) return rct;
    if (D == 1) return launch_1d(st, OP1_STEP_FOURIER, batch, C, in, out, 0, 1, 0);
    int rc = need_ws(ws, batch);
    if (rc) return rc;
    Ws w = carve(ws, batch);
    return step_fourier_nd(st, batch, (const cpx<T>*)in, (cpx<T>*)out, w);
  }

  int step(cudaStream_t st, int64_t batch, const void* in, void* out, void* ws) override {
    return rollout(st, batch, 1, 1, EXB_ROLLOUT_FINAL_ONLY, in, out, ws);
  }

  int rollout(cudaStream_t st, int64_t batch, int64_t n_saved, int substeps, unsigned flags, const void* u0,
              void* out, void* ws) override {
    if (nranks > 1) return fail(EXB_EINVAL, "slab plans only support exb_slab_pass");
    if (n_saved < 0 || substeps < 1) return fail(EXB_EINVAL, "n_saved must be >= 0 and substeps >= 1");
    if (int rct = set_table_rep(batch)) return rct;
    const bool include_init = flags & EXB_ROLLOUT_INCLUDE_INIT;
    const bool layout_tb = flags & EXB_ROLLOUT_LAYOUT_TB;
    const bool final_only = flags & EXB_ROLLOUT_FINAL_ONLY;
    const bool spectral_carry = flags & EXB_ROLLOUT_SPECTRAL_CARRY;
    if (final_only && n_saved < 1) return fail(EXB_EINVAL, "FINAL_ONLY needs n_saved >= 1");
    if (D == 1) {
      if (n_saved == 0 && !include_init) return EXB_OK;
      return launch_1d(st, OP1_ROLLOUT, batch, C, u0, out, n_saved, substeps, flags);
    }
    int rc = need_ws(ws, batch);
    if (rc) return rc;
    // Optional trajectory-major chunking (EXB_ND_CHUNK=n): trajectories are independent, so the batch can
    // be processed n at a time through ALL steps (bounds the live workspace; see init() for why it is off).
    const long long cb = chunk_batch(batch);
    const long long fsz0 = (long long)C * G;
    const long long Tn0 = final_only ? 1 : n_saved + (include_init ? 1 : 0);
    for (long long b0 = 0; b0 < batch; b0 += cb) {
      const long long nb = batch - b0 < cb ? batch - b0 : cb;
      const T* u0c = (const T*)u0 + (size_t)b0 * fsz0;
      T* outc = (T*)out + (size_t)b0 * ((final_only || layout_tb) ? fsz0 : Tn0 * fsz0);
      forcing_b0 = b0;
      rc = rollout_nd_chunk(st, nb, batch, n_saved, substeps, flags, u0c, outc, ws);
      if (rc) return rc;
    }
    return EXB_OK;
  }

  // number of trajectories processed together (N-D).  EXB_ND_CHUNK overrides (0 = whole batch).
  long long chunk_batch(long long batch) const {
    if (table_sets > 1) return batch;   // (table sets are indexed by the position in the whole batch)
    if (const char* e = getenv("EXB_ND_CHUNK")) {
      long long v = atoll(e);
      return v <= 0 ? batch : (v < batch ? v : batch);
    }
    return nd_chunk > 0 && nd_chunk < batch ? nd_chunk : batch;
  }
  long long nd_chunk = 0;  // set in init()
  long long forcing_b0 = 0;  // first trajectory of the chunk being processed (forcing offset)

  int rollout_forced(cudaStream_t st, int64_t batch, int64_t n_saved, int substeps, unsigned flags, const void* u0,
                     void* out, void* ws, const void* fhat, int64_t fstep, int64_t fbatch, double scale) override {
    if (!fhat) return rollout(st, batch, n_saved, substeps, flags, u0, out, ws);
    if (substeps != 1) return fail(EXB_EINVAL, "a forcing needs substeps = 1 (one forcing per ETDRK step)");
    if (flags & EXB_ROLLOUT_SPECTRAL_CARRY)
      return fail(EXB_EINVAL, "a forcing is added to the physical-space carry: no SPECTRAL_CARRY");
    frc.f = (const cpx<T>*)fhat;
    frc.step = fstep;
    frc.batch = fbatch;
    frc.scale = (T)scale;
    const int rc = rollout(st, batch, n_saved, substeps, flags, u0, out, ws);
    frc = Forcing();
    return rc;
  }

  int rollout_nd_chunk(cudaStream_t st, long long batch, long long batch_total, int64_t n_saved, int substeps,
                       unsigned flags, const T* u0, T* out, void* ws) {
    const bool include_init = flags & EXB_ROLLOUT_INCLUDE_INIT;
    const bool layout_tb = flags & EXB_ROLLOUT_LAYOUT_TB;
    const bool final_only = flags & EXB_ROLLOUT_FINAL_ONLY;
    const bool spectral_carry = flags & EXB_ROLLOUT_SPECTRAL_CARRY;
    int rc;
    Ws w = carve(ws, batch);
    const long long fsz = (long long)C * G;
    const long long Tn = final_only ? 1 : n_saved + (include_init ? 1 : 0);
    T* o = out;
    auto slot = [&](long long s, long long& bs) -> T* {
      if (final_only) {
        bs = fsz;
        return o;
      }
      if (layout_tb) {
        bs = fsz;
        return o + (size_t)s * batch_total * fsz;
      }
      bs = Tn * fsz;
      return o + (size_t)s * fsz;
    };
    if (include_init && !final_only) {
      long long bs;
      T* dst = slot(0, bs);
      long long total = batch * fsz;
      int grid = (int)((total + 255) / 256 < (long long)sm_count * 16 ? (total + 255) / 256 : (long long)sm_count * 16);
      copy_batched_kernel<T><<<grid, 256, 0, st>>>((const T*)u0, fsz, dst, bs, fsz, batch);
      ++launches;
      CUDA_OK(cudaGetLastError());
    }
    rc = fft_nd(st, batch, C, (const T*)u0, fsz, w.Uh);
    if (rc) return rc;
    T* phys_scratch = (T*)w.Wfwd;  // >= C*G reals per batch element (M*2 >= G)
    bool have_winv = false;
    for (long long s = 0; s < n_saved; ++s) {
      if (frc.f) {  // ForcedStepper: u_hat += dt * f_hat of step s
        const long long per = (long long)C * M, total = per * batch;
        int grid = (int)((total + 255) / 256 < (long long)sm_count * 16 ? (total + 255) / 256 : (long long)sm_count * 16);
        add_forcing_kernel<T><<<grid, 256, 0, st>>>(w.Uh, frc.f + s * frc.step + forcing_b0 * frc.batch, per, frc.batch,
                                                    batch, frc.scale);
        ++launches;
        CUDA_OK(cudaGetLastError());
      }
      for (int sub = 0; sub < substeps; ++sub) {
        // the next ETDRK step starts from this step's spectral result when another sub-step follows, or (spectral
        // carry) when the next saved step does without the physical-space round trip; a forcing changes the state
        // in between, so no fusion across it
        const bool more_sub = sub + 1 < substeps;
        const bool more_saved = s + 1 < n_saved && spectral_carry && (final_only) && !frc.f;
        const bool want_next = more_sub || more_saved;
        rc = step_fourier_nd(st, batch, w.Uh, w.Uh, w, have_winv, want_next);
        if (rc) return rc;
        have_winv = want_next && can_fuse_prologue();
      }
      const bool last = s == n_saved - 1;
      const bool store = !final_only || last;
      if (!store && spectral_carry) continue;
      long long bs = fsz;
      T* dst = phys_scratch;
      if (store) dst = slot(final_only ? 0 : s + (include_init ? 1 : 0), bs);
      rc = ifft_nd(st, batch, C, w.Uh, w.Winv, dst, bs);
      if (rc) return rc;
      if (last) break;
      if (!spectral_carry) {
        rc = fft_nd(st, batch, C, dst, bs, w.Uh);
        if (rc) return rc;
      }
    }
    return EXB_OK;
  }
};

// ---------------------------------------------------------------------------------- C ABI
template <class T>
static int metric_sums_t(cudaStream_t st, int64_t nfields, int64_t npoints, const void* a, const void* b, double p,
                         double* out) {
  int dev = 0, sms = 0;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long nchunk = (npoints + 8191) / 8192;
  const long long target = std::max<long long>(1, (8ll * sms + nfields - 1) / nfields);
  if (nchunk > target) nchunk = target;
  const long long chunk = (npoints + nchunk - 1) / nchunk;
  const int pi = p == 1.0 ? 1 : (p == 2.0 ? 2 : 0);
  CUDA_OK(cudaMemsetAsync(out, 0, (size_t)nfields * 4 * sizeof(double), st));
  dim3 grid((unsigned)nchunk, (unsigned)nfields);
  metric_sums_kernel<T><<<grid, 256, 0, st>>>((const T*)a, (const T*)b, npoints, chunk, (T)p, pi, out);
  CUDA_OK(cudaGetLastError());
  return EXB_OK;
}

template <class T>
static int ic_normalize_t(cudaStream_t st, int64_t nfields, int64_t npoints, void* u, int zero_mean, int std_one,
                          int max_one, double* stats) {
  int dev = 0, sms = 0;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long nchunk = (npoints + 8191) / 8192;
  const long long target = std::max<long long>(1, (8ll * sms + nfields - 1) / nfields);
  if (nchunk > target) nchunk = target;
  const long long chunk = (npoints + nchunk - 1) / nchunk;
  const long long total = nfields * npoints;
  dim3 grid((unsigned)nchunk, (unsigned)nfields);
  CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)nfields * 4 * sizeof(double), st));
  ic_shift_kernel<T><<<(unsigned)((nfields + 255) / 256), 256, 0, st>>>((const T*)u, npoints, nfields, stats);
  ic_stats_kernel<T><<<grid, 256, 0, st>>>((const T*)u, npoints, chunk, stats);
  if (max_one) ic_maxabs_kernel<T><<<grid, 256, 0, st>>>((const T*)u, npoints, chunk, zero_mean, stats);
  ic_apply_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((T*)u, npoints, total, zero_mean, std_one,
                                                                     max_one, stats);
  CUDA_OK(cudaGetLastError());
  return EXB_OK;
}

extern "C" {

const char* exb_last_error(void) { return g_err.c_str(); }
const char* exb_version(void) { return "exb 0.1 (sm_100a)"; }

int exb_plan_create(const exb_desc* desc, exb_plan** out) {
  if (!desc || !out) return fail(EXB_EINVAL, "null argument");
  if (desc->struct_size != (int32_t)sizeof(exb_desc))
    return fail(EXB_EINVAL, "exb_desc size mismatch (got %d, expected %zu)", desc->struct_size, sizeof(exb_desc));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(EXB_ECUDA, "no CUDA device available (%s); libexb has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  *out = nullptr;
  int rc;
  if (desc->dtype == EXB_F32) {
    auto* p = new PlanImpl<float>();
    rc = p->init(*desc);
    if (rc) {
      delete p;
      return rc;
    }
    *out = p;
  } else if (desc->dtype == EXB_F64) {
    auto* p = new PlanImpl<double>();
    rc = p->init(*desc);
    if (rc) {
      delete p;
      return rc;
    }
    *out = p;
  } else {
    return fail(EXB_EINVAL, "unknown dtype %d", desc->dtype);
  }
  return EXB_OK;
}

void exb_plan_destroy(exb_plan* plan) { delete plan; }

size_t exb_workspace_bytes(const exb_plan* plan, int64_t batch) { return plan ? plan->workspace_bytes(batch) : 0; }

int exb_fft(exb_plan* plan, void* stream, int64_t batch, int32_t channels, const void* u, void* u_hat, void* ws) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->fft((cudaStream_t)stream, batch, channels, u, u_hat, ws);
}
int exb_ifft(exb_plan* plan, void* stream, int64_t batch, int32_t channels, const void* u_hat, void* u, void* ws) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->ifft((cudaStream_t)stream, batch, channels, u_hat, u, ws);
}
int exb_nonlinear_fun(exb_plan* plan, void* stream, int64_t batch, const void* u_hat, void* out_hat, void* ws) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->nonlinear((cudaStream_t)stream, batch, u_hat, out_hat, ws);
}
int exb_step_fourier(exb_plan* plan, void* stream, int64_t batch, const void* in, void* out, void* ws) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->step_fourier((cudaStream_t)stream, batch, in, out, ws);
}
int exb_step(exb_plan* plan, void* stream, int64_t batch, const void* in, void* out, void* ws) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->step((cudaStream_t)stream, batch, in, out, ws);
}
int exb_rollout(exb_plan* plan, void* stream, int64_t batch, int64_t n_saved, int32_t substeps, uint32_t flags,
                const void* u0, void* out, void* ws) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->rollout((cudaStream_t)stream, batch, n_saved, substeps, flags, u0, out, ws);
}
int exb_rollout_forced(exb_plan* plan, void* stream, int64_t batch, int64_t n_saved, int32_t substeps, uint32_t flags,
                       const void* u0, void* out, void* ws, const void* forcing_hat, int64_t step_stride,
                       int64_t batch_stride, double scale) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->rollout_forced((cudaStream_t)stream, batch, n_saved, substeps, flags, u0, out, ws, forcing_hat, step_stride,
                              batch_stride, scale);
}
int64_t exb_launch_count(const exb_plan* plan) { return plan ? plan->launches : 0; }
int exb_plan_fused_ok(const exb_plan* plan) { return plan ? plan->fused_ok() : 0; }
int exb_metric_sums(void* stream, int32_t dtype, int64_t nfields, int64_t npoints, const void* a, const void* b,
                    double p, double* out) {
  if (nfields < 1 || npoints < 1 || !a || !out) return fail(EXB_EINVAL, "exb_metric_sums: bad arguments");
  if (nfields > 65535) return fail(EXB_EINVAL, "exb_metric_sums: at most 65535 fields per call");
  if (dtype == EXB_F32) return metric_sums_t<float>((cudaStream_t)stream, nfields, npoints, a, b, p, out);
  if (dtype == EXB_F64) return metric_sums_t<double>((cudaStream_t)stream, nfields, npoints, a, b, p, out);
  return fail(EXB_EINVAL, "exb_metric_sums: unknown dtype");
}
int exb_ic_shape(exb_plan* plan, void* stream, int64_t nfields, void* u_hat, int32_t kind, double param,
                 double domain_extent, double dc_value) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->ic_shape((cudaStream_t)stream, nfields, u_hat, kind, param, domain_extent, dc_value);
}
int exb_ic_normalize(void* stream, int32_t dtype, int64_t nfields, int64_t npoints, void* u, int32_t zero_mean,
                     int32_t std_one, int32_t max_one, double* stats) {
  if (nfields < 1 || npoints < 1 || !u || !stats) return fail(EXB_EINVAL, "exb_ic_normalize: bad arguments");
  if (nfields > 65535) return fail(EXB_EINVAL, "exb_ic_normalize: at most 65535 fields per call");
  if (dtype == EXB_F32)
    return ic_normalize_t<float>((cudaStream_t)stream, nfields, npoints, u, zero_mean, std_one, max_one, stats);
  if (dtype == EXB_F64)
    return ic_normalize_t<double>((cudaStream_t)stream, nfields, npoints, u, zero_mean, std_one, max_one, stats);
  return fail(EXB_EINVAL, "exb_ic_normalize: unknown dtype");
}
int exb_fourier_sums(exb_plan* plan, void* stream, int64_t nfields, const void* x_hat, double p, int32_t low,
                     int32_t high, double derivative_order, double domain_extent, double* out) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->fourier_sums((cudaStream_t)stream, nfields, x_hat, p, low, high, derivative_order, domain_extent, out);
}
int exb_leray(exb_plan* plan, void* stream, int64_t nfields, const void* u_hat, void* out_hat, double domain_extent) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->leray((cudaStream_t)stream, nfields, u_hat, out_hat, domain_extent);
}
int exb_derivative(exb_plan* plan, void* stream, int64_t nfields, const void* u_hat, void* out_hat, int32_t order,
                   double domain_extent) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->derivative((cudaStream_t)stream, nfields, u_hat, out_hat, order, domain_extent);
}
int exb_spectrum(exb_plan* plan, void* stream, int64_t nfields, const void* u_hat, void* out, int32_t power,
                 int32_t average, void* counts) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->spectrum((cudaStream_t)stream, nfields, u_hat, out, power, average, counts);
}
int exb_slab_pass(exb_plan* plan, void* stream, int32_t pass, int32_t nfields, int32_t stage, const void* in,
                  void* out, const void* U, void* OUT, void* const* S) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->slab_pass((cudaStream_t)stream, pass, nfields, stage, in, out, U, OUT, S);
}
int exb_slab_pass_peer(exb_plan* plan, void* stream, int32_t pass, int32_t field0, int32_t nfields, const void* in,
                       void* const* peer_out) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->slab_pass_peer((cudaStream_t)stream, pass, field0, nfields, in, peer_out);
}
int exb_slab_inv_pro_fields(exb_plan* plan, void* stream, int32_t field0, int32_t nfields, const void* in, void* out) {
  if (!plan) return fail(EXB_EINVAL, "null plan");
  return plan->slab_inv_pro_fields((cudaStream_t)stream, field0, nfields, in, out);
}
int exb_plan_field_pitch(const exb_plan* plan) { return plan ? plan->field_pitch() : 0; }
int exb_plan_nl_fields(const exb_plan* plan, int32_t* n_inv, int32_t* n_fwd) {
  if (!plan || !n_inv || !n_fwd) return fail(EXB_EINVAL, "null argument");
  plan->nl_fields(n_inv, n_fwd);
  return EXB_OK;
}

}  // extern "C"
