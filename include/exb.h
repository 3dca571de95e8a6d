/*
 * exb.h -- C ABI of libexb.so, the B200-native (sm_100a) ETDRK spectral
 * time-stepping core behind the exponax stepper API.
 *
 * The reference (Ceyron/exponax) has no FFI layer of its own: the hot path is
 * Python on jax.numpy.  Each entry point below replaces the body of one
 * reference method; a maintainer binds them with ctypes / jax.ffi as shown in
 * INTEGRATION.md.
 *
 *   exb_fft            <- exponax/_spectral.py:614-656   (fft = rfftn, unnormalised)
 *   exb_ifft           <- exponax/_spectral.py:659-721   (ifft = irfftn, 1/N^D)
 *   exb_nonlinear_fun  <- exponax/nonlin_fun/*.py  __call__(u_hat)
 *   exb_step_fourier   <- exponax/_base_stepper.py:222-239 + etdrk/_etdrk_{0..4}.py step_fourier
 *   exb_step           <- exponax/_base_stepper.py:201-220  (fft -> step_fourier -> ifft)
 *   exb_rollout        <- exponax/_utils.py:92-254 (rollout / repeat) composed with
 *                         exponax/_repeated_stepper.py:56-102 (sub-steps with spectral carry)
 *   exb_spectrum       <- exponax/_spectral.py:866-1030  (get_spectrum after its fft)
 *   exb_ic_shape / exb_ic_normalize <- exponax/ic/_truncated_fourier_series.py:65-100,
 *                         ic/_gaussian_random_field.py:64-93, ic/_base_ic.py:16-33
 *   exb_derivative     <- exponax/_spectral.py:724-792  (derivative between its fft and ifft)
 *   exb_fourier_sums   <- exponax/metrics/_fourier.py:15-140 (fourier_aggregator after its fft)
 *   exb_metric_sums    <- exponax/metrics/_spatial.py:8-196, metrics/_correlation.py:6-60
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer owned by the caller (C-contiguous,
 *     reference layout: state (batch, C, N, .., N) real; spectral state
 *     (batch, C, N, .., N/2+1) complex interleaved re,im);
 *   - `stream` is a cudaStream_t passed as void*; the library only enqueues work
 *     on it -- no allocation, no synchronisation, no other stream -- so calls can
 *     be captured in CUDA graphs / XLA command buffers;
 *   - `workspace` is a caller-owned device buffer of at least
 *     exb_workspace_bytes(plan, batch) bytes (may be NULL when that is 0);
 *   - device memory is allocated only in exb_plan_create (twiddles, coefficient
 *     tables) and released in exb_plan_destroy;
 *   - return value 0 on success, <0 on error; exb_last_error() gives the
 *     thread-local message.  There is NO CPU fallback: without a CUDA device
 *     every compute entry point fails with EXB_ECUDA.
 */
#ifndef EXB_H_
#define EXB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXB_OK 0
#define EXB_EINVAL (-1)
#define EXB_EUNSUPPORTED (-2)
#define EXB_ECUDA (-3)

#define EXB_F32 0
#define EXB_F64 1

/* nonlinear function kinds (exponax/nonlin_fun/) */
#define EXB_NL_ZERO 0            /* _zero.py:34-38 */
#define EXB_NL_CONVECTION 1      /* _convection.py:104-245 (4 variants via flags) */
#define EXB_NL_GRADIENT_NORM 2   /* _gradient_norm.py:78-101 */
#define EXB_NL_POLYNOMIAL 3      /* _polynomial.py:64-76 */
#define EXB_NL_VORTICITY_2D 4    /* _vorticity_convection.py:78-99, 166-182 */
#define EXB_NL_PROJECTED_3D 5    /* _projected_convection.py:114-136, 202-226 (+ _leray.py:114-136) */
#define EXB_NL_GENERAL 6         /* _general_nonlinear.py:78-119 */
#define EXB_NL_GRAY_SCOTT 7      /* stepper/reaction/_gray_scott.py:30-45   (general_scales = feed, kill) */
#define EXB_NL_CAHN_HILLIARD 8   /* stepper/reaction/_cahn_hilliard.py:29-37 (nl_scale * laplace * F[u^3]) */

#define EXB_MAX_POLY 8

/* exb_rollout flags */
#define EXB_ROLLOUT_INCLUDE_INIT 1u   /* prepend u0 (rollout(include_init=True)) */
#define EXB_ROLLOUT_LAYOUT_TB 2u      /* output (T, B, C, ..) instead of (B, T, C, ..) */
#define EXB_ROLLOUT_FINAL_ONLY 4u     /* ex.repeat: write only the last state, shape (B, C, ..) */
#define EXB_ROLLOUT_SPECTRAL_CARRY 8u /* keep the carry in Fourier space between saved steps
                                         (Hermitian-projected; differs from the reference's
                                         ifft->fft round trip at rounding level only) */

typedef struct exb_plan exb_plan;

typedef struct exb_desc {
  int32_t struct_size;       /* sizeof(exb_desc), ABI guard */
  int32_t num_spatial_dims;  /* D in {1,2,3} */
  int32_t num_points;        /* N per axis */
  int32_t num_channels;      /* C */
  int32_t lin_channels;      /* E in {1, C}: leading extent of the coefficient arrays */
  int32_t order;             /* ETDRK order 0..4 */
  int32_t dtype;             /* EXB_F32 / EXB_F64 */
  int32_t dealias_kmax;      /* keep |k_d| <= kmax on every axis; <0: no dealiasing mask */
  double domain_extent;      /* L */
  /* nonlinear function */
  int32_t nl_kind;
  int32_t single_channel;    /* convection */
  int32_t conservative;      /* convection */
  int32_t zero_mode_fix;     /* gradient norm / general */
  double nl_scale;           /* convection scale / gradient-norm scale / vorticity convection scale */
  int32_t n_poly;            /* polynomial: number of coefficients (<= EXB_MAX_POLY) */
  int32_t table_sets;        /* stepper ENSEMBLES (docs/examples/performance_hints.ipynb, eqx.filter_vmap over constructor
                                arguments): > 1 = the coefficient arrays carry a leading axis of that many table sets,
                                (table_sets, E, modes); trajectory b of a batch uses set b / (batch / table_sets).
                                0 or 1 = one set shared by the whole batch */
  double poly[EXB_MAX_POLY];
  double general_scales[3];  /* general: (b0, b1, b2) as in GeneralNonlinearFun.scale_list */
  /* Kolmogorov injection: constant real value added (un-masked) to channel 0 of N(u) at one
     spectral index; has_injection = 0 disables */
  int32_t has_injection;
  int32_t injection_index[3];
  double injection_value;
  /* ETDRK coefficient tables, HOST pointers, copied to the device in exb_plan_create.
     exp_term/half_exp_term: complex (E, modes); coef[i]: real (E, modes); unused ones NULL.
     modes = N^(D-1) * (N/2+1), C-order.   (etdrk/_base_etdrk.py:61-62, _etdrk_{1..4}.py) */
  const void *exp_term;
  const void *half_exp_term;
  const void *coef[6];
  /* slab decomposition of ONE 3-D field over slab_nranks processes (0/1: none).  Physical slab:
     (C, N/P, N, N), x-planes [rank*N/P, (rank+1)*N/P); spectral slab: (C, N, N/P, N/2+1), axis-1
     wavenumber indices [rank*N/P, (rank+1)*N/P).  Coefficient tables are the LOCAL spectral
     slices (E, N, N/P, N/2+1).  Only exb_slab_pass may be used with such a plan. */
  int32_t slab_nranks;
  int32_t slab_rank;
  /* non-zero: exp_term / half_exp_term / coef[] are DEVICE pointers that stay valid for the life of
     the plan (tables built on the GPU for grids whose tables do not fit the host comfortably) */
  int32_t tables_on_device;
  int32_t slab_cyclic;       /* slab plans: 0 = rank r owns the contiguous axis-1 indices [r N/P, (r+1) N/P); 1 = rank r owns
                                the indices r, r + P, r + 2P, ... (cyclic).  With a dealiasing mask the cyclic distribution
                                gives every rank the same share of the wavenumbers inside the mask (block: at P = 8 two
                                ranks own nothing but dealiased modes and sit idle in the axis-0 passes and transposes) */
  int32_t reserved2;
  int32_t lin_matrix;        /* order 0 only: exp_term is a per-mode C x C matrix, (C*C, modes) with entry [i*C + j] =
                                contribution of input channel j to output channel i (lin_channels = C*C); used by the
                                Wave stepper, whose step is a 2 x 2 map of (h, v) per mode (exponax/stepper/_wave.py:175-197) */
} exb_desc;

/* thread-local description of the last error on this thread */
const char *exb_last_error(void);
/* library version string */
const char *exb_version(void);

/* create a plan on the CURRENT cuda device */
int exb_plan_create(const exb_desc *desc, exb_plan **out);
void exb_plan_destroy(exb_plan *plan);

/* bytes of caller-provided workspace the calls below need for `batch` trajectories */
size_t exb_workspace_bytes(const exb_plan *plan, int64_t batch);

/* u (batch, C, N..N) real  ->  u_hat (batch, C, N.., N/2+1) complex; channels: number of
   leading fields per batch element (pass plan's C for a state) */
int exb_fft(exb_plan *plan, void *stream, int64_t batch, int32_t channels, const void *u,
            void *u_hat, void *workspace);
/* inverse; u_hat is NOT modified */
int exb_ifft(exb_plan *plan, void *stream, int64_t batch, int32_t channels, const void *u_hat,
             void *u, void *workspace);

/* out_hat = N(u_hat) (dealiasing, derivative operators, projection, injection included) */
int exb_nonlinear_fun(exb_plan *plan, void *stream, int64_t batch, const void *u_hat,
                      void *out_hat, void *workspace);

/* one ETDRK step in Fourier space; u_hat_in and u_hat_out may alias */
int exb_step_fourier(exb_plan *plan, void *stream, int64_t batch, const void *u_hat_in,
                     void *u_hat_out, void *workspace);

/* one ETDRK step in physical space; u_in and u_out may alias */
int exb_step(exb_plan *plan, void *stream, int64_t batch, const void *u_in, void *u_out,
             void *workspace);

/* n_saved * substeps ETDRK steps starting from u0 (batch, C, N..N).  After every `substeps`
   steps the state is transformed to physical space and (unless FINAL_ONLY) stored:
     default        out (batch, T, C, N..N)      T = n_saved (+1 with INCLUDE_INIT)
     LAYOUT_TB      out (T, batch, C, N..N)
     FINAL_ONLY     out (batch, C, N..N)         (ex.repeat) */
int exb_rollout(exb_plan *plan, void *stream, int64_t batch, int64_t n_saved, int32_t substeps,
                uint32_t flags, const void *u0, void *out, void *workspace);

/* exb_rollout with a forcing: ForcedStepper / aux-taking rollouts (exponax/_forced_stepper.py:61-62, 85-86:
   `stepper.step(u + dt * f)`; exponax/_utils.py:137-163: `rollout(..., takes_aux=True)`), fused into the same
   kernels.  `forcing_hat` = exb_fft of the forcing (the transform is linear: fft(u + dt f) = u_hat + dt f_hat),
   complex (.., C, modes); before ETDRK step i of trajectory b the state receives
       u_hat += scale * forcing_hat[i * step_stride + b * batch_stride + ...]        (strides in complex elements)
   step_stride = 0: constant forcing (`constant_aux=True`); batch_stride = 0: one forcing shared by the batch;
   scale = dt.  substeps must be 1 and SPECTRAL_CARRY must not be set.  forcing_hat == NULL: plain exb_rollout.
   (A plan must not be driven from two host threads at the same time.) */
int exb_rollout_forced(exb_plan *plan, void *stream, int64_t batch, int64_t n_saved, int32_t substeps,
                       uint32_t flags, const void *u0, void *out, void *workspace, const void *forcing_hat,
                       int64_t step_stride, int64_t batch_stride, double scale);

/* ---- slab-decomposed 3-D transforms (one field too large for one GPU) ----
   One local pass of the distributed step; the caller (exponax_b200/_slab.py) performs the
   all-to-all transposes between passes with NCCL.  Layout A = physical slab (F, N/P, N, N/2+1 | N),
   layout B = spectral slab (F, N, N/P, N/2+1).  All buffers are caller-owned device memory. */
#define EXB_SLAB_ROW_R2C 0      /* in: real (nfields, N/P, N, N)       -> out: A half-complex        */
#define EXB_SLAB_ROW_C2R 1      /* in: A half-complex                  -> out: real                   */
#define EXB_SLAB_COL1_FWD 2     /* A, FFT along axis 1, in place when in == out                       */
#define EXB_SLAB_COL1_INV 3
#define EXB_SLAB_COL0_FWD 4     /* B, FFT along axis 0                                                */
#define EXB_SLAB_COL0_INV 5
#define EXB_SLAB_COL0_INV_PRO 6 /* in: stage input state B (C fields) -> out: n_inv fields B          */
#define EXB_SLAB_ROW_NL 7       /* in: n_inv fields A                  -> out: n_fwd fields A         */
#define EXB_SLAB_COL0_FWD_EPI 8 /* in: n_fwd fields B; ETDRK stage update on (U, OUT, S[0..3])        */
#define EXB_SLAB_COL1_INV_NL 9  /* COL1_INV with dealiasing-aware pruning (inside N(u) only)          */
#define EXB_SLAB_COL1_FWD_NL 10 /* COL1_FWD with dealiasing-aware pruning (inside N(u) only)          */
/* OR-ed into a COL1 pass id: layout A is the raw all-to-all buffer [peer][x][k1 or y within peer][K]
   (no pack / unpack copies around the transposes; rows are merely permuted, which the pointwise row
   pass does not care about).  The pass runs in place. */
#define EXB_SLAB_SEGMENTED 0x100
int exb_slab_pass(exb_plan *plan, void *stream, int32_t pass, int32_t nfields, int32_t stage,
                  const void *in, void *out, const void *U, void *OUT, void *const *S);
/* EXB_SLAB_COL0_INV_PRO restricted to the inverse fields [field0, field0 + nfields): lets the caller
   overlap the all-to-all of field f with the prologue pass of field f+1 (out = base of ALL fields) */
int exb_slab_inv_pro_fields(exb_plan *plan, void *stream, int32_t field0, int32_t nfields, const void *in,
                            void *out);
/* The two passes that feed a transpose, with the transpose fused into their stores: the results go straight
   into the peers' buffers over NVLink (peer_out[r] = base of rank r's destination buffer for ALL fields,
   mapped into this process -- CUDA IPC / symmetric memory), so no collective is needed, only a barrier
   before the consumers read.  Fast kernels only (EXB_EUNSUPPORTED otherwise: use the all-to-all path).
     pass = EXB_SLAB_COL0_INV_PRO:                      in = stage input (layout B);   peers' n_inv-field A buffers
     pass = EXB_SLAB_COL1_FWD[_NL] | EXB_SLAB_SEGMENTED: in = local raw-order A field(s) [field0, +nfields);
                                                         peers' n_fwd-field B buffers
   (in points at field `field0` for the COL1 pass, at the stage input for COL0_INV_PRO.) */
int exb_slab_pass_peer(exb_plan *plan, void *stream, int32_t pass, int32_t field0, int32_t nfields,
                       const void *in, void *const *peer_out);
/* last-axis pitch (complex elements) of the FIELD buffers the slab passes exchange: the inverse fields written by
   COL0_INV_PRO / read by COL1_INV_NL and ROW_NL, the forward fields written by ROW_NL / COL1_FWD_NL and read by
   COL0_FWD_EPI are (nfields, N/P, N, pitch) resp. (nfields, N, N/P, pitch).  pitch = kmax + 1 rounded up to 16 when the
   plan runs the fast kernels with a dealiasing mask (the transposes then ship only the wavenumbers inside the mask),
   N/2 + 1 otherwise.  State buffers (U, OUT, S) and the plain transform passes stay dense (N/2 + 1). */
int exb_plan_field_pitch(const exb_plan *plan);
/* number of single-field inverse / forward transforms per N(u) evaluation of this plan */
int exb_plan_nl_fields(const exb_plan *plan, int32_t *n_inv, int32_t *n_fwd);

/* ---- consumer next to the loop: radially binned spectrum ----
   exponax/_spectral.py:866-1030 `get_spectrum`, applied to u_hat = exb_fft(state).  u_hat: (nfields, N..,
   N/2+1) complex of this plan's grid; out: (nfields, N/2+1) real.  power != 0: power spectrum, else
   amplitude spectrum; average != 0: radial_binning="average" (needs `counts`, a device scratch of
   N/2+1 uint32), else "sum" (`counts` may be NULL). */
int exb_spectrum(exb_plan *plan, void *stream, int64_t nfields, const void *u_hat, void *out,
                 int32_t power, int32_t average, void *counts);

/* ---- producers in front of the loop: random initial conditions ----
   exponax/ic generators are  white noise -> exb_fft -> exb_ic_shape -> exb_ifft -> exb_ic_normalize.
   exb_ic_shape: per-mode real factor, in place on u_hat (nfields, N.., N/2+1) of this plan's grid
     kind 0  RandomTruncatedFourierSeries (ic/_truncated_fourier_series.py:65-100): keep |k_d| <= param on every
             axis, then the DC entry := dc_value (unnormalised coefficient, as the reference sets it)
     kind 1  GaussianRandomField (ic/_gaussian_random_field.py:64-93): |2 pi k / domain_extent|^(-param / 2),
             DC factor 1
   exb_ic_normalize: normalize_ic (ic/_base_ic.py:16-33), in place per field of npoints values: subtract the
     mean, divide by the (population) standard deviation, divide by max |x| -- each if flagged;
     stats: device scratch double[nfields * 4]. */
int exb_ic_shape(exb_plan *plan, void *stream, int64_t nfields, void *u_hat, int32_t kind, double param,
                 double domain_extent, double dc_value);
int exb_ic_normalize(void *stream, int32_t dtype, int64_t nfields, int64_t npoints, void *u,
                     int32_t zero_mean, int32_t std_one, int32_t max_one, double *stats);

/* fused reductions behind exponax.metrics (metrics/_spatial.py:8-196, metrics/_correlation.py:6-60), no plan
   needed.  a, b: (nfields, npoints) real of `dtype` (b may be NULL = zeros); out: device double[nfields*4]:
     out[4f+0] = sum |a-b|^p   out[4f+1] = sum |b|^p   out[4f+2] = sum |a|^p   out[4f+3] = sum a*b
   A field is one channel of one sample; the caller combines the sums (scale (L/N)^D, outer exponent,
   normalised / symmetric ratios, channel sum). */
int exb_metric_sums(void *stream, int32_t dtype, int64_t nfields, int64_t npoints, const void *a,
                    const void *b, double p, double *out);

/* exponax.derivative (exponax/_spectral.py:724-792) between its fft and ifft: out_hat (nfields, D, N.., N/2+1)
   = u_hat (nfields, N.., N/2+1) * (i 2 pi k_d / domain_extent)^order for every axis d; order >= 0 integer;
   out_hat must not alias u_hat. */
int exb_derivative(exb_plan *plan, void *stream, int64_t nfields, const void *u_hat, void *out_hat,
                   int32_t order, double domain_extent);

/* exponax.nonlin_fun.Leray.__call__ (exponax/nonlin_fun/_leray.py:114-136, order 2) on its own: the projection of a
   D-channel spectral field onto its divergence-free part, u_hat / out_hat: (nfields, D, N.., N/2+1) of this plan's
   grid (D = the plan's number of spatial dimensions); out_hat may alias u_hat.  Inside the projected-convection
   nonlinearity the same arithmetic is fused into the epilogue of the forward transform. */
int exb_leray(exb_plan *plan, void *stream, int64_t nfields, const void *u_hat, void *out_hat,
              double domain_extent);

/* Fourier-space aggregation behind exponax.metrics.fourier_* / H1_* (metrics/_fourier.py:15-140), applied to
   x_hat = exb_fft(x), x_hat: (nfields, N.., N/2+1) complex of this plan's grid.  Per field and derivative
   component d:  out[f * ncomp + d] = sum_modes band * (|x_hat| * |2 pi k_d / domain_extent|^order)^p / recon
   with |x_hat| < 1e-5 dropped, band = not(all |k_d| <= low-1) and (all |k_d| <= high) when low >= 0 or
   high >= 0 (a negative bound = not given), recon = the "reconstruction" scaling of the rfft layout.
   derivative_order < 0: no derivative, ncomp = 1; otherwise ncomp = D.  out: device double[nfields * ncomp]. */
int exb_fourier_sums(exb_plan *plan, void *stream, int64_t nfields, const void *x_hat, double p,
                     int32_t low, int32_t high, double derivative_order, double domain_extent, double *out);

/* ---- measurement utilities ("the FP32 peak is not in MEASURED_PEAKS.json -- measure it") ----
   The 1-D persistent kernel (config c2) is bound by FP32 issue and shared-memory bandwidth, not by HBM; these
   two micro-benchmarks give the denominators bench.py reports its fractions against.  Both allocate 256 B of
   scratch, time 5 launches with CUDA events on `stream` and SYNCHRONISE (measurement only, never on the hot path).
     exb_peak_fp32: tflops[0] = scalar FFMA chains, tflops[1] = packed fma.rn.f32x2 (FFMA2) chains   [TFLOP/s]
     exb_peak_smem: gbs[0] = conflict-free 8-byte loads (the kernels' complex<float> accesses), gbs[1] = 16-byte [GB/s] */
int exb_peak_fp32(void *stream, double *tflops);
int exb_peak_smem(void *stream, double *gbs);

/* 1 when the fused entry points (exb_step, exb_step_fourier, exb_rollout, exb_nonlinear_fun) are available for
   this plan; 0 for 1-D grids whose state + ETDRK stage buffers do not fit the shared memory of one SM: those plans
   still serve exb_fft / exb_ifft, and the host evaluates the stage formulas (exponax/etdrk/_etdrk_{0..4}.py) and
   the nonlinear function (exponax/nonlin_fun/*.py) around them with device-array arithmetic. */
int exb_plan_fused_ok(const exb_plan *plan);

/* number of kernel launches issued through this plan so far (bench bookkeeping) */
int64_t exb_launch_count(const exb_plan *plan);

#ifdef __cplusplus
}
#endif
#endif /* EXB_H_ */
