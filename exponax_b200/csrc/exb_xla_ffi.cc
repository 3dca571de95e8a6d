// XLA-FFI adapter: thin jax.ffi handlers over the C ABI of libexb (include/exb.h).
//
// NOT compiled in this repository's default build: jaxlib's headers (xla/ffi/api/ffi.h) are not
// installable in the build image.  Build where JAX exists.
// The handlers only translate buffers/attributes and enqueue on the stream XLA provides; all
// arithmetic lives behind exb_* and is what the parity tests exercise through ctypes.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define EXB_HAVE_XLA_FFI 1
#endif
#endif

#ifdef EXB_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/exb.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error to_error(int rc) {
  if (rc == EXB_OK) return ffi::Error::Success();
  return ffi::Error(rc == EXB_EINVAL ? ffi::ErrorCode::kInvalidArgument
                                     : (rc == EXB_EUNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                               : ffi::ErrorCode::kInternal),
                    exb_last_error());
}

// leading (vmapped) dimensions of `buf` beyond the per-trajectory rank `state_rank`
static int64_t batch_of(const ffi::AnyBuffer& buf, size_t state_rank) {
  auto dims = buf.dimensions();
  int64_t b = 1;
  for (size_t i = 0; i + state_rank < dims.size(); ++i) b *= dims[i];
  return b;
}

// u0 (batch.., C, N..) -> trajectory (batch.., T, C, N..); attrs: plan handle, n_saved, substeps, flags, rank
static ffi::Error RolloutImpl(cudaStream_t stream, ffi::AnyBuffer u0, ffi::Result<ffi::AnyBuffer> out,
                              ffi::Result<ffi::AnyBuffer> workspace, uint64_t plan, int64_t n_saved,
                              int32_t substeps, int32_t flags, int32_t state_rank) {
  int64_t batch = batch_of(u0, (size_t)state_rank);
  return to_error(exb_rollout(reinterpret_cast<exb_plan*>(plan), stream, batch, n_saved, substeps,
                              (uint32_t)flags, u0.untyped_data(), out->untyped_data(),
                              workspace->untyped_data()));
}

static ffi::Error StepFourierImpl(cudaStream_t stream, ffi::AnyBuffer u_hat, ffi::Result<ffi::AnyBuffer> out,
                                  ffi::Result<ffi::AnyBuffer> workspace, uint64_t plan, int32_t state_rank) {
  int64_t batch = batch_of(u_hat, (size_t)state_rank);
  return to_error(exb_step_fourier(reinterpret_cast<exb_plan*>(plan), stream, batch, u_hat.untyped_data(),
                                   out->untyped_data(), workspace->untyped_data()));
}

static ffi::Error FftImpl(cudaStream_t stream, ffi::AnyBuffer u, ffi::Result<ffi::AnyBuffer> out,
                          ffi::Result<ffi::AnyBuffer> workspace, uint64_t plan, int32_t num_spatial_dims) {
  int64_t fields = batch_of(u, (size_t)num_spatial_dims);
  return to_error(exb_fft(reinterpret_cast<exb_plan*>(plan), stream, fields, 1, u.untyped_data(),
                          out->untyped_data(), workspace->untyped_data()));
}

static ffi::Error IfftImpl(cudaStream_t stream, ffi::AnyBuffer u_hat, ffi::Result<ffi::AnyBuffer> out,
                           ffi::Result<ffi::AnyBuffer> workspace, uint64_t plan, int32_t num_spatial_dims) {
  int64_t fields = batch_of(u_hat, (size_t)num_spatial_dims);
  return to_error(exb_ifft(reinterpret_cast<exb_plan*>(plan), stream, fields, 1, u_hat.untyped_data(),
                           out->untyped_data(), workspace->untyped_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(exb_xla_rollout, RolloutImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<uint64_t>("plan")
                                  .Attr<int64_t>("n_saved")
                                  .Attr<int32_t>("substeps")
                                  .Attr<int32_t>("flags")
                                  .Attr<int32_t>("state_rank"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(exb_xla_step_fourier, StepFourierImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<uint64_t>("plan")
                                  .Attr<int32_t>("state_rank"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(exb_xla_fft, FftImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<uint64_t>("plan")
                                  .Attr<int32_t>("num_spatial_dims"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(exb_xla_ifft, IfftImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<uint64_t>("plan")
                                  .Attr<int32_t>("num_spatial_dims"));
#endif  // EXB_HAVE_XLA_FFI
