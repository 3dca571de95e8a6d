// XLA-FFI adapter: thin jax.ffi handlers over the C ABI of libexb (include/exb.h).
//
// jaxlib's headers (xla/ffi/api/ffi.h) are not installable in the build image: the default build
// only SYNTAX-CHECKS this file against tests/mock_xla (a mock of the header declaring the symbols used here;
// `__graft_entry__.build()` and tests/test_host_logic.py::test_xla_ffi_adapter_compiles).  Build it for real where
// JAX exists.  The handlers only translate buffers / attributes and enqueue on the stream XLA
// provides; all arithmetic lives behind exb_* and is what the parity tests exercise through ctypes.
//
// Plans and devices: an `exb_plan` owns device-resident tables, so one plan exists PER DEVICE.  The traced
// function carries a device-independent `plan_id` attribute; the host shim registers the plan it created on each
// device with exb_xla_register_plan(plan_id, device_ordinal, plan), and a handler resolves (plan_id, the device it
// is running on) -- correct under multi-device `shard_map` / `pmap`, where one executable runs on every device.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define EXB_HAVE_XLA_FFI 1
#endif
#endif

#ifdef EXB_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <utility>

#include "../../include/exb.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

std::mutex g_mu;                                              // handlers run concurrently, one host thread per device
std::map<std::pair<int64_t, int>, exb_plan*> g_plans;         // (plan_id, device ordinal) -> plan

ffi::Error to_error(int rc) {
  if (rc == EXB_OK) return ffi::Error::Success();
  return ffi::Error(rc == EXB_EINVAL ? ffi::ErrorCode::kInvalidArgument
                                     : (rc == EXB_EUNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                               : ffi::ErrorCode::kInternal),
                    exb_last_error());
}

// the plan registered for `plan_id` on the device this handler runs on
ffi::Error resolve(int64_t plan_id, exb_plan** plan) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ffi::Error(ffi::ErrorCode::kInternal, "cudaGetDevice failed");
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_plans.find({plan_id, dev});
  if (it == g_plans.end())
    return ffi::Error(ffi::ErrorCode::kNotFound, "exb: no plan registered for this plan_id on this device");
  *plan = it->second;
  return ffi::Error::Success();
}

// product of the leading (vmapped) dimensions of `buf` beyond its trailing `rank` dimensions
int64_t leading(const ffi::AnyBuffer& buf, int32_t rank) {
  auto dims = buf.dimensions();
  int64_t b = 1;
  for (size_t i = 0; i + (size_t)rank < dims.size(); ++i) b *= dims[i];
  return b;
}

// u0 (batch.., C, N..) -> (batch.., T, C, N..) [default] / (T, batch.., C, N..) [LAYOUT_TB] / (batch.., C, N..)
// [FINAL_ONLY]; `flags` = EXB_ROLLOUT_* bits (include_init, layout, final only, spectral carry), all forwarded
ffi::Error RolloutImpl(cudaStream_t stream, ffi::AnyBuffer u0, ffi::Result<ffi::AnyBuffer> out,
                       ffi::Result<ffi::AnyBuffer> workspace, int64_t plan_id, int64_t n_saved, int32_t substeps,
                       int32_t flags, int32_t state_rank) {
  exb_plan* plan = nullptr;
  if (ffi::Error e = resolve(plan_id, &plan); !e.success()) return e;
  return to_error(exb_rollout(plan, stream, leading(u0, state_rank), n_saved, substeps, (uint32_t)flags,
                              u0.untyped_data(), out->untyped_data(), workspace->untyped_data()));
}

// physical -> physical, BaseStepper.step (exponax/_base_stepper.py:201-220)
ffi::Error StepImpl(cudaStream_t stream, ffi::AnyBuffer u, ffi::Result<ffi::AnyBuffer> out,
                    ffi::Result<ffi::AnyBuffer> workspace, int64_t plan_id, int32_t state_rank) {
  exb_plan* plan = nullptr;
  if (ffi::Error e = resolve(plan_id, &plan); !e.success()) return e;
  return to_error(exb_step(plan, stream, leading(u, state_rank), u.untyped_data(), out->untyped_data(),
                           workspace->untyped_data()));
}

// spectral -> spectral, BaseStepper.step_fourier (exponax/_base_stepper.py:222-239)
ffi::Error StepFourierImpl(cudaStream_t stream, ffi::AnyBuffer u_hat, ffi::Result<ffi::AnyBuffer> out,
                           ffi::Result<ffi::AnyBuffer> workspace, int64_t plan_id, int32_t state_rank) {
  exb_plan* plan = nullptr;
  if (ffi::Error e = resolve(plan_id, &plan); !e.success()) return e;
  return to_error(exb_step_fourier(plan, stream, leading(u_hat, state_rank), u_hat.untyped_data(),
                                   out->untyped_data(), workspace->untyped_data()));
}

// N(u_hat) of a built-in nonlinear function (exponax/nonlin_fun/*.py __call__)
ffi::Error NonlinearFunImpl(cudaStream_t stream, ffi::AnyBuffer u_hat, ffi::Result<ffi::AnyBuffer> out,
                            ffi::Result<ffi::AnyBuffer> workspace, int64_t plan_id, int32_t state_rank) {
  exb_plan* plan = nullptr;
  if (ffi::Error e = resolve(plan_id, &plan); !e.success()) return e;
  return to_error(exb_nonlinear_fun(plan, stream, leading(u_hat, state_rank), u_hat.untyped_data(),
                                    out->untyped_data(), workspace->untyped_data()));
}

// exponax.fft / exponax.ifft (exponax/_spectral.py:614-721): every leading dimension is a field
ffi::Error FftImpl(cudaStream_t stream, ffi::AnyBuffer u, ffi::Result<ffi::AnyBuffer> out,
                   ffi::Result<ffi::AnyBuffer> workspace, int64_t plan_id, int32_t num_spatial_dims) {
  exb_plan* plan = nullptr;
  if (ffi::Error e = resolve(plan_id, &plan); !e.success()) return e;
  return to_error(exb_fft(plan, stream, leading(u, num_spatial_dims), 1, u.untyped_data(), out->untyped_data(),
                          workspace->untyped_data()));
}

ffi::Error IfftImpl(cudaStream_t stream, ffi::AnyBuffer u_hat, ffi::Result<ffi::AnyBuffer> out,
                    ffi::Result<ffi::AnyBuffer> workspace, int64_t plan_id, int32_t num_spatial_dims) {
  exb_plan* plan = nullptr;
  if (ffi::Error e = resolve(plan_id, &plan); !e.success()) return e;
  return to_error(exb_ifft(plan, stream, leading(u_hat, num_spatial_dims), 1, u_hat.untyped_data(),
                           out->untyped_data(), workspace->untyped_data()));
}

}  // namespace

extern "C" {
// Host shim (Python, via ctypes), outside any traced function: make `plan` (created with exb_plan_create while
// `device_ordinal` was current) the target of `plan_id` on that device; plan == NULL removes the entry.
int exb_xla_register_plan(int64_t plan_id, int32_t device_ordinal, exb_plan* plan) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (plan) g_plans[{plan_id, (int)device_ordinal}] = plan;
  else g_plans.erase({plan_id, (int)device_ordinal});
  return EXB_OK;
}
}

#define EXB_STATE_HANDLER(symbol, impl)                                 \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl,                           \
                                ffi::Ffi::Bind()                        \
                                    .Ctx<ffi::PlatformStream<cudaStream_t>>() \
                                    .Arg<ffi::AnyBuffer>()              \
                                    .Ret<ffi::AnyBuffer>()              \
                                    .Ret<ffi::AnyBuffer>()              \
                                    .Attr<int64_t>("plan_id")           \
                                    .Attr<int32_t>("state_rank"))

XLA_FFI_DEFINE_HANDLER_SYMBOL(exb_xla_rollout, RolloutImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("plan_id")
                                  .Attr<int64_t>("n_saved")
                                  .Attr<int32_t>("substeps")
                                  .Attr<int32_t>("flags")
                                  .Attr<int32_t>("state_rank"));
EXB_STATE_HANDLER(exb_xla_step, StepImpl);
EXB_STATE_HANDLER(exb_xla_step_fourier, StepFourierImpl);
EXB_STATE_HANDLER(exb_xla_nonlinear_fun, NonlinearFunImpl);

XLA_FFI_DEFINE_HANDLER_SYMBOL(exb_xla_fft, FftImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("plan_id")
                                  .Attr<int32_t>("num_spatial_dims"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(exb_xla_ifft, IfftImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("plan_id")
                                  .Attr<int32_t>("num_spatial_dims"));
#endif  // EXB_HAVE_XLA_FFI
