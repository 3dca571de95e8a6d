// Minimal MOCK of jaxlib's xla/ffi/api/ffi.h: declares exactly the symbols exponax_b200/csrc/exb_xla_ffi.cc
// uses, with the signatures of the real header (XLA FFI, jax >= 0.4.38), so that the adapter is syntax- and
// type-checked in an image without jaxlib (compile-only: `g++ -fsyntax-only`).  NOT a functional FFI runtime.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla {
namespace ffi {

template <typename T> class Span {
 public:
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
  const T* begin() const { return d_; }
  const T* end() const { return d_ + n_; }

 private:
  const T* d_;
  size_t n_;
};

enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kNotFound = 5, kUnimplemented = 12, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

enum class DataType { F32, F64, C64, C128, U8 };

class AnyBuffer {
 public:
  using Dimensions = Span<int64_t>;
  DataType element_type() const { return DataType::F32; }
  Dimensions dimensions() const { return Dimensions(nullptr, 0); }
  void* untyped_data() const { return nullptr; }
  size_t size_bytes() const { return 0; }
};

template <typename T> class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};

template <typename T> struct PlatformStream {};
struct DeviceOrdinal {};

// what the handler implementation receives for a bound slot (real header: CtxDecoding / ArgDecoding / ...)
template <typename T> struct Decoded { using type = T; };
template <typename T> struct Decoded<PlatformStream<T>> { using type = T; };

template <typename... Ts> class Binding {
 public:
  template <typename T> Binding<Ts..., T> Ctx() && { return {}; }
  template <typename T> Binding<Ts..., T> Arg() && { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() && { return {}; }
  template <typename T> Binding<Ts..., T> Attr(std::string) && { return {}; }
  // the mock's whole point: the implementation must be callable with exactly the bound slots
  template <typename Fn> int To(Fn) && {
    static_assert(std::is_invocable_r_v<Error, Fn, typename Decoded<Ts>::type...>,
                  "handler implementation does not match its XLA FFI binding");
    return 0;
  }
};

class Ffi {
 public:
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines `extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*)` dispatching to `impl` through `binding`
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                    \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame* call_frame) {             \
    static int bound = (binding).To(impl);                                    \
    (void)bound;                                                              \
    (void)call_frame;                                                         \
    return nullptr;                                                           \
  }
