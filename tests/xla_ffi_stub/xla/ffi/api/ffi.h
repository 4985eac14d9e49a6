// API-SHAPED STAND-IN for jaxlib's xla/ffi/api/ffi.h -- TEST INFRASTRUCTURE ONLY.
//
// jaxlib is not installable in this image (no network, no wheel), so jax_ib_b200/csrc/xla_ffi_shim.cc cannot be
// compiled against the real header here.  This file declares, from the published XLA FFI C++ API, exactly the
// names the shim uses (Ffi::Bind().Ctx/Arg/Ret/Attr/Attrs, Buffer, Result, ScratchAllocator, PlatformStream,
// Dictionary, Error / ErrorCode, XLA_FFI_DEFINE_HANDLER_SYMBOL) with the same shapes, so that
// tests/test_cpu_host.py can at least TYPE-CHECK the shim on the CPU box: every handler must be invocable with
// the argument list its binding describes.  Nothing here executes; it is never shipped or linked.
#ifndef JAXIB_TESTS_XLA_FFI_STUB_H_
#define JAXIB_TESTS_XLA_FFI_STUB_H_

#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <string_view>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
template <DataType dt> struct NativeTypeOf;
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<F64> { using type = double; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };
template <> struct NativeTypeOf<U8> { using type = uint8_t; };

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kResourceExhausted, kUnimplemented };

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

template <typename T>
class ErrorOr {
 public:
  ErrorOr(T v) : v_(std::move(v)) {}
  ErrorOr(Error e) : e_(std::move(e)) {}
  bool has_value() const { return v_.has_value(); }
  T& value() { return *v_; }
  T& operator*() { return *v_; }

 private:
  std::optional<T> v_;
  Error e_;
};

template <typename T>
class Span {
 public:
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }

 private:
  const T* p_ = nullptr;
  size_t n_ = 0;
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeTypeOf<dt>::type;
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  Span<int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

class ScratchAllocator {
 public:
  std::optional<void*> Allocate(size_t size, size_t alignment = 1) { (void)size; (void)alignment; return std::nullopt; }
};

template <typename T> struct PlatformStream {};

class Dictionary {
 public:
  template <typename T> ErrorOr<T> get(std::string_view name) const { (void)name; return ErrorOr<T>(T{}); }
  bool contains(std::string_view name) const { (void)name; return false; }
};

namespace internal {
template <typename T> struct CtxArg { using type = T; };
template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };
template <typename... Ts> struct TypeList {};
template <typename Fn, typename List> struct Invocable;
template <typename Fn, typename... Ts>
struct Invocable<Fn, TypeList<Ts...>> : std::is_invocable_r<Error, Fn, Ts...> {};
}  // namespace internal

template <typename... Ts>
class Binding {
 public:
  using Args = internal::TypeList<Ts...>;
  template <typename T> Binding<Ts..., typename internal::CtxArg<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(std::string name) const { (void)name; return {}; }
  template <typename T = Dictionary> Binding<Ts..., T> Attrs() const { return {}; }
};

class Ffi {
 public:
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;

// the real macro defines `extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame*)` that decodes the call frame as the
// binding says and calls impl; the stand-in keeps the symbol and checks that impl accepts exactly those types
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)                                                        \
  static_assert(::xla::ffi::internal::Invocable<decltype(&impl), typename decltype(binding)::Args>::value,      \
                #impl " is not invocable with the arguments its binding declares");                             \
  extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame* call_frame) { (void)call_frame; return nullptr; }

#endif  // JAXIB_TESTS_XLA_FFI_STUB_H_
