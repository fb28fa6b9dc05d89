// Minimal stand-in for jaxlib's xla/ffi/api/ffi.h -- TEST INFRASTRUCTURE, not the real header.
//
// jax / jaxlib are not installable in this image, so csrc/xla_ffi.cc can neither be compiled against the real
// header nor be called by XLA here.  This file restates just the surface that xla_ffi.cc uses -- Buffer<dtype>, Result<>,
// Error / ErrorCode, PlatformStream<>, Ffi::Bind().Ctx/Arg/Attr/Ret and XLA_FFI_DEFINE_HANDLER_SYMBOL -- with the same
// spelling and argument-decoding order as the documented API (operands in order, then attributes by name, then results),
// so that type errors in the handlers surface at build time and the handlers can be driven by a test through a plain
// call frame.  It says nothing about the real ABI of XLA_FFI_CallFrame.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <tuple>
#include <utility>
#include <variant>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { U8 = 1, F32 = 2 };
template <DataType dt> struct NativeType;
template <> struct NativeType<U8> { using type = uint8_t; };
template <> struct NativeType<F32> { using type = float; };

class Span {
 public:
  Span(const int64_t* p, size_t n) : p_(p), n_(n) {}
  size_t size() const { return n_; }
  int64_t operator[](size_t i) const { return p_[i]; }
  bool operator==(const Span& o) const {
    if (n_ != o.n_) return false;
    for (size_t i = 0; i < n_; ++i)
      if (p_[i] != o.p_[i]) return false;
    return true;
  }
  bool operator!=(const Span& o) const { return !(*this == o); }

 private:
  const int64_t* p_;
  size_t n_;
};

struct AnyBuffer {
  void* data = nullptr;
  DataType dtype = U8;
  std::vector<int64_t> dims;
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : dims) n *= (size_t)d;
    return n;
  }
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeType<dt>::type;
  Buffer() = default;
  explicit Buffer(AnyBuffer* b) : b_(b) {}
  void* untyped_data() const { return b_->data; }
  T* typed_data() const { return static_cast<T*>(b_->data); }
  Span dimensions() const { return Span(b_->dims.data(), b_->dims.size()); }
  size_t element_count() const { return b_->element_count(); }
  size_t size_bytes() const { return b_->element_count() * sizeof(T); }

 private:
  AnyBuffer* b_ = nullptr;
};

template <typename T>
class Result {
 public:
  explicit Result(T v) : v_(v) {}
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};

enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  ErrorCode code() const { return code_; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T> struct PlatformStream {};

struct CallFrame {
  void* stream = nullptr;
  std::vector<AnyBuffer> args, rets;
  std::map<std::string, std::variant<int64_t, float>> attrs;
};

namespace internal {
template <typename T> struct CtxTag {};
template <typename T> struct ArgTag {};
template <typename T> struct AttrTag {};
template <typename T> struct RetTag {};

struct DecodeState {
  CallFrame* frame;
  const std::vector<std::string>* names;
  size_t arg = 0, ret = 0, attr = 0;
  Error error;
};

template <typename Tag> struct Decode;
template <typename S> struct Decode<CtxTag<PlatformStream<S>>> {
  using type = S;
  static type get(DecodeState& st) { return reinterpret_cast<S>(st.frame->stream); }
};
template <DataType dt> struct Decode<ArgTag<Buffer<dt>>> {
  using type = Buffer<dt>;
  static type get(DecodeState& st) {
    static AnyBuffer none;
    if (st.arg >= st.frame->args.size()) { st.error = Error(ErrorCode::kInvalidArgument, "too few operands"); ++st.arg; return type(&none); }
    AnyBuffer* b = &st.frame->args[st.arg++];
    if (b->dtype != dt) st.error = Error(ErrorCode::kInvalidArgument, "operand dtype mismatch");
    return type(b);
  }
};
template <DataType dt> struct Decode<RetTag<Buffer<dt>>> {
  using type = Result<Buffer<dt>>;
  static type get(DecodeState& st) {
    static AnyBuffer none;
    if (st.ret >= st.frame->rets.size()) { st.error = Error(ErrorCode::kInvalidArgument, "too few results"); ++st.ret; return type(Buffer<dt>(&none)); }
    AnyBuffer* b = &st.frame->rets[st.ret++];
    if (b->dtype != dt) st.error = Error(ErrorCode::kInvalidArgument, "result dtype mismatch");
    return type(Buffer<dt>(b));
  }
};
template <typename T> struct Decode<AttrTag<T>> {
  using type = T;
  static type get(DecodeState& st) {
    const std::string& name = (*st.names)[st.attr++];
    auto it = st.frame->attrs.find(name);
    if (it == st.frame->attrs.end() || !std::holds_alternative<T>(it->second)) {
      st.error = Error(ErrorCode::kInvalidArgument, "missing or mistyped attribute " + name);
      return T();
    }
    return std::get<T>(it->second);
  }
};

template <typename Fn, typename... Tags>
class Handler {
 public:
  Handler(Fn fn, std::vector<std::string> names) : fn_(fn), names_(std::move(names)) {}
  Error* Call(CallFrame* frame) const {
    DecodeState st{frame, &names_};
    // braced initialisation evaluates left to right: operands, attributes and results are consumed in binding order
    std::tuple<typename Decode<Tags>::type...> vals{Decode<Tags>::get(st)...};
    Error e = st.error.failure() ? st.error : std::apply(fn_, std::move(vals));
    return e.success() ? nullptr : new Error(e);
  }

 private:
  Fn fn_;
  std::vector<std::string> names_;
};
}  // namespace internal

template <typename... Tags>
class Binding {
 public:
  Binding() = default;
  explicit Binding(std::vector<std::string> names) : names_(std::move(names)) {}
  template <typename T> Binding<Tags..., internal::CtxTag<T>> Ctx() const { return Binding<Tags..., internal::CtxTag<T>>(names_); }
  template <typename T> Binding<Tags..., internal::ArgTag<T>> Arg() const { return Binding<Tags..., internal::ArgTag<T>>(names_); }
  template <typename T> Binding<Tags..., internal::RetTag<T>> Ret() const { return Binding<Tags..., internal::RetTag<T>>(names_); }
  template <typename T> Binding<Tags..., internal::AttrTag<T>> Attr(std::string name) const {
    std::vector<std::string> n = names_;
    n.push_back(std::move(name));
    return Binding<Tags..., internal::AttrTag<T>>(std::move(n));
  }
  template <typename Fn> internal::Handler<Fn, Tags...> To(Fn fn) const { return internal::Handler<Fn, Tags...>(fn, names_); }

 private:
  std::vector<std::string> names_;
};

struct Ffi {
  static Binding<> Bind() { return Binding<>(); }
};

}  // namespace ffi
}  // namespace xla

using XLA_FFI_CallFrame = xla::ffi::CallFrame;
using XLA_FFI_Error = xla::ffi::Error;
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)        \
  extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame* call_frame) { \
    static const auto handler = (binding).To(impl);             \
    return handler.Call(call_frame);                            \
  }
