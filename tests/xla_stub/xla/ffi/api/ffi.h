// TEST-ONLY stand-in for the part of XLA's header-only FFI API (xla/ffi/api/ffi.h) that unidom_b200/csrc/xla_ffi.cc
// uses.  Neither jax nor the XLA headers exist in this image, so the adapter could never be compiled; this stub lets
// the build compile it and lets tests/test_xla_ffi.py CALL the handlers through a plain C entry point
// (ud_stub_call below) to check what the adapter itself is responsible for: buffer order, attribute decoding,
// workspace plumbing and error mapping.  It is written from the documented shape of the API (Ffi::Bind() builder,
// RemainingArgs/RemainingRets, AnyBuffer, Error, XLA_FFI_DEFINE_HANDLER_SYMBOL), not from XLA's sources, and says
// nothing about ABI compatibility with a real XLA -- that still needs a machine with jax.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  ErrorCode code() const { return code_; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

struct Dims {
  const int64_t* p = nullptr;
  size_t n = 0;
  int64_t operator[](size_t i) const { return p[i]; }
  size_t size() const { return n; }
};

class AnyBuffer {
 public:
  AnyBuffer() = default;
  AnyBuffer(void* data, const int64_t* dims, size_t rank, size_t elem_bytes)
      : data_(data), dims_{dims, rank}, elem_bytes_(elem_bytes) {}
  void* untyped_data() const { return data_; }
  Dims dimensions() const { return dims_; }
  size_t element_count() const {
    size_t c = 1;
    for (size_t i = 0; i < dims_.n; ++i) c *= (size_t)dims_.p[i];
    return c;
  }
  size_t size_bytes() const { return element_count() * elem_bytes_; }

 private:
  void* data_ = nullptr;
  Dims dims_;
  size_t elem_bytes_ = 4;
};

template <class T>
class Result {
 public:
  explicit Result(T v) : v_(v) {}
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};

template <class T>
class ErrorOr {
 public:
  explicit ErrorOr(T v) : v_(v), ok_(true) {}
  ErrorOr() : ok_(false) {}
  bool has_value() const { return ok_; }
  T& value() { return v_; }

 private:
  T v_{};
  bool ok_;
};

class RemainingArgs {
 public:
  explicit RemainingArgs(const std::vector<AnyBuffer>* b) : b_(b) {}
  size_t size() const { return b_->size(); }
  template <class T>
  ErrorOr<T> get(size_t i) const {
    return i < b_->size() ? ErrorOr<T>((*b_)[i]) : ErrorOr<T>();
  }

 private:
  const std::vector<AnyBuffer>* b_;
};

class RemainingRets {
 public:
  explicit RemainingRets(const std::vector<AnyBuffer>* b) : b_(b) {}
  size_t size() const { return b_->size(); }
  template <class T>
  ErrorOr<Result<T>> get(size_t i) const {
    return i < b_->size() ? ErrorOr<Result<T>>(Result<T>((*b_)[i])) : ErrorOr<Result<T>>(Result<T>(T()));
  }

 private:
  const std::vector<AnyBuffer>* b_;
};

template <class S>
struct PlatformStream {
  using stream_type = S;
};

// what a call carries in this stub
struct StubAttr {
  std::string name;
  bool is_double;
  int64_t i;
  double d;
};
struct StubFrame {
  void* stream;
  std::vector<AnyBuffer> args, rets;
  std::vector<StubAttr> attrs;
  std::string error;   // filled by the handler symbol
};

template <class T>
inline bool stub_attr(const StubFrame& f, const std::string& name, T* out) {
  for (const StubAttr& a : f.attrs)
    if (a.name == name) {
      *out = a.is_double ? (T)a.d : (T)a.i;
      return true;
    }
  return false;
}

// Ffi::Bind().Ctx<PlatformStream<S>>().RemainingArgs().RemainingRets().Attr<T>(name)...: the builder records the
// attribute names; To(fn) decodes them in order and calls fn(stream, RemainingArgs, RemainingRets, attrs...).
template <class S, class... As>
class Binding {
 public:
  Binding() = default;
  explicit Binding(std::vector<std::string> names) : names_(std::move(names)) {}
  Binding<S, As...> RemainingArgs() const { return *this; }
  Binding<S, As...> RemainingRets() const { return *this; }
  template <class T>
  Binding<S, As..., T> Attr(const char* name) const {
    std::vector<std::string> n = names_;
    n.emplace_back(name);
    return Binding<S, As..., T>(n);
  }
  template <class C>
  auto Ctx() const {
    return Binding<typename C::stream_type, As...>(names_);
  }
  const std::vector<std::string>& names() const { return names_; }

  template <class Fn>
  Error Call(Fn fn, StubFrame& f) const {
    return CallImpl(fn, f, std::index_sequence_for<As...>{});
  }

 private:
  template <class Fn, size_t... I>
  Error CallImpl(Fn fn, StubFrame& f, std::index_sequence<I...>) const {
    std::tuple<As...> vals;
    bool ok = true;
    (void)std::initializer_list<int>{(ok = ok && stub_attr(f, names_[I], &std::get<I>(vals)), 0)...};
    if (!ok) return Error(ErrorCode::kInvalidArgument, "missing attribute");
    return fn((S)f.stream, ffi::RemainingArgs(&f.args), ffi::RemainingRets(&f.rets), std::get<I>(vals)...);
  }
  std::vector<std::string> names_;
};
struct Ffi {
  static Binding<void*> Bind() { return Binding<void*>(); }
};
}  // namespace ffi
}  // namespace xla

#define XLA_FFI_STUB 1
// The handler symbol: extern "C" int symbol(StubFrame*): 0 on success, the ErrorCode otherwise (message in frame.error)
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, fn, binding)                          \
  extern "C" int symbol(::xla::ffi::StubFrame* frame) {                             \
    static const auto b = (binding);                                                \
    ::xla::ffi::Error e = b.Call(fn, *frame);                                       \
    frame->error = e.message();                                                     \
    return (int)e.code();                                                           \
  }
