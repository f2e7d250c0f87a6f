// Minimal stand-in for jaxlib's xla/ffi/api/ffi.h — TEST INFRASTRUCTURE, not the real header.
// It implements just the surface the generated csrc/ffi/dogstep_ffi.cc uses (Ffi::Bind().Ctx/Arg/Ret/Attr, AnyBuffer,
// Result, Error, PlatformStream, XLA_FFI_DEFINE_HANDLER_SYMBOL) with the same shapes, so that tests/test_ffi.py can compile
// the shim on a box without jaxlib, statically check every handler's signature against its binding, and call handlers
// through a fake call frame.  Semantics follow the XLA FFI documentation: Bind order = parameter order; Ret<T> arrives as
// Result<T> (pointer-like); Attr<T>(name) arrives by value.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <variant>
#include <vector>

namespace xla::ffi {

template <typename T>
struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

class AnyBuffer {
 public:
  AnyBuffer() = default;
  AnyBuffer(void* data, size_t bytes, size_t count, std::vector<int64_t> dims = {})
      : data_(data), bytes_(bytes), count_(count), dims_(std::move(dims)) {}
  void* untyped_data() const { return data_; }
  size_t size_bytes() const { return bytes_; }
  size_t element_count() const { return count_; }
  Span<int64_t> dimensions() const { return Span<int64_t>{dims_.data(), dims_.size()}; }

 private:
  void* data_ = nullptr;
  size_t bytes_ = 0, count_ = 0;
  std::vector<int64_t> dims_;
};

template <typename T>
class Result {
 public:
  explicit Result(T v) : v_(std::move(v)) {}
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};

class Error {
 public:
  static Error Success() { return Error(true, ""); }
  static Error InvalidArgument(std::string m) { return Error(false, "INVALID_ARGUMENT: " + m); }
  static Error Internal(std::string m) { return Error(false, "INTERNAL: " + m); }
  bool success() const { return ok_; }
  const std::string& message() const { return msg_; }

 private:
  Error(bool ok, std::string m) : ok_(ok), msg_(std::move(m)) {}
  bool ok_;
  std::string msg_;
};

template <typename T>
struct PlatformStream {};

using AttrValue = std::variant<int32_t, int64_t, uint32_t, float>;
struct CallFrame {
  void* stream = nullptr;
  std::vector<AnyBuffer> args, rets;
  std::map<std::string, AttrValue> attrs;
};

namespace stub {
template <typename T> struct CtxTag {};
template <typename T> struct ArgTag {};
template <typename T> struct RetTag {};
template <typename T> struct AttrTag {};
template <typename Tag> struct ParamOf;
template <typename S> struct ParamOf<CtxTag<PlatformStream<S>>> { using type = S; };
template <typename T> struct ParamOf<ArgTag<T>> { using type = T; };
template <typename T> struct ParamOf<RetTag<T>> { using type = Result<T>; };
template <typename T> struct ParamOf<AttrTag<T>> { using type = T; };

struct Cursor { size_t a = 0, r = 0, k = 0; };
template <typename S>
S decode(CtxTag<PlatformStream<S>>, CallFrame& f, Cursor&, const std::vector<std::string>&) { return reinterpret_cast<S>(f.stream); }
template <typename T>
T decode(ArgTag<T>, CallFrame& f, Cursor& c, const std::vector<std::string>&) { return f.args.at(c.a++); }
template <typename T>
Result<T> decode(RetTag<T>, CallFrame& f, Cursor& c, const std::vector<std::string>&) { return Result<T>(f.rets.at(c.r++)); }
template <typename T>
T decode(AttrTag<T>, CallFrame& f, Cursor& c, const std::vector<std::string>& names) { return std::get<T>(f.attrs.at(names.at(c.k++))); }
}  // namespace stub

template <typename... Tags>
struct Binding {
  std::vector<std::string> attr_names;
  template <typename T> Binding<Tags..., stub::CtxTag<T>> Ctx() && { return {std::move(attr_names)}; }
  template <typename T> Binding<Tags..., stub::ArgTag<T>> Arg() && { return {std::move(attr_names)}; }
  template <typename T> Binding<Tags..., stub::RetTag<T>> Ret() && { return {std::move(attr_names)}; }
  template <typename T> Binding<Tags..., stub::AttrTag<T>> Attr(std::string name) && {
    attr_names.push_back(std::move(name));
    return {std::move(attr_names)};
  }
  // number of operands / results / attributes this binding declares
  static constexpr size_t num_args = (0 + ... + (std::is_same_v<Tags, stub::ArgTag<AnyBuffer>> ? 1 : 0));
  static constexpr size_t num_rets = (0 + ... + (std::is_same_v<Tags, stub::RetTag<AnyBuffer>> ? 1 : 0));
  template <typename Fn>
  static constexpr bool matches = std::is_invocable_r_v<Error, Fn, typename stub::ParamOf<Tags>::type...>;
  template <typename Fn>
  Error Call(Fn fn, CallFrame& f) const {
    stub::Cursor c;
    std::tuple<typename stub::ParamOf<Tags>::type...> params{stub::decode(Tags{}, f, c, attr_names)...};  // left to right
    return std::apply(fn, std::move(params));
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

struct StubHandler {
  std::function<Error(CallFrame&)> call;
  std::vector<std::string> attr_names;
  size_t num_args, num_rets;
};

template <typename Fn, typename B>
StubHandler MakeStubHandler(Fn fn, B binding) {
  static_assert(B::template matches<Fn>, "handler signature does not match its binding (order: Ctx, Arg..., Ret..., Attr...)");
  auto names = binding.attr_names;
  return StubHandler{[fn, binding](CallFrame& f) { return binding.Call(fn, f); }, names, B::num_args, B::num_rets};
}

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding) ::xla::ffi::StubHandler sym = ::xla::ffi::MakeStubHandler(&impl, binding)
