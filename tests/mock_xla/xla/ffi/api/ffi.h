// Mock of the slice of XLA's FFI binding API (xla/ffi/api/ffi.h) that integration/updes_jax_ffi.cc uses -- TEST
// INFRASTRUCTURE.  jaxlib and its headers are absent from this image, so the adapter cannot be compiled against the real
// thing; with this header it compiles (the Bind() chain must match the handler's signature or std::apply fails to
// compile -- the class of bug an unbuilt FFI shim would otherwise hide) and each XLA_FFI_DEFINE_HANDLER_SYMBOL becomes an
// extern "C" entry point taking a MockCallFrame, which tests drive from Python (oracle/refshim/jax/ffi.py) the way XLA
// would: context stream, attributes by name, argument and result buffers in order.  Names and call shapes follow the
// public XLA FFI API; nothing here is XLA code.
#ifndef MOCK_XLA_FFI_API_FFI_H_
#define MOCK_XLA_FFI_API_FFI_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

extern "C" {
struct MockBuffer { void *data; int32_t rank; int64_t dims[6]; };
struct MockAttr { const char *name; int32_t is_double; double d; int32_t i; };
struct MockCallFrame {
  void *stream;
  int32_t nargs; MockBuffer *args;
  int32_t nrets; MockBuffer *rets;
  int32_t nattrs; MockAttr *attrs;
  char error[256];
};
}

namespace xla {
namespace ffi {

enum DataType { F64, S32 };
template <DataType> struct NativeTypeOf;
template <> struct NativeTypeOf<F64> { using type = double; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };

class Dimensions {
 public:
  Dimensions(const int64_t *p, size_t n) : p_(p), n_(n) {}
  int64_t operator[](size_t i) const { return p_[i]; }
  size_t size() const { return n_; }
 private:
  const int64_t *p_; size_t n_;
};

template <DataType T> class Buffer {
 public:
  using Native = typename NativeTypeOf<T>::type;
  Buffer() : b_(nullptr) {}
  explicit Buffer(MockBuffer *b) : b_(b) {}
  Native *typed_data() const { return static_cast<Native *>(b_->data); }
  Dimensions dimensions() const { return Dimensions(b_->dims, (size_t)b_->rank); }
 private:
  MockBuffer *b_;
};

template <class B> class Result {
 public:
  explicit Result(B b) : b_(b) {}
  B *operator->() { return &b_; }
  B &operator*() { return b_; }
 private:
  B b_;
};
template <DataType T> using ResultBuffer = Result<Buffer<T>>;

class Error {
 public:
  static Error Success() { return Error(false, ""); }
  static Error Internal(std::string m) { return Error(true, std::move(m)); }
  bool failure() const { return fail_; }
  const std::string &message() const { return msg_; }
 private:
  Error(bool f, std::string m) : fail_(f), msg_(std::move(m)) {}
  bool fail_; std::string msg_;
};

template <class T> struct PlatformStream {};

namespace mock {
struct Cursor { MockCallFrame *f; const std::vector<std::string> *names; size_t arg = 0, ret = 0, attr = 0; bool bad = false; };
template <class T> struct CtxStage;
template <class T> struct CtxStage<PlatformStream<T>> {
  using type = T;
  static type decode(Cursor &c) { return reinterpret_cast<T>(c.f->stream); }
};
template <class T> struct AttrStage {
  using type = T;
  static type decode(Cursor &c) {
    const std::string &want = (*c.names)[c.attr++];
    for (int k = 0; k < c.f->nattrs; k++)
      if (want == c.f->attrs[k].name) return c.f->attrs[k].is_double ? (T)c.f->attrs[k].d : (T)c.f->attrs[k].i;
    c.bad = true;
    return T();
  }
};
template <class B> struct ArgStage {
  using type = B;
  static type decode(Cursor &c) { if ((int)c.arg >= c.f->nargs) { c.bad = true; return B(); } return B(&c.f->args[c.arg++]); }
};
template <class B> struct RetStage {
  using type = Result<B>;
  static type decode(Cursor &c) { if ((int)c.ret >= c.f->nrets) { c.bad = true; return type(B()); } return type(B(&c.f->rets[c.ret++])); }
};
}  // namespace mock

template <class... S> class Binding {
 public:
  std::vector<std::string> names;
  template <class T> Binding<S..., mock::CtxStage<T>> Ctx() { return next<mock::CtxStage<T>>(); }
  template <class T> Binding<S..., mock::AttrStage<T>> Attr(const char *name) { auto b = next<mock::AttrStage<T>>(); b.names.push_back(name); return b; }
  template <class B> Binding<S..., mock::ArgStage<B>> Arg() { return next<mock::ArgStage<B>>(); }
  template <class B> Binding<S..., mock::RetStage<B>> Ret() { return next<mock::RetStage<B>>(); }

  template <class Fn> int Call(Fn fn, MockCallFrame *f) const {
    mock::Cursor c{f, &names};
    std::tuple<typename S::type...> t{S::decode(c)...};        // braced init: decoded left to right
    if (c.bad || (int)c.arg != f->nargs || (int)c.ret != f->nrets) {
      std::strncpy(f->error, "call frame does not match the handler's binding (arity or attribute names)", sizeof(f->error) - 1);
      return 2;
    }
    Error e = std::apply(fn, std::move(t));
    if (e.failure()) { std::strncpy(f->error, e.message().c_str(), sizeof(f->error) - 1); return 1; }
    return 0;
  }
 private:
  template <class N> Binding<S..., N> next() const { Binding<S..., N> b; b.names = names; return b; }
};

struct Ffi { static Binding<> Bind() { return Binding<>(); } };

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                         \
  extern "C" int name(MockCallFrame *frame) {                                      \
    static const auto b = (binding);                                               \
    return b.Call(impl, frame);                                                    \
  }

#endif  // MOCK_XLA_FFI_API_FFI_H_
