// TEST INFRASTRUCTURE: drives the XLA-FFI handlers of csrc/xla_ffi.cc through the stand-in call frame of
// tests/ffi_stub/xla/ffi/api/ffi.h (jax is absent from this image, so XLA cannot call them here).
#include <cstdio>
#include <cstring>

#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using HandlerFn = XLA_FFI_Error* (*)(XLA_FFI_CallFrame*);

static int finish(XLA_FFI_Error* e, char* err, int errlen) {
  if (!e) {
    if (err && errlen > 0) err[0] = 0;
    return 0;
  }
  if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e->message().c_str());
  const int code = (int)e->code();
  delete e;
  return code;
}

static ffi::AnyBuffer buf(void* p, ffi::DataType dt, std::vector<int64_t> dims) {
  ffi::AnyBuffer b;
  b.data = p;
  b.dtype = dt;
  b.dims = std::move(dims);
  return b;
}

// handlers bound as (stream, u8 workspace, attr plan) -> u8 workspace: JaxentBvStep / ForwardInit / Pass / Tail / JaxentBgStep
extern "C" int ffi_drv_workspace_call(void* fn, void* stream, void* ws_in, void* ws_out, long long ws_bytes, long long plan, char* err,
                                      int errlen) {
  ffi::CallFrame f;
  f.stream = stream;
  f.args.push_back(buf(ws_in, ffi::U8, {ws_bytes}));
  f.rets.push_back(buf(ws_out, ffi::U8, {ws_bytes}));
  f.attrs["plan"] = (int64_t)plan;
  return finish(reinterpret_cast<HandlerFn>(fn)(&f), err, errlen);
}

// JaxentBvPack: (stream, f32 heavy (R,F), f32 acceptor (R,F)) -> f32 packed (n,)
extern "C" int ffi_drv_pack_call(void* fn, void* stream, float* heavy, float* acceptor, long long R, long long F, long long F_acceptor,
                                 float* packed, long long n_packed, char* err, int errlen) {
  ffi::CallFrame f;
  f.stream = stream;
  f.args.push_back(buf(heavy, ffi::F32, {R, F}));
  f.args.push_back(buf(acceptor, ffi::F32, {R, F_acceptor}));
  f.rets.push_back(buf(packed, ffi::F32, {n_packed}));
  return finish(reinterpret_cast<HandlerFn>(fn)(&f), err, errlen);
}

// JaxentBvPredict: (stream, packed, k_ints, timepoints; attrs) -> out
extern "C" int ffi_drv_predict_call(void* fn, void* stream, float* packed, long long n_packed, float* k_ints, float* tp, long long R,
                                    long long F, long long T, float bc, float bh, float* out, long long n_out, char* err, int errlen) {
  ffi::CallFrame f;
  f.stream = stream;
  f.args.push_back(buf(packed, ffi::F32, {n_packed}));
  f.args.push_back(buf(k_ints, ffi::F32, {R}));
  f.args.push_back(buf(tp, ffi::F32, {T > 0 ? T : 1}));
  f.rets.push_back(buf(out, ffi::F32, {n_out}));
  f.attrs["n_residues"] = (int64_t)R;
  f.attrs["n_frames"] = (int64_t)F;
  f.attrs["n_timepoints"] = (int64_t)T;
  f.attrs["bc"] = bc;
  f.attrs["bh"] = bh;
  return finish(reinterpret_cast<HandlerFn>(fn)(&f), err, errlen);
}

// a call frame with a missing attribute / wrong dtype: the binding layer must refuse it before the handler runs
extern "C" int ffi_drv_malformed_call(void* fn, void* stream, void* ws, long long ws_bytes, int which, char* err, int errlen) {
  ffi::CallFrame f;
  f.stream = stream;
  f.args.push_back(buf(ws, which == 1 ? ffi::F32 : ffi::U8, {ws_bytes}));
  f.rets.push_back(buf(ws, ffi::U8, {ws_bytes}));
  if (which == 2) f.attrs["plan"] = 1.0f;  // wrong attribute type
  if (which == 1) f.attrs["plan"] = (int64_t)1;
  return finish(reinterpret_cast<HandlerFn>(fn)(&f), err, errlen);
}
