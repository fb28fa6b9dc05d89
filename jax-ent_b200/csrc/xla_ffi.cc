// XLA FFI (jax.ffi) custom-call wrappers over the C ABI of include/jaxent_b200.h.
//
// Compiled only when jaxlib's header tree is on the include path
// (`-I $(python -c "import jax; print(jax.ffi.include_dir())")`); this image has no jax, so the
// file is inert here and is NOT part of libjaxent_b200.so.  Each handler only
// enqueues work on the stream XLA hands it; the plan handle travels as an i64 attribute and the
// workspace buffer is aliased input->output so XLA treats the step as an in-place update.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include "jaxent_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Check(int rc) {
  if (rc == JAXENT_OK) return ffi::Error::Success();
  const ffi::ErrorCode code = (rc == JAXENT_E_CUDA) ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument;
  return ffi::Error(code, jaxent_last_error());
}

// one optimise step: pass(1) + reduce + tail on a single device
static ffi::Error BvStepImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> workspace, int64_t plan,
                             ffi::Result<ffi::Buffer<ffi::U8>> workspace_out) {
  (void)workspace;
  (void)workspace_out;
  return Check(jaxent_bv_step(reinterpret_cast<JaxentPlan*>(plan), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxentBvStep, BvStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("plan")
                                  .Ret<ffi::Buffer<ffi::U8>>());

// the three phases separately, for shard_map + lax.psum between them (frame-sharded runs)
#define JAXENT_PHASE_HANDLER(NAME, CALL)                                                                 \
  static ffi::Error NAME##Impl(cudaStream_t stream, ffi::Buffer<ffi::U8> workspace, int64_t plan,        \
                               ffi::Result<ffi::Buffer<ffi::U8>> workspace_out) {                        \
    (void)workspace;                                                                                     \
    (void)workspace_out;                                                                                 \
    return Check(CALL);                                                                                  \
  }                                                                                                      \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(NAME, NAME##Impl,                                                        \
                                ffi::Ffi::Bind()                                                         \
                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()                            \
                                    .Arg<ffi::Buffer<ffi::U8>>()                                         \
                                    .Attr<int64_t>("plan")                                               \
                                    .Ret<ffi::Buffer<ffi::U8>>())

JAXENT_PHASE_HANDLER(JaxentBvForwardInit, jaxent_bv_forward_init(reinterpret_cast<JaxentPlan*>(plan), stream));
JAXENT_PHASE_HANDLER(JaxentBvPass, jaxent_bv_pass(reinterpret_cast<JaxentPlan*>(plan), 1, stream));
JAXENT_PHASE_HANDLER(JaxentBvReduce, jaxent_bv_reduce(reinterpret_cast<JaxentPlan*>(plan), stream));
JAXENT_PHASE_HANDLER(JaxentBvTail, jaxent_bv_tail(reinterpret_cast<JaxentPlan*>(plan), stream));

// one-off feature packing: (R,F) fp32 heavy / acceptor -> packed tile layout
static ffi::Error BvPackImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> heavy, ffi::Buffer<ffi::F32> acceptor,
                             ffi::Result<ffi::Buffer<ffi::F32>> packed) {
  const auto dims = heavy.dimensions();
  if (dims.size() != 2 || acceptor.dimensions() != dims)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "heavy/acceptor must both be (n_residues, n_frames)");
  const int64_t R = dims[0], F = dims[1];
  return Check(jaxent_bv_pack_features_f32(heavy.typed_data(), acceptor.typed_data(), (int32_t)R, F, F, 0, F,
                                           packed->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxentBvPack, BvPackImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// batched sweep (batch_optimise): one call advances every replica of a JaxentBatchPlan
static ffi::Error BgStepImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> workspace, int64_t plan,
                             ffi::Result<ffi::Buffer<ffi::U8>> workspace_out) {
  (void)workspace;
  (void)workspace_out;
  return Check(jaxent_bg_step(reinterpret_cast<JaxentBatchPlan*>(plan), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxentBgStep, BgStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("plan")
                                  .Ret<ffi::Buffer<ffi::U8>>());

// Simulation.predict: per-frame forward pass on the packed features (T = 0: log_Pf (R,F); T > 0: uptake (T,R,F))
static ffi::Error BvPredictImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> packed, ffi::Buffer<ffi::F32> k_ints,
                                ffi::Buffer<ffi::F32> timepoints, int64_t n_residues, int64_t n_frames, int64_t n_timepoints, float bc,
                                float bh, ffi::Result<ffi::Buffer<ffi::F32>> out) {
  return Check(jaxent_bv_predict_frames(packed.typed_data(), (int32_t)n_residues, n_frames, (int32_t)n_timepoints, k_ints.typed_data(),
                                        timepoints.typed_data(), bc, bh, out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxentBvPredict, BvPredictImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("n_residues")
                                  .Attr<int64_t>("n_frames")
                                  .Attr<int64_t>("n_timepoints")
                                  .Attr<float>("bc")
                                  .Attr<float>("bh")
                                  .Ret<ffi::Buffer<ffi::F32>>());
#endif
#endif
