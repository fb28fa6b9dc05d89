// XLA FFI (jax.ffi) custom-call wrappers over the C ABI of include/jaxent_b200.h.
//
// Compiled only when an xla/ffi/api/ffi.h is on the include path: jaxlib's
// (`-I $(python -c "import jax; print(jax.ffi.include_dir())")`), or -- in this image, which has no jax --
// the minimal stand-in under tests/ffi_stub that the test-suite builds it against (type-checks the handlers
// and drives them through a plain call frame, tests/test_ffi_*.py).  It is NOT part of libjaxent_b200.so
// (see INTEGRATION.md).  Each handler only enqueues work on the stream XLA hands it; the plan travels as a
// handle (an i64 attribute, jaxent_handle_register) and the workspace buffer is aliased input->output so XLA
// treats the step as an in-place update.
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

// Resolves the plan of a single-ensemble call and ties it to the buffers XLA actually passed:
//  * the `plan` attribute is a handle from jaxent_handle_register, not a pointer: a handle whose plan has been destroyed
//    is rejected here instead of being dereferenced;
//  * a step updates the workspace in place, so operand 0 and result 0 must be the same buffer
//    (ffi_call(..., input_output_aliases={0: 0})); without the alias XLA would hand out a fresh, uninitialised result;
//  * if XLA could not donate the operand it passes a COPY of the workspace at a new address: the plan is rebound to it
//    (same contents, every derived pointer re-based) rather than silently updating the stale allocation.
static ffi::Error ResolveBv(int64_t handle, ffi::Buffer<ffi::U8>& workspace, ffi::Result<ffi::Buffer<ffi::U8>>& workspace_out,
                            JaxentPlan** plan) {
  JaxentPlan* p = static_cast<JaxentPlan*>(jaxent_handle_lookup(handle, JAXENT_HANDLE_BV_PLAN));
  if (!p) return ffi::Error(ffi::ErrorCode::kInvalidArgument, jaxent_last_error());
  if (workspace_out->untyped_data() != workspace.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "the workspace operand and result must alias (input_output_aliases={0: 0})");
  void* base = nullptr;
  size_t bytes = 0;
  if (int rc = jaxent_bv_plan_workspace(p, &base, &bytes)) return Check(rc);
  if (workspace.size_bytes() < bytes) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "workspace buffer smaller than the plan's workspace");
  if (workspace.untyped_data() != base)
    if (int rc = jaxent_bv_plan_rebind_workspace(p, workspace.untyped_data(), workspace.size_bytes())) return Check(rc);
  *plan = p;
  return ffi::Error::Success();
}

#define JAXENT_BV_HANDLER(NAME, CALL)                                                                    \
  static ffi::Error NAME##Impl(cudaStream_t stream, ffi::Buffer<ffi::U8> workspace, int64_t plan,        \
                               ffi::Result<ffi::Buffer<ffi::U8>> workspace_out) {                        \
    JaxentPlan* p = nullptr;                                                                             \
    ffi::Error e = ResolveBv(plan, workspace, workspace_out, &p);                                        \
    if (e.failure()) return e;                                                                           \
    return Check(CALL);                                                                                  \
  }                                                                                                      \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(NAME, NAME##Impl,                                                        \
                                ffi::Ffi::Bind()                                                         \
                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()                            \
                                    .Arg<ffi::Buffer<ffi::U8>>()                                         \
                                    .Attr<int64_t>("plan")                                               \
                                    .Ret<ffi::Buffer<ffi::U8>>())

// one optimise step (pass + tail) on a single device; the phases separately for shard_map + lax.psum between them
JAXENT_BV_HANDLER(JaxentBvStep, jaxent_bv_step(p, stream));
JAXENT_BV_HANDLER(JaxentBvForwardInit, jaxent_bv_forward_init(p, stream));
JAXENT_BV_HANDLER(JaxentBvPass, jaxent_bv_pass(p, 1, stream));
JAXENT_BV_HANDLER(JaxentBvTail, jaxent_bv_tail(p, stream));

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

// batched sweep (batch_optimise): one call advances every replica of a JaxentBatchPlan.  The batched plan is not
// re-bindable: a workspace that moved is an error (donate the operand).
static ffi::Error BgStepImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> workspace, int64_t plan,
                             ffi::Result<ffi::Buffer<ffi::U8>> workspace_out) {
  JaxentBatchPlan* p = static_cast<JaxentBatchPlan*>(jaxent_handle_lookup(plan, JAXENT_HANDLE_BATCH_PLAN));
  if (!p) return ffi::Error(ffi::ErrorCode::kInvalidArgument, jaxent_last_error());
  void* base = nullptr;
  size_t bytes = 0;
  if (int rc = jaxent_bg_plan_workspace(p, &base, &bytes)) return Check(rc);
  if (workspace.untyped_data() != base || workspace_out->untyped_data() != base || workspace.size_bytes() < bytes)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument,
                      "the batched plan's workspace must be passed and aliased in place (donate_argnums + input_output_aliases={0: 0})");
  return Check(jaxent_bg_step(p, stream));
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
