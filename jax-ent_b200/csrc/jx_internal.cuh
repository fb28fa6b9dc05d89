// Internal declarations shared by the CUDA translation units.  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "jaxent_b200.h"

namespace jx {

constexpr int FT = 8;           // frames per tile (2 quads of 4 frames)
constexpr int QT = FT / 4;      // quads per tile
constexpr int MAX_STAGES = 8;   // bulk-copy ring depth upper bound
constexpr int U8_SUB_TILES = 4; // u8 features: one bulk-copy stage carries 4 tiles of 8 frames (32 frames)
constexpr int PASS_THREADS_MAX = 512;
constexpr int TAIL_THREADS = 1024;
constexpr int KL_N = 4;         // MaxEnt partial sums: sum w*kappa, sum p(log p - log q), sum(q - p), sum w
constexpr int FX_REP = 4;       // replicas of the fixed-point row-sum accumulators (CTA c adds into replica c % FX_REP)

// Device-resident control block; two copies, indexed by (epoch & 1).  Written only by the tail
// kernel (and state_init), read by the pass kernel that follows it.
struct alignas(16) Ctl {
  int step;         // number of updates applied to the current parameters (reference state.step)
  int mask_idx;     // optimizer._current_gradient_mask_idx
  int branch;       // which plateau branch of state set (epoch & 1) holds the current z/m/v
  int pending;      // a speculative update (pass mode 1) precedes the next tail
  int have_prev;    // state.gradients is not None
  int count_frame;  // optax adam count of the "frame" label
  int count_model;  // optax adam count of the "model" label
  int kl_slot;      // index of the MaxEnt slot or -1
  float lr, model_lr;          // optimizer._current_lr / _current_model_lr
  float bc, bh;                // BV parameters
  float mb[2], vb[2];          // Adam moments of (bc, bh)
  float gb_prev[2], gb[2];     // masked d loss / d (bc,bh): previous and current step
  // prepared for the pass that follows:
  float lrA, lrB, mlrA, mlrB;  // speculative learning rates (base, base / plateau_denominator)
  float mask_frame, mask_model;
  float bc1, bc2;              // 1 - b1^t, 1 - b2^t of the coming frame Adam update
  float lse;                   // log-sum-exp shift: exp(z - lse) are the softmax weights
  float kl_scale;              // fw*fs of the MaxEnt slot
  float last_dot, last_lr;
  int last_branch;
  int rec_idx;                 // ring index of this step's loss record
  double hbar_data;            // data part of hbar = sum_j w_j H_j MINUS what the pass subtracts itself (sum_r fl(k_r), see tail_body)
  float fw[JAXENT_MAX_SLOTS], fs[JAXENT_MAX_SLOTS];
  // ---- on-device convergence tracking (reference opt/run.py:272-348, opt/track.py:128-197)
  int trk_stop;                // 1: the loop has ended; pass / reduce / tail launches become no-ops
  int trk_reason;              // 1 all thresholds completed, 2 tolerance reached / non-finite loss
  int trk_stop_step;           // loop index of the step that ended the loop
  int trk_have_prev, trk_have_ema;
  int trk_steps_since, trk_idx, trk_nsnap;
  float trk_ema_bc, trk_ema_bh;
  int epoch_id;                // the epoch this block belongs to: lets the tail pick the current block before the pass has finished
  int trk_pad;
  double trk_prev_loss, trk_ema;
};
static_assert(sizeof(Ctl) % 16 == 0, "Ctl is staged to shared memory with 16-byte vector copies");

// one data slot's peptide maps / targets gathered into one contiguous word array (see bv_tail.cu)
struct BlobLayout {
  int tr_ptr, tr_col, tr_tptr, tr_trow, va_ptr, va_col, tr_val, tr_tval, tr_y, va_val, va_y, words;
};
struct BlobSet {
  int* blob[JAXENT_MAX_SLOTS];
  int words[JAXENT_MAX_SLOTS];
};

// ---- exchange buffer of the frame-sharded step (one per rank, mapped into every peer; see jaxent_b200.h)
//   plain  red [part_n] double, kl [KL_N] double              world == 1 (also what an NCCL-driven run all-reduces)
//   LL     red [2 parities][world][part_n] x 16 B             written by every rank's reduce kernel (slot = source rank)
//          kl  [2 parities][world][KL_N]  x 16 B              written by every rank's tail kernel
//   err    u64                                                 set when a wait timed out
// LL ("low latency") words carry their own validity stamp, as in NCCL's LL protocol: a double travels as one
// 16-byte store {lo, stamp, hi, stamp} whose two 8-byte halves are each written atomically, the stamp being the
// epoch of the step.  The consumer spins on the word itself until both stamps match -- no flags, no memory fences
// and no ordering requirement between words, so an exchange costs one NVLink store latency.  Parity double
// buffering keeps a fast rank's stores of epoch e+2 out of the slots a slow rank may still be reading for epoch e
// (a rank cannot publish e+2 before every rank has published e+1, i.e. finished reading e).
struct Xchg {
  int world, rank, part_n, pad;
  unsigned char* local;                    // == base[rank]; kernels never index base[] dynamically (the struct is a
  unsigned char* base[JAXENT_MAX_RANKS];   // by-value kernel parameter: a dynamic index would force a local-memory copy)
};
__host__ __device__ inline size_t xchg_plain_red_off(const Xchg&) { return 0; }
__host__ __device__ inline size_t xchg_plain_kl_off(const Xchg& x) { return (size_t)x.part_n * 8; }
__host__ __device__ inline size_t xchg_ll_base(const Xchg& x) { return ((size_t)(x.part_n + KL_N) * 8 + 255) / 256 * 256; }
__host__ __device__ inline size_t xchg_red_off(const Xchg& x, int par, int src) {
  return xchg_ll_base(x) + ((size_t)(par * x.world + src) * x.part_n) * 16;
}
__host__ __device__ inline size_t xchg_kl_off(const Xchg& x, int par, int src) {
  return xchg_ll_base(x) + (size_t)2 * x.world * x.part_n * 16 + (size_t)(par * x.world + src) * KL_N * 16;
}
__host__ __device__ inline size_t xchg_err_off(const Xchg& x) {
  return xchg_ll_base(x) + (x.world > 1 ? (size_t)2 * x.world * (x.part_n + KL_N) * 16 : 0);
}
__host__ __device__ inline size_t xchg_bytes(int part_n, int world) {
  Xchg x;
  x.world = world;
  x.part_n = part_n;
  return (xchg_err_off(x) + 8 + 255) / 256 * 256;
}

#ifdef __CUDACC__
__device__ __forceinline__ void ll_store(unsigned char* slot, double v, unsigned stamp) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"(lo), "r"(stamp), "r"(hi), "r"(stamp) : "memory");
}
// bounded spin (a few seconds): a peer that never arrives must not hang the GPU; the error word is raised instead
__device__ __forceinline__ double ll_load(unsigned char* err_word, const unsigned char* slot, unsigned stamp) {
  unsigned a = 0, b = 0, c = 0, d = 0;
  for (long long i = 0; i < (1ll << 24); ++i) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(slot) : "memory");
    if (b == stamp && d == stamp) return __hiloint2double((int)c, (int)a);
    if (i > 1024) __nanosleep(100);
  }
  *reinterpret_cast<volatile unsigned long long*>(err_word) = 1ull;
  return __hiloint2double((int)c, (int)a);
}
// element j of the partial sums of epoch `ep`, summed over the ranks in rank order (identical on every rank).
// All ranks' words are requested at once (independent loads, one L2 round trip) and, while any of them has not landed,
// ALL are requested again together: the first poll of a step normally precedes the peers' stores, and re-polling the
// missing words one after the other cost one round trip per word (8 ranks x 4 columns per thread: ~10 us, measured).
static __device__ __noinline__ double red_total_ll(const unsigned char* slot0, size_t rstride, int world, unsigned ep, unsigned char* err_word) {
  uint4 w[JAXENT_MAX_RANKS];
  for (long long it = 0;; ++it) {
#pragma unroll
    for (int r = 0; r < JAXENT_MAX_RANKS; ++r)
      if (r < world) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[r].x), "=r"(w[r].y), "=r"(w[r].z), "=r"(w[r].w) : "l"(slot0 + r * rstride) : "memory");
    bool ok = true;
#pragma unroll
    for (int r = 0; r < JAXENT_MAX_RANKS; ++r)
      if (r < world) ok = ok && (w[r].y == ep) && (w[r].w == ep);
    if (ok) break;
    if (it > (1ll << 22)) {  // bounded (seconds): a peer that never arrives must not hang the GPU
      *reinterpret_cast<volatile unsigned long long*>(err_word) = 1ull;
      break;
    }
    if (it > 4096) __nanosleep(100);
  }
  double t = 0.0;
#pragma unroll
  for (int r = 0; r < JAXENT_MAX_RANKS; ++r)
    if (r < world) t += __hiloint2double((int)w[r].z, (int)w[r].x);
  return t;
}
__device__ __forceinline__ double red_total(const Xchg& x, unsigned ep, int j) {
  if (x.world == 1) return __ldcg(reinterpret_cast<const double*>(x.local + xchg_plain_red_off(x)) + j);
  return red_total_ll(x.local + xchg_red_off(x, (int)(ep & 1u), 0) + (size_t)j * 16, (size_t)x.part_n * 16, x.world, ep, x.local + xchg_err_off(x));
}
__device__ __forceinline__ void red_publish(const Xchg& x, unsigned ep, int j, double v) {
  if (x.world == 1) {
    reinterpret_cast<double*>(x.local + xchg_plain_red_off(x))[j] = v;
    return;
  }
  const size_t off = xchg_red_off(x, (int)(ep & 1u), x.rank) + (size_t)j * 16;
#pragma unroll
  for (int r = 0; r < JAXENT_MAX_RANKS; ++r)
    if (r < x.world) ll_store(x.base[r] + off, v, ep);
}
__device__ __forceinline__ void kl_publish(const Xchg& x, unsigned ep, int k, double v) {
  if (x.world == 1) {
    reinterpret_cast<double*>(x.local + xchg_plain_kl_off(x))[k] = v;
    return;
  }
  const size_t off = xchg_kl_off(x, (int)(ep & 1u), x.rank) + (size_t)k * 16;
#pragma unroll
  for (int r = 0; r < JAXENT_MAX_RANKS; ++r)
    if (r < x.world) ll_store(x.base[r] + off, v, ep);
}
// MaxEnt sums of epoch `ep`, gathered by one whole warp (spins on the peers' words when sharded)
__device__ __forceinline__ void kl_gather_warp(const Xchg& x, unsigned ep, int lane, double (&kl)[KL_N]) {
  const int W = x.world;
  double v = 0.0;
  if (lane < W * KL_N) {
    const int r = lane / KL_N, k = lane % KL_N;
    if (W == 1) v = __ldcg(reinterpret_cast<const double*>(x.local + xchg_plain_kl_off(x)) + k);
    else v = ll_load(x.local + xchg_err_off(x), x.local + xchg_kl_off(x, (int)(ep & 1u), r) + (size_t)k * 16, ep);
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < KL_N; ++k) kl[k] = 0.0;
  for (int r = 0; r < W; ++r) {
#pragma unroll
    for (int k = 0; k < KL_N; ++k) kl[k] += __shfl_sync(0xffffffffu, v, r * KL_N + k);
  }
}
#endif

#ifdef __CUDACC__
// ---- fixed-point row sums of the pass
// |row sum| <= (sum of the un-normalised weights) * max |feature|.  In a step the weights are exp(z' - lse) with
// sum exp(z - lse) = 1 and |z' - z| <= (1 - b1) / sqrt(1 - b2) < 3.2 for optax adam (the update does not depend on the
// learning rate, which pre-scales the gradient): sum e' < e^3.2 < 64.  In the init pass e = exp(z0 - max z0) <= 1 per frame.
// scale = 2^(61 - e) with bound < 2^e, so a sum stays below 2^61 and carries >= 2^-61 of the bound as resolution
// (the fp32 tile sums that enter are good to 2^-24).
__device__ __forceinline__ double fx_scale(float nmax, int mode, long long F) {
  const double bound = (mode ? 64.0 : (double)F) * (double)fmaxf(nmax, 1e-30f);
  const int e = ((__double2hiint(bound) >> 20) & 0x7ff) - 1022;  // bound < 2^e
  return __hiloint2double((61 - e + 1023) << 20, 0);
}
__device__ __forceinline__ void fx_add(unsigned long long* dst, float v, double scale) {
  atomicAdd(dst, (unsigned long long)__double2ll_rn((double)v * scale));  // REDG.ADD.64: fire and forget
}
// 1 / scale of the sums a pass of `mode` left behind (NaN when a feature is not finite: integers cannot carry it, the tail
// then sees NaN sums and the loss becomes non-finite -- the reference treats that as data, not as an error)
__device__ __forceinline__ double fx_inv_scale(float nmax, bool per_frame_mode, int mode, long long F) {
  const double scale = fx_scale(per_frame_mode ? 1.f : nmax, mode, F);
  return (nmax <= 3.0e38f) ? 1.0 / scale : __longlong_as_double(0x7ff8000000000000ll);
}
// one column of the accumulators: the FX_REP replicas are integers, any order of addition gives the same sum.  Every
// column has exactly ONE reader per step, which also clears it for the next pass (the next pass cannot add to it before the
// reading kernel has completed).
__device__ __forceinline__ long long fx_load_col(const unsigned long long* fx, int n_rows, int col) {
  unsigned long long v[FX_REP];
#pragma unroll
  for (int r = 0; r < FX_REP; ++r) v[r] = __ldcg(fx + (size_t)r * n_rows + col);
  long long t = 0;
#pragma unroll
  for (int r = 0; r < FX_REP; ++r) t += (long long)v[r];
  return t;
}
__device__ __forceinline__ void fx_clear_col(unsigned long long* fx, int n_rows, int col) {
#pragma unroll
  for (int r = 0; r < FX_REP; ++r) fx[(size_t)r * n_rows + col] = 0ull;
}
// the per-CTA scalars (0 sum e' branch A, 1 branch B, 2 grad-dot) summed over the pass CTAs by ONE WARP in a fixed order
// (lane-strided, then a butterfly): every warp that evaluates this gets the same bits.  All loads of a lane are issued
// before the first addition (one L2 round trip for the whole fold).
constexpr int MAX_PASS_CTAS = 2 * 160;
__device__ __forceinline__ void fold_scal3(const double* cta_scal, int n_ctas, int lane, double (&out)[3]) {
  constexpr int U = MAX_PASS_CTAS / 32;
  double2 ab[U];
  double g[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int c = lane + 32 * u;
    const bool ok = c < n_ctas;
    ab[u] = ok ? __ldcg(reinterpret_cast<const double2*>(cta_scal + (size_t)c * 4)) : make_double2(0.0, 0.0);
    g[u] = ok ? __ldcg(cta_scal + (size_t)c * 4 + 2) : 0.0;
  }
  double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    t0 += ab[u].x;
    t1 += ab[u].y;
    t2 += g[u];
  }
  for (int o = 16; o; o >>= 1) {
    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
  }
  out[0] = t0;
  out[1] = t1;
  out[2] = t2;
}
#endif

struct StateSets {
  float* z[2][2];
  float* m[2][2];
  float* v[2][2];
  float* g[2];
};

struct PassArgs {
  const float* packed;
  int R;
  int mode;
  int stages;
  int part_n;         // 4R + 4
  int format;          // 0 fp32 features, 1 u8 features
  int sub_tiles;       // 8-frame tiles per bulk-copy stage
  uint32_t subtile_bytes;
  uint32_t tile_bytes; // bytes of one stage (= sub_tiles * subtile_bytes)
  uint32_t stage_stride;
  long long n_stage_units;
  long long n_tiles;
  long long F;
  StateSets st;
  const float* w;
  const float* kappa;
  const float* kappa_lo;  // kappa = kappa + kappa_lo carries dKL/dw to ~2^-48 relative (see frame_terms)
  const float* cc;
  const float* ch;
  const float* ck;          // [R] fl(cc_r <Nc>_r + ch_r <Nh>_r): subtracted inside every row's product (centred column sums)
  // per-frame uptake mode (ForwardPass.average_first = False): pf_T > 0
  int pf_T;                 // time points
  const float* pf_G;        // [T][R] dL/d<uptake>[t,r] from the tail
  const float* pf_k;        // [R] intrinsic rates
  const float* pf_tp;       // [T] time points
  // Row sums of all CTAs meet in ONE set of 64-bit fixed-point accumulators (integer atomics: exact, hence independent of
  // the order in which the CTAs arrive -> bitwise reproducible without a fixed-order reduction kernel).  The scale is a
  // power of two derived from a bound on the sums (see fx_scale()).  The kernel that follows (tail / reduce) converts and
  // clears them.
  unsigned long long* fx;   // [FX_REP][part_n - 4]
  double* cta_scal;         // [CTA][4]: sum e' (both branches), grad-dot -- summed in a fixed order by the consumer (fold_scal)
  const float* nmax;        // max |feature| of this rank's packed arrays (state_init)
  Xchg x;
  const Ctl* ctl;
  unsigned* epoch;
  JaxentLossRecord* records;
  float b1, b2, om_b1, om_b2, eps, clip;
  double* dbg;
  // per-frame uptake mode with optimised BV parameters: the forward half of the sweep uses the UPDATED parameters of each
  // plateau branch (derive_spec on the current control block), and a sweep of its own forms d loss / d (bc, bh) first
  int pf_db;
  JaxentOptConfig cfg;
};

// powf out of line: one copy per translation unit, the same instruction sequence wherever the bias corrections are formed
static __device__ __noinline__ float fpow(float a, float b) { return powf(a, b); }

// BV-parameter Adam update for BOTH plateau branches from the previous control block (lane b of the calling warp takes branch
// b; all 32 lanes must call).  Used by the tail (before the pass has finished) and by the per-frame pass when the BV
// parameters are optimised (its forward half needs the updated parameters of both branches): out of line, so both callers
// execute the same instruction sequence and agree bit for bit.
struct SpecScalars {
  float lr[2], mlr[2], bc[2], bh[2], mb0[2], mb1[2], vb0[2], vb1[2];
  double dot_model;  // sum over (bc, bh) of prev * cur gradient: the BV-parameter part of the grad-dot
};

static __device__ __noinline__ void derive_spec(SpecScalars* out, const Ctl* __restrict__ old, const JaxentOptConfig cfg) {
  const int lane = threadIdx.x & 31;
  const int cm1 = old->count_model + (old->pending ? 1 : 0);
  // lanes 0,1: b1^cm1, b2^cm1 (bias corrections of the BV-parameter Adam update)
  float pw = 0.f;
  if (lane < 2) pw = fpow((lane & 1) ? (float)cfg.b2 : (float)cfg.b1, (float)cm1);
  const float pw0 = __shfl_sync(0xffffffffu, pw, 0), pw1 = __shfl_sync(0xffffffffu, pw, 1);
  if (lane >= 2) return;
  const int b = lane;  // plateau branch
  const float g0 = old->gb[0], g1 = old->gb[1];
  const float lr = b ? old->lrB : old->lrA, mlr = b ? old->mlrB : old->mlrA;
  // optax adam + keep_params_nonnegative on (bc, bh) with the LR-pre-scaled gradients
  const float b1 = (float)cfg.b1, b2 = (float)cfg.b2, eps = (float)cfg.eps, clip = cfg.clip_value;
  const float om_b1 = (float)(1.0 - cfg.b1), om_b2 = (float)(1.0 - cfg.b2);
  const float c1 = 1.0f - pw0, c2 = 1.0f - pw1;
  float sg0 = g0 * mlr, sg1 = g1 * mlr;
  if (clip > 0.f) {
    sg0 = fminf(fmaxf(sg0, -clip), clip);
    sg1 = fminf(fmaxf(sg1, -clip), clip);
  }
  const float mb0 = om_b1 * sg0 + b1 * old->mb[0];
  const float mb1 = om_b1 * sg1 + b1 * old->mb[1];
  const float vb0 = om_b2 * (sg0 * sg0) + b2 * old->vb[0];
  const float vb1 = om_b2 * (sg1 * sg1) + b2 * old->vb[1];
  float u0 = -((mb0 / c1) / (sqrtf(vb0 / c2) + eps));
  float u1 = -((mb1 / c1) / (sqrtf(vb1 / c2) + eps));
  if (old->bc + u0 < 0.f) u0 = -old->bc;  // optax.keep_params_nonnegative
  if (old->bh + u1 < 0.f) u1 = -old->bh;
  out->lr[b] = lr;
  out->mlr[b] = mlr;
  out->bc[b] = old->bc + u0;
  out->bh[b] = old->bh + u1;
  out->mb0[b] = mb0;
  out->mb1[b] = mb1;
  out->vb0[b] = vb0;
  out->vb1[b] = vb1;
  if (b == 0) {
    const float p0 = old->have_prev ? old->gb_prev[0] : g0, p1 = old->have_prev ? old->gb_prev[1] : g1;
    out->dot_model = (double)p0 * (double)g0 + (double)p1 * (double)g1;
  }
}


constexpr int TRK_MAX_THR = 8;
// static configuration of the on-device loop (jaxent_bv_track_init)
struct TrackConfig {
  int on;
  int n_thr;
  int min_steps;
  int initial_steps;
  float thr[TRK_MAX_THR];     // descending, already multiplied by the learning rate (opt/track.py:12-21)
  float ema_alpha;
  float tolerance;
  float* ema_w;               // [F] EMA of the softmax weights
  float* snap_w;              // [n_thr + 1][F] snapshots of the softmax weights
  float* snap_ema;            // [n_thr + 1][F] snapshots of the EMA weights
  JaxentTrackSnapshot* snaps; // [n_thr + 1]
};

struct TailArgs {
  JaxentBvProblem prob;   // by value: device pointers + dims
  TrackConfig trk;
  JaxentOptConfig cfg;
  StateSets st;
  Xchg x;
  Ctl* ctl;
  unsigned* counter;
  unsigned* epoch;            // advanced by the tail (CTA 0) once the new control block is written
  // the pass's raw sums (see PassArgs) and how they reach this kernel
  unsigned long long* fx;     // [FX_REP][part_n - 4]
  const double* cta_scal;     // [pass CTA][4]
  const float* nmax;
  int pass_ctas;
  int fold_here;              // 1: this kernel completes the sums (and publishes them to the peers when frame-sharded);
                              // 0: jaxent_bv_reduce already wrote them to JAXENT_BUF_RED (three-phase ABI, NCCL in between)
  float* w;
  float* kappa;
  float* kappa_lo;
  float* cc;
  float* ch;
  float* ck;
  float* logpf;
  float* uptake;
  float* avg;
  float* pf_G;         // per-frame uptake mode: [T][R] dL/d<uptake> for the next pass
  int part_n;          // length of the reduced vector (4R+4, or 2TR+4 in per-frame uptake mode)
  double* klpart;
  JaxentLossRecord* records;
  BlobSet blobs;
  double* cl_scratch;  // 64 words of debug timestamps (JX_TAIL_PROFILE builds)
  int p_max_tr, p_max_va, nnz_max_tr, nnz_max_va;
  long long frames_per_cta;  // frames handled by each per-frame CTA of the tail kernel
  float eps_kl;   // 1e-10 / F_total
};

#ifdef JX_PROFILE
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define JPROF(buf, i) do { reinterpret_cast<long long*>(buf)[i] = jx::gtime(); } while (0)
#define JCLK_DECL(v) long long v = 0
#define JCLK_TIC(t) const long long t = clock64()
#define JCLK_TOC(acc, t) acc += clock64() - t
#define JCLK_OUT(buf, i, v) do { reinterpret_cast<long long*>(buf)[i] = (v); } while (0)
#else
#define JPROF(buf, i) do {} while (0)
#define JCLK_DECL(v) do {} while (0)
#define JCLK_TIC(t) do {} while (0)
#define JCLK_TOC(acc, t) do {} while (0)
#define JCLK_OUT(buf, i, v) do {} while (0)
#endif

void set_error(const char* fmt, ...);

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its
// predecessor in the stream is still draining; everything before pdl_wait() must only touch data the
// predecessor does not write (features, constants, shared memory).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename Kernel, typename Args>
inline cudaError_t launch_pdl(Kernel kernel, int grid, int block, size_t smem, cudaStream_t s, const Args& args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args);
}

int launch_pass(const PassArgs& a, int threads, int rpt, int ctas, size_t smem, cudaStream_t s);
constexpr int PF_DB_MAX_CTAS = 592;
int launch_pf_dbeta(const PassArgs& a, double* part, Ctl* ctl, JaxentLossRecord* records, cudaStream_t s);
// number of doubles every pass CTA contributes: row sums (+ sum e for both branches, grad-dot, pad)
__host__ __device__ inline int pass_part_n(int R, int T, int uptake_mode) { return uptake_mode == 2 ? 2 * T * R + 4 : 4 * R + 4; }
int launch_tail(const TailArgs& a, int ctas, size_t smem, cudaStream_t s);
int launch_reduce_fold(const TailArgs& a, cudaStream_t s);
int launch_finish_record(const Ctl* ctl, const unsigned* epoch, const Xchg& x, JaxentLossRecord* rec, int n_slots, cudaStream_t s);
size_t tail_smem_bytes(int R, int T, int p_tr, int p_va, int nnz_tr, int nnz_va);
int blob_words(int R, int T, int p_tr, int nnz_tr, int p_va, int nnz_va);
int launch_build_blobs(const JaxentBvProblem& pb, const BlobSet& bs, cudaStream_t s);
int configure_kernels();

__device__ __forceinline__ void finish_record(const Ctl* c, double kl_value, JaxentLossRecord* records, int n_slots) {
  JaxentLossRecord* r = records + c->rec_idx;
  if (c->kl_slot >= 0) {
    const float kl = (float)kl_value;
    r->train[c->kl_slot] = kl;
    r->val[c->kl_slot] = kl;
  }
  float tt = 0.f, tv = 0.f;
  for (int i = 0; i < n_slots; ++i) {
    const float st = r->train[i] * c->fw[i] * c->fs[i];
    const float sv = r->val[i] * c->fw[i] * c->fs[i];
    r->scaled_train[i] = st;
    r->scaled_val[i] = sv;
    tt += st;
    tv += sv;
  }
  r->total_train = tt;
  r->total_val = tv;
  r->complete = 1;
}

}  // namespace jx
