// Batched optimise step: B replicas (hyper-parameter sweep / replicates) that share one feature matrix.
//
// Reference: batch_optimise (opt/batch.py:51-192) vmaps _optimise_pure (opt/run.py:141-232) over
// (forward_model_weights, forward_model_scaling, learning_rate); every replica runs the optimise step of
// OptaxOptimizer._step_with_rates (opt/optimiser.py:337-445).  Here the two feature sweeps of all replicas are
// dense tensor-core GEMMs (bg_gemm.cu) and the per-(frame, replica) work is elementwise:
//
//   gemm1      H[f,b] = sum_{a,r} N_a[r,f] c_a[r,b]                         backward of step k
//   bt_grad    g = mask * w * (H + kappa - hbar), per-replica grad-dot      opt/optimiser.py:394-401
//   bt_decide  plateau rule -> learning rate, BV-parameter Adam, counters   opt/optimiser.py:403-424
//   bt_adam    optax adam on the frame logits, e' = exp(z' - lse), bf16 pieces of e' for gemm2
//   gemm2      Acc[a*Rp+r, b] = sum_f N_a[r,f] e'[f,b]                      forward of step k+1
//   bt_reduce2 k-split partials -> per-replica row sums (fp64) and S_b = sum_f e'
//   bt_tail    one CTA per replica: the R x T tail (bv_tail_body.cuh), c -> bf16 pieces for gemm1, loss record,
//              learning rates / masks of the coming update
//   bt_kl      maxent_convexKL_loss terms per (frame, replica), sums folded by the last block, loss record completed
//
// State arrays are [frame][replica] fp32 (replica contiguous): z, m, v, e, g (two buffers).  All reductions are
// carried in fp64 and folded in a fixed order (bitwise reproducible).
#include <cuda_bf16.h>

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "bg_mma.cuh"
#include "bv_tail_body.cuh"
#include "jx_internal.cuh"

namespace jx {
namespace bg {

constexpr int RB = 256;          // loss-record ring depth per replica
constexpr int MAX_EW_BLOCKS = 592;

struct alignas(16) BCtl {
  int step, mask_idx, have_prev, count_frame, count_model, kl_slot, stop, last_branch;
  float lr, model_lr, cfg_lr, lse;
  float bc, bh, mb[2], vb[2], gb_prev[2], gb[2];
  float lrA, lrB, mlrA, mlrB, mask_frame, mask_model;
  float lr_eff, bc1, bc2, kl_scale, last_dot, pad0;
  double S, hbar_data, zp;  // zp = sum_f p_f z_f of the current logits
  double lse_d, pad_d;      // log-sum-exp of the current logits, unrounded: (double)lse_old + log S (the MaxEnt loss needs more than fp32 here)
  double klsum[KL_N];
  float fw[JAXENT_MAX_SLOTS], fs[JAXENT_MAX_SLOTS];
  // ---- per-replica ConvergenceTracker (reference opt/track.py:128-197, opt/run.py:285-348); see derive_scalars (bv_tail.cu)
  int stopping;           // the loop of this replica ended at this step: the update is still applied and the new state
                          // evaluated (snapshot of the post-update weights), then `stop` freezes the replica
  int trk_reason, trk_stop_step, trk_have_prev, trk_have_ema, trk_steps_since, trk_idx, trk_nsnap;
  int snap_now, ema_update, ema_first, loop_step;
  int trk_have_best, trk_best_step, trk_pad_i[2];  // pure-loop mode: loop index whose validation loss is the best so far
  float trk_ema_bc, trk_ema_bh, trk_best_val, trk_pad;
  float thr[JAXENT_TRACK_MAX_THRESHOLDS];  // descending, already multiplied by this replica's learning rate
  double trk_prev_loss, trk_ema;
};
static_assert(sizeof(BCtl) % 16 == 0, "BCtl is copied with 16-byte loads");

struct BatchArgs {
  JaxentBvProblem prob;  // .packed = GF bf16 features
  JaxentOptConfig cfg;
  BlobSet blobs;
  int Bn, NS, n_fb, n_rg, Rp, ksplit;
  long long F;
  float *z, *m, *v, *e, *g[2];
  float* H;               // [Fp][Bn]
  float* P;               // [ksplit][2Rp][Bn]
  unsigned char* cc3;     // [NS][2][Rp/8][Bn][8] bf16
  unsigned char* e3;      // [n_fb][NS][16][Bn][8] bf16
  double* red;            // [Bn][2Rp]
  double* part_dot;       // [MAX_EW_BLOCKS][Bn]
  double* part_s;         // [MAX_EW_BLOCKS][Bn]
  double* part_kl;        // [MAX_EW_BLOCKS][KL_N][Bn]
  BCtl* ctl;              // [Bn]
  JaxentLossRecord* records;  // [RB][Bn]
  float* avg;             // [Bn][2R]
  float* logpf;           // [Bn][R]
  float* uptake;          // [Bn][T*R]
  const float* row_off;   // [2][Rp] integers the packed features were centred with (jaxent_bg_pack_features_centred), or nullptr
  float* pnorm;           // [F] normalised prior p_f = (|prior_f| + eps) / sum (fp32, as frame_terms forms it)
  double* kl_const;       // [2] sum_f p_f log p_f, sum_f p_f
  unsigned* gsel;         // which g buffer holds the current masked gradients
  unsigned* counter;      // ticket of bt_kl
  int p_max_tr, p_max_va;
  int blob_words_max;     // words reserved for the blob image in the tail's shared memory
  int ew_blocks, fy;      // elementwise grid (2 blocks of 512 threads per SM) / frames per block iteration
  int adam_blocks, adam_threads;  // bt_adam: one replica per thread (see the kernel)
  int dot_blocks;         // rows of part_dot written by the gradient kernel (bt_grad, or the fused tensor-core kernel)
  float eps_kl;
  // on-device loop (jaxent_bg_track_init)
  int trk_on, trk_n_thr, trk_min_steps;
  int trk_pure;  // 1: the functional tracker of _optimise_pure (opt/track.py:40-125), the loop batch_optimise vmaps; 0: ConvergenceTracker of the Python loop
  float trk_alpha, trk_tolerance;
  float* ema_w;                // [F][Bn] EMA of the softmax weights
  float* snap_w;               // [n_thr + 1][F][Bn]
  float* snap_ema;             // [n_thr + 1][F][Bn]
  JaxentTrackSnapshot* snaps;  // [n_thr + 1][Bn]
};

// fp32 -> NS bf16 pieces (round to nearest at every level): x ~= p0 + p1 + p2
__device__ __forceinline__ void split3(float x, int NS, __nv_bfloat16& p0, __nv_bfloat16& p1, __nv_bfloat16& p2) {
  p0 = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(p0);
  p1 = (NS > 1) ? __float2bfloat16_rn(r) : __float2bfloat16_rn(0.f);
  r -= __bfloat162float(p1);
  p2 = (NS > 2) ? __float2bfloat16_rn(r) : __float2bfloat16_rn(0.f);
}

// kappa (= d KL / d w, scaled) and the normalised weight of one (frame, replica).  Same quantities as frame_terms
// (bv_tail.cu), but p/q is carried as an unevaluated fp32 pair r0 + corr (r0 = fl(p/q), corr = fl((p - r0*q)/q) with the
// exact FMA residual) instead of an fp64 quotient: 256 replicas x 1e5 frames of fp64 division and log would make this
// kernel arithmetic bound.  1 - p/q = (1 - r0) - corr keeps full relative accuracy when p/q is close to 1.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// invS = 1 / S_b.  The quotient p/q only has to be *some* fp32 approximation r0: the FMA residual p - r0*q is exact, so
// r0 + corr carries p/q to ~1e-14 whatever the error of the approximate reciprocal.
// kappa is returned in fp64 (two conversions and three fp64 operations, no fp64 division): once a few frames dominate,
// g_f = w_f (H_f + kappa_f - hbar) is a difference of O(lambda) numbers and needs more than 24 bits (see frame_terms).
__device__ __forceinline__ void weight_terms(float e, float invS, float p, bool has_kl, float eps, float lam, float& w, double& kap, float& q) {
  w = e * invS;
  kap = 0.0;
  q = 0.f;
  if (has_kl) {
    q = fabsf(w) + eps;
    const float rq = rcp_approx(q);
    const float r0 = p * rq;
    const float corr = fmaf(-r0, q, p) * rq;
    const double G_ = (1.0 - (double)r0) - (double)corr;
    kap = (w > 0.f) ? (double)lam * G_ : 0.0;
  }
}

// The same kappa as weight_terms, as an unevaluated fp32 pair (kh + kl carries ~45 bits) and without a single fp64
// instruction or conversion: float <-> double conversions run at 16 lanes/clk on sm_100, and the fused gradient kernel has
// only its epilogue warps for this arithmetic.  1 - r0 is formed with its exact rounding error (2Sum), the product with
// lambda with its exact FMA residual.
__device__ __forceinline__ void weight_terms_pair(float e, float invS, float p, bool has_kl, float eps, float lam, float& w, float& kh, float& kl) {
  w = e * invS;
  kh = 0.f;
  kl = 0.f;
  if (has_kl && w > 0.f) {
    const float q = fabsf(w) + eps;
    const float rq = rcp_approx(q);
    const float r0 = p * rq;
    const float corr = fmaf(-r0, q, p) * rq;          // p/q = r0 + corr to ~1e-14
    const float gh = 1.0f - r0;                         // 2Sum(1, -r0): gh + ge == 1 - r0 exactly
    const float bb = gh - 1.0f;
    const float ge = (1.0f - (gh - bb)) + (-r0 - bb);
    const float gl = ge - corr;                         // G = 1 - p/q = gh + gl
    kh = lam * gh;
    kl = fmaf(lam, gh, -kh) + lam * gl;                 // lam * G = kh + kl
  }
}

// ------------------------------------------------------------------------------------------ bt_grad
// thread = (4 consecutive replicas bq = tid % (Bn/4), frame lane fy = tid / (Bn/4)): 16-byte loads / stores, per-replica
// constants in registers; frames strided over the grid, two frames (6 vector loads) in flight per thread
__global__ void __launch_bounds__(512) bt_grad_kernel(const BatchArgs a) {
  extern __shared__ double sh_d[];  // [fy][Bn]
  const int Bn = a.Bn, Bq = Bn >> 2, bq = threadIdx.x % Bq, fy = threadIdx.x / Bq, b0 = 4 * bq;
  const unsigned sel = *a.gsel;
  const float4* __restrict__ gprev = reinterpret_cast<const float4*>(a.g[sel]);
  float4* __restrict__ gout = reinterpret_cast<float4*>(a.g[sel ^ 1u]);
  const float4* __restrict__ ep = reinterpret_cast<const float4*>(a.e);
  const float4* __restrict__ Hp = reinterpret_cast<const float4*>(a.H);
  const float* __restrict__ pn = a.pnorm;
  // per-replica constants: one thread per replica reads its control block, everybody else takes them from shared memory
  // (a strided control-block read per thread costs more than the whole frame loop)
  __shared__ float s_f[3][256];
  __shared__ double s_hbar[256];
  __shared__ int s_i[256];
  for (int b = threadIdx.x; b < Bn; b += blockDim.x) {
    const BCtl* c = a.ctl + b;
    s_f[0][b] = (float)(1.0 / c->S);
    s_f[1][b] = c->kl_scale;
    s_f[2][b] = c->mask_frame;
    s_hbar[b] = c->hbar_data + c->klsum[0];  // hbar = sum_j w_j (H_j + kappa_j): data part + MaxEnt part (see gather_hbar, bv_pass.cu)
    s_i[b] = (c->have_prev ? 1 : 0) | ((c->kl_slot >= 0 && a.prob.prior != nullptr) ? 2 : 0) | (c->stop ? 0 : 4);
  }
  __syncthreads();
  float invS[4], lam[4], maskf[4];
  double hbar[4];
  int have_prev[4];
  bool has_kl[4], active[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    invS[j] = s_f[0][b0 + j];
    lam[j] = s_f[1][b0 + j];
    maskf[j] = s_f[2][b0 + j];
    hbar[j] = s_hbar[b0 + j];
    const int fl = s_i[b0 + j];
    have_prev[j] = fl & 1;
    has_kl[j] = fl & 2;
    active[j] = fl & 4;
  }
  double dot[4] = {0.0, 0.0, 0.0, 0.0};
  const long long fstep = (long long)gridDim.x * a.fy;
  for (long long f0 = (long long)blockIdx.x * a.fy + fy; f0 < a.F; f0 += 2 * fstep) {
    float4 ev[2], hv[2], gv[2];
    float pv[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long f = f0 + k * fstep;
      const long long fc = f < a.F ? f : f0;
      ev[k] = ep[(size_t)fc * Bq + bq];
      hv[k] = Hp[(size_t)fc * Bq + bq];
      gv[k] = gprev[(size_t)fc * Bq + bq];
      pv[k] = pn ? pn[fc] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long f = f0 + k * fstep;
      if (f < a.F) {
        const float e4[4] = {ev[k].x, ev[k].y, ev[k].z, ev[k].w}, h4[4] = {hv[k].x, hv[k].y, hv[k].z, hv[k].w};
        const float g4[4] = {gv[k].x, gv[k].y, gv[k].z, gv[k].w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float w, q;
          double kap;
          weight_terms(e4[j], invS[j], pv[k], has_kl[j], a.eps_kl, lam[j], w, kap, q);
          const float gf = active[j] ? maskf[j] * (w * (float)(((double)h4[j] + kap) - hbar[j])) : 0.f;
          const float gpv = have_prev[j] ? g4[j] : gf;
          dot[j] += (double)(gpv * gf);
          o[j] = gf;
        }
        gout[(size_t)f * Bq + bq] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sh_d[fy * Bn + b0 + j] = dot[j];
  __syncthreads();
  for (int b = threadIdx.x; b < Bn; b += blockDim.x) {
    double t = 0.0;
    for (int k = 0; k < a.fy; ++k) t += sh_d[k * Bn + b];
    a.part_dot[(size_t)blockIdx.x * Bn + b] = t;
  }
}

// ------------------------------------------------------------------------------------------ fused gemm1 + gradient
// H never leaves the chip: the backward product runs TRANSPOSED on the tensor cores,
//     H^T[b, f] = sum_{a,r} c_a[r, b] * N_a[r, f]        M = 128 replicas, N = 256 frames, K = 2 Rp (x NS pieces of c)
// so that a TMEM lane is a REPLICA and the columns are frames.  An epilogue thread then owns one replica: its constants
// (1/S, lambda, mask, hbar) live in registers, its grad-dot is a plain register accumulation (no cross-lane reduction), and
// for a given frame the 32 lanes of a warp touch 32 consecutive replicas of the [frame][replica] state arrays -- every
// load and store is a coalesced 128-byte line.  The epilogue does what bt_grad did (weights, dKL/dw, masked gradient,
// grad-dot partial) straight from tcgen05.ld, overlapped with the MMAs of the next tile (two TMEM accumulators).
//   A operand  c pieces   global CC3 [s][a][kc][Bn][8]  -> smem [kc: 4][128 replicas][16 B]            K-major  (SBO 128, LBO 2048)
//   B operand  features   global GF  [fb][a][rg][16 fc][8 r][8 f] -> smem [rg: 4][fc: 32][8 r][8 f]    MN-major (SBO 128, LBO 4096)
//              (two frame blocks side by side: 32 chunks of 8 frames = N 256)
// 5 stages of K = 32 (8 KB of features + NS x 8 KB of c pieces) cover the L2 latency; features cross L2 -> SM twice (one
// tile per replica half), c pieces once per tile.  Used when Bn is a multiple of 128; otherwise gemm1 + bt_grad.
constexpr int FG_STAGE_K = 32;
constexpr int FG_STAGES = 5;
constexpr int FG_B_BYTES = 256 * FG_STAGE_K * 2;  // 16 KB: features, 256 frames x 32 rows
constexpr int FG_A_BYTES = 128 * FG_STAGE_K * 2;  // 8 KB per piece: 128 replicas x 32 rows
constexpr int FG_EPI_WARPS = 16;                    // 4 per TMEM lane quarter: enough loads in flight to stream the state arrays
constexpr int FG_THREADS = 64 + 32 * FG_EPI_WARPS;

__global__ void __launch_bounds__(FG_THREADS, 1) bt_fused_grad_kernel(const BatchArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Bn = a.Bn, NS = a.NS;
  const uint32_t stage_bytes = FG_B_BYTES + (uint32_t)NS * FG_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)FG_STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + FG_STAGES;
  uint64_t* tfull = bars + 2 * FG_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  double* s_dot = reinterpret_cast<double*>(tmem_slot + 4);  // [FG_EPI_WARPS][32]

  // tile = (pair of frame blocks, replica half); a CTA keeps ONE replica half (the grid is even when there are two)
  const int halves = Bn >> 7;
  const int n_fp = (a.n_fb + 1) >> 1;
  const int h = (halves == 2) ? (int)(blockIdx.x & 1u) : 0;
  const int fp0 = (halves == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int fp_step = (halves == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_stage = 2 * (a.n_rg / 4);  // both arrays, 32-row chunks

  if (tid == 0) {
    for (int s = 0; s < FG_STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull[s]), 1);
      mbar_init(smem_u32(&tempty[s]), FG_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // two accumulators of 256 fp32 columns: all of tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ producer
    if (lane == 0) {
      int slot = 0;
      uint32_t par = 0;
      const unsigned char* gf = reinterpret_cast<const unsigned char*>(a.prob.packed);
      for (int fp = fp0; fp < n_fp; fp += fp_step) {
        const int fb0 = 2 * fp, fb1 = (2 * fp + 1 < a.n_fb) ? 2 * fp + 1 : fb0;  // odd tail: the second half repeats the first (ignored: f >= F)
        for (int st = 0; st < n_stage; ++st) {
          const int arr = st / (a.n_rg / 4), rg0 = (st - arr * (a.n_rg / 4)) * 4;
          mbar_wait(smem_u32(&empty[slot]), par ^ 1u);
          const uint32_t bar = smem_u32(&full[slot]);
          unsigned char* dst = smem + (size_t)slot * stage_bytes;
          mbar_expect_tx(bar, stage_bytes);
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // features: 2 KB per (frame block, 8-row group) -> [rg][fc 0..15 | fc 16..31]
            bulk_g2s(smem_u32(dst + g * 4096), gf + (((size_t)fb0 * 2 + arr) * a.n_rg + rg0 + g) * 2048, 2048, bar);
            bulk_g2s(smem_u32(dst + g * 4096 + 2048), gf + (((size_t)fb1 * 2 + arr) * a.n_rg + rg0 + g) * 2048, 2048, bar);
          }
          for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int g = 0; g < 4; ++g)  // c pieces: 128 replicas x 16 B per k-chunk
              bulk_g2s(smem_u32(dst + FG_B_BYTES + s * FG_A_BYTES + g * 2048),
                       a.cc3 + ((((size_t)s * 2 + arr) * a.n_rg + rg0 + g) * Bn + (size_t)h * 128) * 16, 2048, bar);
          if (++slot == FG_STAGES) {
            slot = 0;
            par ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      // D fp32, A / B bf16, A K-major, B MN-major (bit 16), N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      int slot = 0, acc = 0;
      uint32_t par = 0, tpar = 0;
      for (int fp = fp0; fp < n_fp; fp += fp_step) {
        mbar_wait(smem_u32(&tempty[acc]), tpar ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * 256);
        uint32_t accumulate = 0;
        for (int st = 0; st < n_stage; ++st) {
          mbar_wait(smem_u32(&full[slot]), par);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)slot * stage_bytes);
#pragma unroll
          for (int j = 0; j < FG_STAGE_K / 16; ++j) {
            const uint64_t bd = make_desc(sa + j * 2 * 4096, 4096, 128);
            for (int s = 0; s < NS; ++s) {
              const uint64_t ad = make_desc(sa + FG_B_BYTES + s * FG_A_BYTES + j * 2 * 2048, 2048, 128);
              umma_bf16(d, ad, bd, idesc, accumulate);
              accumulate = 1;
            }
          }
          tc_commit(smem_u32(&empty[slot]));
          if (++slot == FG_STAGES) {
            slot = 0;
            par ^= 1u;
          }
        }
        tc_commit(smem_u32(&tfull[acc]));
        if (++acc == 2) {
          acc = 0;
          tpar ^= 1u;
        }
      }
    }
  } else {
    // ================================================================ epilogue: one replica per thread
    // 16 warps: warp (2 + w) reads TMEM lane quarter (2 + w) % 4 and takes the 16-frame column chunks j = (w / 4) mod 4 of a
    // tile.  The state loads of the NEXT chunk are issued before the current one is evaluated (one chunk = 2 x 16 coalesced
    // 128-byte lines per warp).
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    const int b = h * 128 + quarter * 32 + lane;
    const BCtl* c = a.ctl + b;
    const unsigned sel = *a.gsel;
    const float* __restrict__ gprev = a.g[sel];
    float* __restrict__ gout = a.g[sel ^ 1u];
    const float* __restrict__ ep = a.e;
    const float invS = (float)(1.0 / c->S), lam = c->kl_scale, maskf = c->mask_frame;
    const double hbar = c->hbar_data + c->klsum[0];  // data part + MaxEnt part (see gather_hbar, bv_pass.cu)
    const float hb_hi = (float)hbar, hb_lo = (float)(hbar - (double)hb_hi);
    const bool have_prev = c->have_prev != 0, has_kl = c->kl_slot >= 0 && a.prob.prior != nullptr, active = c->stop == 0;
    const float* __restrict__ pn = a.pnorm;
    const long long Fm1 = a.F - 1;
    double dot = 0.0;
    int acc = 0;
    uint32_t tpar = 0;
    float ev[16], gv[16], pv;
    auto load_chunk = [&](long long f0) {  // lanes = consecutive replicas of one frame: coalesced lines; lane k also fetches p of frame k
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const long long f = f0 + k;
        const size_t i = (size_t)(f < a.F ? f : Fm1) * Bn + b;
        ev[k] = ep[i];
        gv[k] = gprev[i];
      }
      const long long fl = f0 + (lane & 15);
      pv = pn ? pn[fl < a.F ? fl : Fm1] : 0.f;
    };
    if (fp0 < n_fp) load_chunk((long long)fp0 * 256 + sub * 16);
    for (int fp = fp0; fp < n_fp; fp += fp_step) {
      mbar_wait(smem_u32(&tfull[acc]), tpar);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
      const long long fbase = (long long)fp * 256;
#pragma unroll 1
      for (int j = sub; j < 16; j += 4) {
        const long long f0 = fbase + j * 16;
        uint32_t hv[16];
        tmem_ld16_issue(taddr + j * 16, hv);
        float ec[16], gc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          ec[k] = ev[k];
          gc[k] = gv[k];
        }
        const float pc = pv;
        {  // next chunk of this warp: same tile, or the first one of the next tile
          const bool more = j + 4 < 16;
          const long long fn = more ? f0 + 64 : (long long)(fp + fp_step) * 256 + sub * 16;
          if (more || fp + fp_step < n_fp) load_chunk(fn);
        }
        tmem_ld_wait();
        float prod[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const long long f = f0 + k;
          const float p = __shfl_sync(0xffffffffu, pc, k);
          prod[k] = 0.f;
          if (f < a.F) {
            float w, kh, kl;
            weight_terms_pair(ec[k], invS, p, has_kl, a.eps_kl, lam, w, kh, kl);
            // (H + kappa) - hbar with kappa and hbar as fp32 pairs: the O(lambda) cancellation kh - hb_hi is exact (see bv_pass.cu)
            const float gf = active ? maskf * (w * ((__uint_as_float(hv[k]) + (kh - hb_hi)) + (kl - hb_lo))) : 0.f;
            prod[k] = (have_prev ? gc[k] : gf) * gf;
            gout[(size_t)f * Bn + b] = gf;
          }
        }
        // 16 products: fixed-shape fp32 tree, then one fp64 addition per chunk
#pragma unroll
        for (int st = 8; st; st >>= 1)
#pragma unroll
          for (int k = 0; k < st; ++k) prod[k] += prod[k + st];
        dot += (double)prod[0];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty[acc]));
      if (++acc == 2) {
        acc = 0;
        tpar ^= 1u;
      }
    }
    s_dot[(warp - 2) * 32 + lane] = dot;
  }

  tc_fence_before();
  __syncthreads();
  if (warp >= 2 && warp < 6) {  // this CTA's row of the grad-dot partials (fixed order over the 4 column sub-sets); zeros for the other half
    const int quarter = warp & 3;
    double t = 0.0;
#pragma unroll
    for (int sb = 0; sb < 4; ++sb) {
      // the epilogue warp with lane quarter `quarter` and column sub-set sb is warp index 2 + w with (2 + w) % 4 == quarter, w / 4 == sb
      const int w = ((quarter + 2) & 3) + 4 * sb;
      t += s_dot[w * 32 + lane];
    }
    a.part_dot[(size_t)blockIdx.x * Bn + h * 128 + quarter * 32 + lane] = t;
    if (halves == 2) a.part_dot[(size_t)blockIdx.x * Bn + (h ^ 1) * 128 + quarter * 32 + lane] = 0.0;
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------ bt_decide
// one thread per replica: plateau rule, BV-parameter Adam (optax adam + keep_params_nonnegative), counters
__global__ void bt_decide_kernel(const BatchArgs a) {
  // one warp per replica: the per-block partials are summed lane-strided, then across lanes (fixed order)
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.gsel = *a.gsel ^ 1u;  // bt_grad wrote the other buffer: it now holds the current gradients
  if (b >= a.Bn) return;
  BCtl* c = a.ctl + b;
  if (c->stop) return;
  const JaxentOptConfig& cfg = a.cfg;
  double gdot = 0.0;
  for (int k = lane; k < a.dot_blocks; k += 32) gdot += a.part_dot[(size_t)k * a.Bn + b];
  gdot = wsum(gdot);
  if (lane != 0) return;
  if (a.trk_on && a.trk_pure && c->step == 0) {
    // cond_fn of _optimise_pure (opt/run.py:202-211) looks at the loss of the dense initial state before the first iteration
    const double l0 = (double)a.records[b].total_train;
    if (l0 < (double)a.trk_tolerance || isnan(l0) || isinf(l0)) {
      c->stop = 1;
      c->trk_reason = 2;
      c->trk_stop_step = 0;
      c->snap_now = -1;
      c->ema_update = 0;
      return;
    }
  }
  const float g0 = c->gb[0], g1 = c->gb[1];
  const float p0 = c->have_prev ? c->gb_prev[0] : g0, p1 = c->have_prev ? c->gb_prev[1] : g1;
  const double dot = gdot + (double)p0 * (double)g0 + (double)p1 * (double)g1;
  const int d = ((float)dot < 0.f && c->step > 1) ? 1 : 0;  // opt/optimiser.py:403
  const float lr = d ? c->lrB : c->lrA, mlr = d ? c->mlrB : c->mlrA;
  const float b1 = (float)cfg.b1, b2 = (float)cfg.b2, eps = (float)cfg.eps, clip = cfg.clip_value;
  const float om_b1 = (float)(1.0 - cfg.b1), om_b2 = (float)(1.0 - cfg.b2);
  const int cm = c->count_model + 1, cf = c->count_frame + 1;
  const float c1 = 1.0f - fpow(b1, (float)cm), c2 = 1.0f - fpow(b2, (float)cm);
  float sg0 = g0 * mlr, sg1 = g1 * mlr;
  if (clip > 0.f) {
    sg0 = fminf(fmaxf(sg0, -clip), clip);
    sg1 = fminf(fmaxf(sg1, -clip), clip);
  }
  const float mb0 = om_b1 * sg0 + b1 * c->mb[0], mb1 = om_b1 * sg1 + b1 * c->mb[1];
  const float vb0 = om_b2 * (sg0 * sg0) + b2 * c->vb[0], vb1 = om_b2 * (sg1 * sg1) + b2 * c->vb[1];
  float u0 = -((mb0 / c1) / (sqrtf(vb0 / c2) + eps));
  float u1 = -((mb1 / c1) / (sqrtf(vb1 / c2) + eps));
  if (c->bc + u0 < 0.f) u0 = -c->bc;  // optax.keep_params_nonnegative
  if (c->bh + u1 < 0.f) u1 = -c->bh;
  c->bc += u0;
  c->bh += u1;
  c->mb[0] = mb0; c->mb[1] = mb1; c->vb[0] = vb0; c->vb[1] = vb1;
  c->gb_prev[0] = g0; c->gb_prev[1] = g1;
  c->count_model = cm;
  c->count_frame = cf;
  c->have_prev = 1;
  c->mask_idx = (c->step >= cfg.initial_steps) ? 1 : c->mask_idx;
  c->step += 1;
  c->lr = lr;
  c->model_lr = mlr;
  c->lr_eff = lr;
  c->bc1 = 1.0f - fpow(b1, (float)cf);
  c->bc2 = 1.0f - fpow(b2, (float)cf);
  c->last_dot = (float)dot;
  c->last_branch = d;

  // ---- ConvergenceTracker for loop step k = state.step before this update (same logic as derive_scalars, bv_tail.cu)
  c->snap_now = -1;
  c->ema_update = 0;
  c->ema_first = 0;
  if (a.trk_on && a.trk_pure) {
    // update_convergence / check_and_advance_threshold (opt/track.py:40-125) inside OptaxOptimizer._pure_step
    // (opt/optimiser.py:494-579), fp32 like the reference's carry.  k = state.step before this update.
    const int k = c->step - 1;
    c->loop_step = k;
    const JaxentLossRecord& rk = a.records[(size_t)(k % RB) * a.Bn + b];
    const float loss = rk.total_train;
    const float alpha = a.trk_alpha;
    // previous_loss = the loss stored by iteration k - 1; the dense initial state carries the loss at the initial parameters,
    // which is also what iteration 0 evaluates: first raw delta = 0
    const float prev = c->trk_have_prev ? (float)c->trk_prev_loss : loss;
    const float raw = fabsf(prev - loss);
    const bool first = c->trk_steps_since == 0;  // the EMAs restart at the start and after every threshold advance
    const float ema = first ? raw : alpha * raw + (1.0f - alpha) * (float)c->trk_ema;
    c->trk_ema = (double)ema;
    c->trk_have_ema = 1;
    c->ema_update = 1;
    c->ema_first = first ? 1 : 0;
    c->trk_ema_bc = first ? c->bc : alpha * c->bc + (1.0f - alpha) * c->trk_ema_bc;
    c->trk_ema_bh = first ? c->bh : alpha * c->bh + (1.0f - alpha) * c->trk_ema_bh;
    c->trk_steps_since += 1;
    c->trk_prev_loss = (double)loss;
    c->trk_have_prev = 1;
    const int ti = c->trk_idx < a.trk_n_thr ? c->trk_idx : a.trk_n_thr - 1;
    const float rel = loss > 0.f ? ema / loss : 0.f;
    const bool met = c->trk_steps_since >= a.trk_min_steps && rel < c->thr[ti] && c->step > cfg.initial_steps;  // 1-based state.step
    if (met && c->trk_idx + 1 < a.trk_n_thr) {
      c->trk_idx += 1;
      c->trk_steps_since = 0;
    } else if (met) {
      c->stopping = 1;
      c->trk_reason = 1;
      c->trk_stop_step = k;
    }
    if (!c->stopping && (loss < a.trk_tolerance || isnan(loss) || isinf(loss))) {  // cond_fn before iteration k + 1
      c->stopping = 1;
      c->trk_reason = 2;
      c->trk_stop_step = k;
    }
    // every iteration enters the reference's dense history (save_state = post-update weights, losses of step k) and
    // get_best_state takes the smallest first validation loss (opt/base.py:180-189 over all steps, opt/batch.py:158-180): slot 0
    // of the snapshot buffers follows the running minimum (strictly smaller: the first step that reaches it)
    const float val0 = rk.val[0];
    if (!c->trk_have_best || val0 < c->trk_best_val) {
      c->trk_have_best = 1;
      c->trk_best_val = val0;
      c->trk_best_step = k;
      c->snap_now = 0;
      c->trk_nsnap = 1;
      if (a.snaps != nullptr) {
        JaxentTrackSnapshot* sn = a.snaps + b;
        sn->loop_step = k;
        sn->state_step = c->step;
        sn->have_ema = 0;
        sn->threshold_idx = c->trk_idx;
        sn->bc = c->bc;
        sn->bh = c->bh;
        sn->ema_bc = c->trk_ema_bc;
        sn->ema_bh = c->trk_ema_bh;
        sn->losses = rk;
      }
    }
  } else if (a.trk_on) {
    const int k = c->step - 1;
    c->loop_step = k;
    const double loss = (double)a.records[(size_t)(k % RB) * a.Bn + b].total_train;
    const double alpha = (double)a.trk_alpha;
    if (c->trk_have_prev) {  // tracker.update
      const double raw = fabs(c->trk_prev_loss - loss);
      if (!c->trk_have_ema) {
        c->trk_ema = raw;
        c->trk_have_ema = 1;
        c->ema_first = 1;
      } else {
        c->trk_ema = alpha * raw + (1.0 - alpha) * c->trk_ema;
      }
      c->ema_update = 1;
      c->trk_ema_bc = c->ema_first ? c->bc : (float)(alpha * (double)c->bc + (1.0 - alpha) * (double)c->trk_ema_bc);
      c->trk_ema_bh = c->ema_first ? c->bh : (float)(alpha * (double)c->bh + (1.0 - alpha) * (double)c->trk_ema_bh);
    }
    c->trk_steps_since += 1;
    c->trk_prev_loss = loss;
    c->trk_have_prev = 1;
    const double gd = (k > 0) ? dot : 0.0;  // check_gradient_oscillation: zeros before the first step
    if (gd < 0.0) c->trk_steps_since = 0;
    if (loss < (double)a.trk_tolerance || isnan(loss) || isinf(loss)) {
      c->stopping = 1;
      c->trk_reason = 2;
      c->trk_stop_step = k;
    } else {
      const double rel = (c->trk_have_ema && loss > 0.0) ? c->trk_ema / loss : 0.0;
      const int ti = c->trk_idx < a.trk_n_thr ? c->trk_idx : a.trk_n_thr - 1;
      const bool met = c->trk_steps_since >= a.trk_min_steps && c->trk_have_ema && rel < (double)c->thr[ti] && k > cfg.initial_steps;
      if (k == 0 || met) {
        c->snap_now = c->trk_nsnap;
        c->trk_nsnap += 1;
        if (k > 0) {
          c->trk_idx += 1;
          c->trk_steps_since = 0;
          if (c->trk_idx >= a.trk_n_thr) {
            c->stopping = 1;
            c->trk_reason = 1;
            c->trk_stop_step = k;
          }
        }
      }
    }
    if (c->snap_now >= 0 && a.snaps != nullptr) {
      JaxentTrackSnapshot* sn = a.snaps + (size_t)c->snap_now * a.Bn + b;
      sn->loop_step = k;
      sn->state_step = c->step;
      sn->have_ema = c->trk_have_ema;
      sn->threshold_idx = c->trk_idx;
      sn->bc = c->bc;
      sn->bh = c->bh;
      sn->ema_bc = c->trk_ema_bc;
      sn->ema_bh = c->trk_ema_bh;
      sn->losses = a.records[(size_t)(k % RB) * a.Bn + b];
    }
  }
}

// ------------------------------------------------------------------------------------------ bt_adam
// thread = (ONE replica b, chunk lane cy); a chunk is 8 consecutive frames: z, m, v updated in place, e' written, and the bf16
// pieces of e' stored as 16-byte K-chunks of the gemm2 B operand.  All 32 loads of a chunk (4 arrays x 8 frames, each a
// coalesced 128-byte line per warp: consecutive lanes are consecutive replicas) are issued before the arithmetic; ~64
// registers per thread, four 256-thread blocks per SM.  (The first version gave a thread 4 replicas x 8 frames with 16-byte
// accesses: 128 registers, one block per SM, 197 us for 870 MB.)  update = 0: initial forward (no Adam).
__global__ void __launch_bounds__(512, 2) bt_adam_kernel(const BatchArgs a, int update) {
  extern __shared__ double sh_d[];  // [2][ncy][Bn]
  const int Bn = a.Bn, b = threadIdx.x % Bn, cy = threadIdx.x / Bn, ncy = blockDim.x / Bn;
  const float* __restrict__ gp = a.g[*a.gsel];
  float* __restrict__ zp_ = a.z;
  float* __restrict__ mp = a.m;
  float* __restrict__ vp = a.v;
  float* __restrict__ ep = a.e;
  const BCtl* c = a.ctl + b;
  const float lr = c->lr_eff, lse = c->lse;
  const float ibc1 = 1.0f / c->bc1, ibc2 = 1.0f / c->bc2;  // bias corrections applied as reciprocal multiplies
  const bool active = update && !c->stop;
  const bool has_kl = c->kl_slot >= 0 && a.prob.prior != nullptr;
  const float b1 = (float)a.cfg.b1, b2 = (float)a.cfg.b2, eps = (float)a.cfg.eps, clip = a.cfg.clip_value;
  const float om_b1 = (float)(1.0 - a.cfg.b1), om_b2 = (float)(1.0 - a.cfg.b2);
  const int NS = a.NS;
  const long long n_chunks = (long long)a.n_fb * 16;
  double sum = 0.0, zp = 0.0;  // sum_f e'_f and sum_f p_f z'_f (MaxEnt loss)
  for (long long ch = (long long)blockIdx.x * ncy + cy; ch < n_chunks; ch += (long long)gridDim.x * ncy) {
    const long long f0 = ch * 8;
    __nv_bfloat16 p0[8], p1[8], p2[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // two halves of 4 frames: 16 loads in flight per thread, 64 registers
      float z_[4], m_[4], v_[4], g_[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long f = f0 + 4 * h + k;
        const bool ok = f < a.F;
        const size_t i = (size_t)(ok ? f : 0) * Bn + b;
        z_[k] = ok ? zp_[i] : 0.f;
        if (update) {
          g_[k] = ok ? gp[i] : 0.f;
          m_[k] = ok ? mp[i] : 0.f;
          v_[k] = ok ? vp[i] : 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long f = f0 + 4 * h + k;
        const size_t i = (size_t)f * Bn + b;
        float en = 0.f;
        if (f < a.F) {
          if (update) {
            float mo = m_[k], vo = v_[k];
            if (active) {
              float sg = g_[k] * lr;
              if (clip > 0.f) sg = fminf(fmaxf(sg, -clip), clip);
              mo = om_b1 * sg + b1 * m_[k];
              vo = om_b2 * (sg * sg) + b2 * v_[k];
              z_[k] = z_[k] - (mo * ibc1) / (sqrtf(vo * ibc2) + eps);
            }
            zp_[i] = z_[k];
            mp[i] = mo;
            vp[i] = vo;
          }
          en = __expf(z_[k] - lse);
          sum += (double)en;
          if (has_kl) zp += (double)a.pnorm[f] * (double)z_[k];
          ep[i] = en;
        }
        split3(en, NS, p0[4 * h + k], p1[4 * h + k], p2[4 * h + k]);
      }
    }
    const long long fb = ch >> 4;
    const int kc = (int)(ch & 15);
    uint4* dst = reinterpret_cast<uint4*>(a.e3);
    dst[(((size_t)fb * NS + 0) * 16 + kc) * Bn + b] = *reinterpret_cast<const uint4*>(p0);
    if (NS > 1) dst[(((size_t)fb * NS + 1) * 16 + kc) * Bn + b] = *reinterpret_cast<const uint4*>(p1);
    if (NS > 2) dst[(((size_t)fb * NS + 2) * 16 + kc) * Bn + b] = *reinterpret_cast<const uint4*>(p2);
  }
  sh_d[cy * Bn + b] = sum;
  sh_d[(ncy + cy) * Bn + b] = zp;
  __syncthreads();
  for (int bb = threadIdx.x; bb < Bn; bb += blockDim.x) {
    double t = 0.0, tz = 0.0;
    for (int k = 0; k < ncy; ++k) {
      t += sh_d[k * Bn + bb];
      tz += sh_d[(ncy + k) * Bn + bb];
    }
    a.part_s[(size_t)blockIdx.x * Bn + bb] = t;
    a.part_s[(size_t)(MAX_EW_BLOCKS + blockIdx.x) * Bn + bb] = tz;
  }
}

// ------------------------------------------------------------------------------------------ bt_reduce2
// red[b][row] = sum_ks P[ks][row][b]  (32 x 32 transposing tiles);  block row 0 also folds S_b
__global__ void __launch_bounds__(1024) bt_reduce2_kernel(const BatchArgs a) {
  __shared__ double tile[32][33];
  const int Bn = a.Bn, R2 = 2 * a.Rp;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int row0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  {
    const int row = row0 + ty, b = b0 + tx;
    double s = 0.0;
    if (row < R2 && b < Bn) {
#pragma unroll 8
      for (int ks = 0; ks < a.ksplit; ++ks) s += (double)__ldcg(&a.P[((size_t)ks * R2 + row) * Bn + b]);
    }
    tile[ty][tx] = s;
  }
  __syncthreads();
  {
    const int b = b0 + ty, row = row0 + tx;
    if (row < R2 && b < Bn) a.red[(size_t)b * R2 + row] = tile[tx][ty];
  }
  if (blockIdx.x == 0) {  // S_b = sum_f e' and sum_f p_f z_f: block partials summed ty-strided, then over ty (fixed order)
    const int b = b0 + tx;
    for (int which = 0; which < 2; ++which) {
      __syncthreads();
      double s = 0.0;
      if (b < Bn)
        for (int k = ty; k < a.adam_blocks; k += 32) s += a.part_s[(size_t)(which * MAX_EW_BLOCKS + k) * Bn + b];
      tile[ty][tx] = s;
      __syncthreads();
      if (ty == 0 && b < Bn) {
        double t = 0.0;
        for (int k = 0; k < 32; ++k) t += tile[k][tx];
        if (which == 0) a.ctl[b].S = t;
        else a.ctl[b].zp = t;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ bt_tail
template <bool GENERAL>
__global__ void __launch_bounds__(TAIL_THREADS, 1) bt_tail_kernel(const BatchArgs a) {
  unsigned char* const smem_raw = jx_dyn_smem;
  __shared__ double s_red[32][16];
  __shared__ double s_fin[16];
  __shared__ __align__(16) BCtl s_c;
  const JaxentBvProblem& pb = a.prob;
  const int R = pb.R, T = pb.uptake ? pb.T : 0, Rp = a.Rp, Bn = a.Bn;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const TailSmem tsm = carve_tail_smem(smem_raw, R, T, a.p_max_tr, a.p_max_va);
  // the body's smem is followed by the (cc, ch) staging rows of this replica
  float* s_cc = reinterpret_cast<float*>(tsm.s_blob + ((a.blob_words_max + 3) & ~3));
  float* s_ch = s_cc + Rp;

  const int first_data = stage_tail_constants(pb, a.blobs, tsm);
  if (tid < (int)(sizeof(BCtl) / 16)) reinterpret_cast<int4*>(&s_c)[tid] = reinterpret_cast<const int4*>(a.ctl + b)[tid];
  double acc0 = 0.0, acc1 = 0.0;
  if (tid < R) {
    acc0 = a.red[(size_t)b * 2 * Rp + tid];
    acc1 = a.red[(size_t)b * 2 * Rp + Rp + tid];
  }
  const double off0 = (a.row_off && tid < R) ? (double)a.row_off[tid] : 0.0, off1 = (a.row_off && tid < R) ? (double)a.row_off[Rp + tid] : 0.0;
  if (lane < 16) s_red[warp][lane] = 0.0;
  for (int i = tid; i < 2 * Rp; i += blockDim.x) s_cc[i] = 0.f;
  __syncthreads();
  if (s_c.stop) return;
  const double S = s_c.S;
  // the GEMM swept centred features: sum_f e'_f (N[r,f] - off_r) = acc - off_r S
  acc0 += off0 * S;
  acc1 += off1 * S;
  const double bc = (double)s_c.bc, bh = (double)s_c.bh;
  double my_c, my_a, my_b, my_lp, gbc_reg, gbh_reg;
  TailIO io;
  io.row_off = a.row_off;
  io.row_off_stride = Rp;
  io.avg = a.avg + (size_t)b * 2 * R;
  io.logpf = a.logpf + (size_t)b * R;
  io.uptake = a.uptake + (size_t)b * (T > 0 ? T : 1) * R;
  io.pf_G = nullptr;
  io.pf_mode = 0;
  io.p_max_tr = a.p_max_tr;
  io.p_max_va = a.p_max_va;
  io.c_pieces = a.NS;
  io.prof = nullptr;
  io.ck = nullptr;
  tail_body<GENERAL>(pb, a.blobs, first_data, tsm, acc0, acc1, S, bc, bh, s_c.fw, s_c.fs, io, s_red, s_fin, my_c, my_a, my_b, my_lp, gbc_reg,
                     gbh_reg);
  if (tid < R) {
    s_cc[tid] = (float)(my_c * bc);
    s_ch[tid] = (float)(my_c * bh);
  }
  __syncthreads();
  // c -> bf16 pieces, one 16-byte K-chunk (8 residues) per thread and piece
  {
    const int n_kc = Rp / 8;
    for (int it = tid; it < 2 * n_kc; it += blockDim.x) {
      const int arr = it / n_kc, kc = it - arr * n_kc;
      const float* src = (arr ? s_ch : s_cc) + kc * 8;
      __nv_bfloat16 p0[8], p1[8], p2[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) split3(src[k], a.NS, p0[k], p1[k], p2[k]);
      uint4* dst = reinterpret_cast<uint4*>(a.cc3);
      dst[(((size_t)0 * 2 + arr) * n_kc + kc) * Bn + b] = *reinterpret_cast<const uint4*>(p0);
      if (a.NS > 1) dst[(((size_t)1 * 2 + arr) * n_kc + kc) * Bn + b] = *reinterpret_cast<const uint4*>(p1);
      if (a.NS > 2) dst[(((size_t)2 * 2 + arr) * n_kc + kc) * Bn + b] = *reinterpret_cast<const uint4*>(p2);
    }
  }
  // control block of the coming update + loss record
  if (tid == 0) {
    BCtl* c = a.ctl + b;
    const JaxentOptConfig& cfg = a.cfg;
    const int step = s_c.step;
    const bool switched = step >= cfg.initial_steps;
    const int mi = switched ? 1 : s_c.mask_idx;
    const float base_lr = switched ? s_c.cfg_lr : s_c.lr;
    const float base_mlr = switched ? s_c.cfg_lr * cfg.model_parameters_lr_scale : s_c.model_lr;
    const float mask_model = (mi == 0) ? 0.f : (cfg.optimise_model_parameters ? 1.f : 0.f);
    c->lrA = base_lr;
    c->lrB = base_lr / cfg.plateau_denominator;
    c->mlrA = base_mlr;
    c->mlrB = base_mlr / cfg.plateau_denominator;
    c->mask_frame = (mi == 0) ? 1.f : (cfg.optimise_frame_weights ? 1.f : 0.f);
    c->mask_model = mask_model;
    const double lse_d = (double)s_c.lse + dlog(S);
    c->lse_d = lse_d;
    c->lse = (float)lse_d;
    c->hbar_data = s_fin[12];
    const float gb0 = (float)(s_fin[13] + gbc_reg) * mask_model, gb1 = (float)(s_fin[14] + gbh_reg) * mask_model;
    c->gb[0] = gb0;
    c->gb[1] = gb1;
    JaxentLossRecord* rec = a.records + (size_t)(step % RB) * Bn + b;
    rec->step = step;
    rec->branch = s_c.last_branch;
    rec->lr = s_c.lr;
    rec->grad_dot = s_c.last_dot;
    rec->bc = s_c.bc;
    rec->bh = s_c.bh;
    rec->grad_bc = gb0;
    rec->grad_bh = gb1;
    rec->complete = 0;
    float tt = 0.f, tv = 0.f;
    for (int i = 0; i < JAXENT_MAX_SLOTS; ++i) {
      const bool on = i < pb.n_slots;
      rec->train[i] = on ? (float)s_fin[i] : 0.f;
      rec->val[i] = on ? (float)s_fin[JAXENT_MAX_SLOTS + i] : 0.f;
      rec->scaled_train[i] = rec->train[i] * s_c.fw[i] * s_c.fs[i];
      rec->scaled_val[i] = rec->val[i] * s_c.fw[i] * s_c.fs[i];
      tt += rec->scaled_train[i];
      tv += rec->scaled_val[i];
    }
    rec->total_train = tt;
    rec->total_val = tv;
    if (s_c.kl_slot < 0) rec->complete = 1;
  }
}

// ------------------------------------------------------------------------------------------ bt_kl
// MaxEnt terms per (frame, replica); the last block folds the per-block sums (fixed order) into the control blocks
// and completes the loss records of this step
__global__ void __launch_bounds__(512) bt_kl_kernel(const BatchArgs a) {
  extern __shared__ double sh_d[];  // [fy][KL_N][Bn]
  const int Bn = a.Bn, Bq = Bn >> 2, bq = threadIdx.x % Bq, fy = threadIdx.x / Bq, b0 = 4 * bq;
  // maxent_convexKL_loss = sum_f p_f (log p_f - log q_f) + sum_f (q_f - p_f),  q_f = w_f + eps,  w_f = e_f / S = exp(z_f - lse_new):
  //   sum_f p_f log q_f = sum_f p_f z_f - lse_new * sum_f p_f + sum_f p_f log1p(eps / w_f)
  // sum p z comes from bt_adam (fp64), sum p log p and sum p are constants of the prior (a.kl_const), and log1p(eps / w)
  // only matters for frames whose weight has collapsed to the order of eps = 1e-10 / F: those are evaluated here,
  // every other frame contributes less than 1e-9 * p_f and is skipped.  No logarithm per (frame, replica).
  // thread = (4 consecutive replicas, frame lane): 16-byte loads, 4 frames in flight.
  __shared__ float s_f[3][256];
  __shared__ double s_invS[256];
  __shared__ int s_i[2][256];
  for (int b = threadIdx.x; b < Bn; b += blockDim.x) {  // per-replica constants through shared memory (see bt_grad)
    const BCtl* c = a.ctl + b;
    s_invS[b] = 1.0 / c->S;
    s_f[0][b] = (float)(1.0 / c->S);
    s_f[1][b] = c->lse;
    s_f[2][b] = c->kl_scale;
    s_i[0][b] = ((c->kl_slot >= 0 && a.prob.prior != nullptr) ? 1 : 0) | (c->stop ? 0 : 2) | ((a.trk_on && c->ema_update) ? 4 : 0) | (c->ema_first ? 8 : 0);
    s_i[1][b] = a.trk_on ? c->snap_now : -1;
  }
  __syncthreads();
  float invS[4], lse[4], lam[4];
  bool has_kl[4], run[4];
  int snap[4], ema_update[4], ema_first[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    invS[j] = s_f[0][b0 + j];
    lse[j] = s_f[1][b0 + j];
    lam[j] = s_f[2][b0 + j];
    const int fl = s_i[0][b0 + j];
    has_kl[j] = fl & 1;
    run[j] = fl & 2;
    ema_update[j] = (fl & 4) ? 1 : 0;
    ema_first[j] = (fl & 8) ? 1 : 0;
    snap[j] = s_i[1][b0 + j];
  }
  double s0[4] = {0.0, 0.0, 0.0, 0.0}, s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0}, s3[4] = {0.0, 0.0, 0.0, 0.0};
  const float4* __restrict__ ep = reinterpret_cast<const float4*>(a.e);
  const float* __restrict__ pn = a.pnorm;
  const long long fstep = (long long)gridDim.x * a.fy;
  for (long long f0 = (long long)blockIdx.x * a.fy + fy; f0 < a.F; f0 += 4 * fstep) {
    float4 ev[4];
    float pv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long fk = f0 + k * fstep;
      const long long fc = fk < a.F ? fk : f0;
      ev[k] = ep[(size_t)fc * Bq + bq];
      pv[k] = a.prob.prior != nullptr ? pn[fc] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long f = f0 + k * fstep;
      if (f >= a.F) break;
      const float e4[4] = {ev[k].x, ev[k].y, ev[k].z, ev[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!run[j]) continue;
        const float p = pv[k];
        float w, q;
        double kap;
        weight_terms(e4[j], invS[j], p, has_kl[j], a.eps_kl, lam[j], w, kap, q);  // the same w, kappa bt_grad will form
        if (has_kl[j]) {
          s0[j] += ((double)e4[j] * s_invS[b0 + j]) * kap;  // MaxEnt part of hbar with e_f / S in fp64 (see frame_terms, bv_tail.cu)
          s2[j] += (double)(q - p);
          const float x = a.eps_kl * rcp_approx(fmaxf(w, 1e-38f));
          if (x > 1e-9f) {
            // log(w + eps) - log(w) with log w = z - lse taken from the logits: a weight this small may be a denormal fp32
            // number (2 bits of precision at exp(-100)), so w itself must not enter the logarithm
            const float lw = a.z[(size_t)f * Bn + b0 + j] - lse[j];
            const float le = logf(a.eps_kl);
            s1[j] += (double)p * (double)((le + log1pf(expf(lw - le))) - lw);
          }
        }
        s3[j] += (double)w;
        if (a.trk_on) {  // tracker.ema_params (opt/track.py:160-165) and the history snapshot of the post-update weights
          const size_t i = (size_t)f * Bn + b0 + j;
          if (ema_update[j]) {
            const float em = ema_first[j] ? w : a.trk_alpha * w + (1.0f - a.trk_alpha) * a.ema_w[i];
            a.ema_w[i] = em;
            if (snap[j] >= 0 && !a.trk_pure) a.snap_ema[(size_t)snap[j] * a.F * Bn + i] = em;
          }
          if (snap[j] >= 0) a.snap_w[(size_t)snap[j] * a.F * Bn + i] = w;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double* mine = sh_d + ((size_t)fy * KL_N) * Bn + b0 + j;
    mine[0 * Bn] = s0[j];
    mine[1 * Bn] = s1[j];
    mine[2 * Bn] = s2[j];
    mine[3 * Bn] = s3[j];
  }
  __syncthreads();
  for (int b = threadIdx.x; b < Bn; b += blockDim.x) {
    for (int k = 0; k < KL_N; ++k) {
      double t = 0.0;
      for (int y = 0; y < a.fy; ++y) t += sh_d[((size_t)y * KL_N + k) * Bn + b];
      a.part_kl[((size_t)blockIdx.x * KL_N + k) * Bn + b] = t;
    }
  }
}

// fold of the per-block MaxEnt sums, one warp per replica (block index lane-strided, then a shuffle reduction: fixed order),
// and completion of this step's loss record
__global__ void __launch_bounds__(256) bt_klfold_kernel(const BatchArgs a, int n_blocks) {
  const int Bn = a.Bn;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= Bn) return;
  {
    BCtl* cw = a.ctl + b;
    if (cw->stop) return;
    double kl[KL_N];
    for (int k = 0; k < KL_N; ++k) {
      double t = 0.0;
      for (int blk0 = lane; blk0 < n_blocks; blk0 += 32 * 8) {  // 8 independent loads in flight per lane
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int blk = blk0 + 32 * u;
          v[u] = blk < n_blocks ? __ldcg(&a.part_kl[((size_t)blk * KL_N + k) * Bn + b]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) t += v[u];
      }
      kl[k] = wsum(t);
    }
    if (lane != 0) return;
    if (cw->stopping) {  // this was the last evaluation of the replica: freeze it
      cw->stop = 1;
      cw->stopping = 0;
    }
    cw->snap_now = -1;
    cw->ema_update = 0;
    for (int k = 0; k < KL_N; ++k) cw->klsum[k] = kl[k];
    if (cw->kl_slot >= 0) {
      JaxentLossRecord* rec = a.records + (size_t)(cw->step % RB) * Bn + b;
      // kl[1] holds the rare-frame corrections sum p log1p(eps / w)
      // log w_f = z_f - (lse_old + log S) with the sum in double: the fp32 lse is 5e-7 off at lse ~ 11, which the loss scaling (x 1000) turns into 2e-4
      const double plogq = cw->zp - cw->lse_d * a.kl_const[1] + kl[1];
      const float klv = (float)((a.kl_const[0] - plogq) + kl[2]);
      const int i = cw->kl_slot;
      rec->train[i] = klv;
      rec->val[i] = klv;
      float tt = 0.f, tv = 0.f;
      for (int k = 0; k < a.prob.n_slots; ++k) {
        rec->scaled_train[k] = rec->train[k] * cw->fw[k] * cw->fs[k];
        rec->scaled_val[k] = rec->val[k] * cw->fw[k] * cw->fs[k];
        tt += rec->scaled_train[k];
        tv += rec->scaled_val[k];
      }
      rec->total_train = tt;
      rec->total_val = tv;
      rec->complete = 1;
    }
  }
}

// normalised prior and its two constants (one block, fixed order)
__global__ void __launch_bounds__(1024) bt_prior_kernel(const BatchArgs a) {
  __shared__ double sh[2][32];
  const double inv_pnorm = 1.0 / a.prob.prior_norm;
  double s_plogp = 0.0, s_p = 0.0;
  for (long long f = threadIdx.x; f < a.F; f += blockDim.x) {
    const float p = (float)(((double)fabsf(a.prob.prior[f]) + (double)a.eps_kl) * inv_pnorm);
    a.pnorm[f] = p;
    s_p += (double)p;
    if (p > 0.f) s_plogp += (double)p * log((double)p);
  }
  s_plogp = wsum(s_plogp);
  s_p = wsum(s_p);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s_plogp; sh[1][threadIdx.x >> 5] = s_p; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sh[threadIdx.x][i];
    a.kl_const[threadIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------ init
__global__ void bt_init_kernel(const BatchArgs a, const float* __restrict__ z0, const BCtl* __restrict__ init) {

  const long long n = (long long)a.F * a.Bn;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i < n; i += stride) {
    const long long f = i / a.Bn;
    a.z[i] = z0[f];
    a.m[i] = 0.f;
    a.v[i] = 0.f;
    a.g[0][i] = 0.f;
    a.g[1][i] = 0.f;
  }
  if (i0 < a.Bn) a.ctl[i0] = init[i0];
  if (i0 == 0) {
    *a.gsel = 0u;
    *a.counter = 0u;
  }
  const long long nrec = (long long)RB * a.Bn * (long long)(sizeof(JaxentLossRecord) / 4);
  for (long long i = i0; i < nrec; i += stride) reinterpret_cast<int*>(a.records)[i] = 0;
}

}  // namespace bg
}  // namespace jx

// ============================================================================================== host side
using namespace jx;

struct JaxentBatchPlan {
  bg::BatchArgs a;
  float conv[JAXENT_TRACK_MAX_THRESHOLDS];
  unsigned char* ws;
  size_t ws_bytes;
  size_t off_ctl_init;
  size_t tail_smem;
  int general;
  int sm_count;
  int fused_grad;  // phase 0 = bt_fused_grad_kernel (Bn a multiple of 128)
  cudaStream_t side;        // jaxent_bg_steps: the MaxEnt terms of step k run here, beside the backward product of step k + 1
  cudaEvent_t ev_tail, ev_kl;
  int initialised;
  long long steps;
};

namespace {
struct BLayout {
  size_t z, m, v, e, g0, g1, pnorm, kl_const, H, P, cc3, e3, red, part_dot, part_s, part_kl, ctl, ctl_init, records, avg, logpf, uptake, gsel, counter;
  size_t blob[JAXENT_MAX_SLOTS];
  int blob_words[JAXENT_MAX_SLOTS];
  size_t total;
};
size_t up(size_t x) { return (x + 255) / 256 * 256; }

BLayout b_layout(const JaxentBvProblem* pb, int Bn, int NS, int ksplit) {
  const size_t F = (size_t)pb->F, R = pb->R, T = pb->uptake ? pb->T : 1;
  const size_t Rp = (R + 127) / 128 * 128, Fp = (F + 127) / 128 * 128;
  BLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = up(o + bytes); return r; };
  const size_t fb = F * Bn * 4;
  L.z = take(fb); L.m = take(fb); L.v = take(fb); L.e = take(fb); L.g0 = take(fb); L.g1 = take(fb);
  L.pnorm = take(F * 4);
  L.H = take(Fp * Bn * 4);
  L.P = take((size_t)ksplit * 2 * Rp * Bn * 4);
  L.cc3 = take((size_t)NS * 2 * Rp * Bn * 2);
  L.e3 = take(Fp * NS * Bn * 2);
  L.red = take((size_t)Bn * 2 * Rp * 8);
  L.part_dot = take((size_t)bg::MAX_EW_BLOCKS * Bn * 8);
  L.part_s = take((size_t)2 * bg::MAX_EW_BLOCKS * Bn * 8);
  L.kl_const = take(2 * 8);
  L.part_kl = take((size_t)bg::MAX_EW_BLOCKS * KL_N * Bn * 8);
  L.ctl = take((size_t)Bn * sizeof(bg::BCtl));
  L.ctl_init = take((size_t)Bn * sizeof(bg::BCtl));
  L.records = take((size_t)bg::RB * Bn * sizeof(JaxentLossRecord));
  L.avg = take((size_t)Bn * 2 * R * 4);
  L.logpf = take((size_t)Bn * R * 4);
  L.uptake = take((size_t)Bn * T * R * 4);
  L.gsel = take(4);
  L.counter = take(4);
  for (int i = 0; i < JAXENT_MAX_SLOTS; ++i) {
    L.blob[i] = 0;
    L.blob_words[i] = 0;
    if (i < pb->n_slots && pb->slots[i].kind <= JAXENT_LOSS_PF_L2) {
      const JaxentLossSlot& sl = pb->slots[i];
      L.blob_words[i] = blob_words(pb->R, pb->uptake ? pb->T : 0, sl.train.n_pep, sl.train.nnz, sl.val.n_pep, sl.val.nnz);
      L.blob[i] = take((size_t)L.blob_words[i] * 4);
    }
  }
  L.total = o;
  return L;
}

int b_check(const JaxentBvProblem* pb, int n_replicas, int NS) {
  if (!pb || pb->R <= 0 || pb->F <= 0) { set_error("batched plan: invalid problem"); return JAXENT_E_INVALID; }
  if (n_replicas < 1 || n_replicas > 256 || n_replicas % 16 != 0) { set_error("batched plan: replicas must be a multiple of 16 in 16..256 (pad by repeating the last)"); return JAXENT_E_INVALID; }
  if (NS < 1 || NS > 3) { set_error("batched plan: 1..3 bf16 pieces"); return JAXENT_E_INVALID; }
  if (pb->R > TAIL_THREADS) { set_error("batched plan: n_residues=%d exceeds %d", pb->R, TAIL_THREADS); return JAXENT_E_UNSUPPORTED; }
  if (pb->feature_format != 2) { set_error("batched plan: features must be in the bf16 GEMM layout (jaxent_bg_pack_features, feature_format = 2)"); return JAXENT_E_INVALID; }
  return JAXENT_OK;
}
int b_ksplit(const JaxentBvProblem* pb, int sm) {
  const int Rp = (pb->R + 127) / 128 * 128, n_mt = 2 * Rp / 128;
  const long long n_hb = 2 * ((pb->F + 127) / 128);
  long long ks = sm / n_mt;
  if (ks < 1) ks = 1;
  if (ks > n_hb) ks = n_hb;
  return (int)ks;
}
}  // namespace

extern "C" size_t jaxent_bg_workspace_bytes(const JaxentBvProblem* pb, int32_t n_replicas, int32_t n_pieces, int32_t sm_count) {
  if (b_check(pb, n_replicas, n_pieces)) return 0;
  if (sm_count <= 0) sm_count = 148;
  return b_layout(pb, n_replicas, n_pieces, b_ksplit(pb, sm_count)).total;
}

extern "C" int jaxent_bg_plan_create(const JaxentBvProblem* pb, const JaxentOptConfig* cfg, int32_t n_replicas, int32_t n_pieces,
                                     void* workspace, size_t workspace_bytes, int32_t sm_count, JaxentBatchPlan** out) {
  if (!out) { set_error("batched plan: NULL output"); return JAXENT_E_INVALID; }
  *out = nullptr;
  int rc = b_check(pb, n_replicas, n_pieces);
  if (rc) return rc;
  if (!cfg || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) { set_error("batched plan: config / 256-byte aligned workspace required"); return JAXENT_E_INVALID; }
  if (sm_count <= 0) sm_count = 148;
  const int ksplit = b_ksplit(pb, sm_count);
  const BLayout L = b_layout(pb, n_replicas, n_pieces, ksplit);
  if (workspace_bytes < L.total) { set_error("batched plan: workspace too small: %zu < %zu", workspace_bytes, L.total); return JAXENT_E_WORKSPACE; }
  int n_kl = 0;
  bool general = false;
  for (int i = 0; i < pb->n_slots; ++i) {
    const int k = pb->slots[i].kind;
    if (k == JAXENT_LOSS_MAXENT_CONVEX_KL) {
      if (!pb->prior || !(pb->prior_norm > 0.0)) { set_error("batched plan: maxent_convexKL_loss needs prior and prior_norm"); return JAXENT_E_INVALID; }
      ++n_kl;
    }
    if (k < 0 || k > JAXENT_LOSS_PARAMS_L2) { set_error("batched plan: unsupported loss kind %d", k); return JAXENT_E_UNSUPPORTED; }
    if ((k <= JAXENT_LOSS_UPTAKE_SIGMA_MSE) != (pb->uptake != 0) && k <= JAXENT_LOSS_PF_L2) { set_error("batched plan: slot %d does not match the forward model", i); return JAXENT_E_UNSUPPORTED; }
    general = general || k == JAXENT_LOSS_UPTAKE_MC_EYE_MSE || k == JAXENT_LOSS_UPTAKE_SIGMA_MSE;
  }
  if (n_kl > 1) { set_error("batched plan: at most one maxent_convexKL_loss slot"); return JAXENT_E_UNSUPPORTED; }
  JaxentBatchPlan* p = new (std::nothrow) JaxentBatchPlan();
  if (!p) { set_error("out of host memory"); return JAXENT_E_INVALID; }
  memset(p, 0, sizeof(*p));
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  p->ws = ws;
  p->ws_bytes = L.total;
  bg::BatchArgs& a = p->a;
  a.prob = *pb;
  a.cfg = *cfg;
  a.Bn = n_replicas;
  a.NS = n_pieces;
  a.Rp = (pb->R + 127) / 128 * 128;
  a.n_rg = a.Rp / 8;
  a.n_fb = (int)((pb->F + 127) / 128);
  a.ksplit = ksplit;
  a.F = pb->F;
  a.z = (float*)(ws + L.z); a.m = (float*)(ws + L.m); a.v = (float*)(ws + L.v); a.e = (float*)(ws + L.e);
  a.g[0] = (float*)(ws + L.g0); a.g[1] = (float*)(ws + L.g1);
  a.pnorm = (float*)(ws + L.pnorm);
  a.kl_const = (double*)(ws + L.kl_const);
  a.H = (float*)(ws + L.H);
  a.P = (float*)(ws + L.P);
  a.cc3 = ws + L.cc3;
  a.e3 = ws + L.e3;
  a.red = (double*)(ws + L.red);
  a.part_dot = (double*)(ws + L.part_dot);
  a.part_s = (double*)(ws + L.part_s);
  a.part_kl = (double*)(ws + L.part_kl);
  a.ctl = (bg::BCtl*)(ws + L.ctl);
  a.records = (JaxentLossRecord*)(ws + L.records);
  a.avg = (float*)(ws + L.avg);
  a.logpf = (float*)(ws + L.logpf);
  a.uptake = (float*)(ws + L.uptake);
  a.gsel = (unsigned*)(ws + L.gsel);
  a.counter = (unsigned*)(ws + L.counter);
  p->off_ctl_init = L.ctl_init;
  for (int i = 0; i < JAXENT_MAX_SLOTS; ++i) {
    a.blobs.words[i] = L.blob_words[i];
    a.blobs.blob[i] = L.blob_words[i] ? (int*)(ws + L.blob[i]) : nullptr;
  }
  int p_tr = 1, p_va = 1;
  for (int i = 0; i < pb->n_slots; ++i)
    if (pb->slots[i].kind <= JAXENT_LOSS_PF_L2) {
      if (pb->slots[i].train.n_pep > p_tr) p_tr = pb->slots[i].train.n_pep;
      if (pb->slots[i].val.n_pep > p_va) p_va = pb->slots[i].val.n_pep;
      if (pb->slots[i].train.n_pep <= 0) { delete p; set_error("batched plan: slot %d has an empty training set", i); return JAXENT_E_INVALID; }
    }
  a.p_max_tr = p_tr;
  a.p_max_va = p_va;
  a.eps_kl = (float)(1e-10 / (double)pb->F_total);
  // elementwise geometry: blocks of Bn x fy threads
  a.fy = 512 / (n_replicas / 4);  // elementwise kernels: Bn/4 threads (4 replicas each) per frame, fy frames per 512-thread block
  long long blocks = ((long long)pb->F + a.fy - 1) / a.fy;
  const long long want = (long long)sm_count * 2;
  if (blocks > want) blocks = want;
  if (blocks > bg::MAX_EW_BLOCKS) blocks = bg::MAX_EW_BLOCKS;
  if (blocks < 1) blocks = 1;
  a.ew_blocks = (int)blocks;
  // bt_adam: one replica per thread, blocks of adam_threads = Bn x (chunks per block), four 256-thread blocks per SM
  a.adam_threads = n_replicas * (n_replicas >= 256 ? 1 : 256 / n_replicas);
  {
    const long long n_chunks = (long long)a.n_fb * 16, per_block = a.adam_threads / n_replicas;
    long long ab = (n_chunks + per_block - 1) / per_block;
    const long long cap = (long long)sm_count * 4 < bg::MAX_EW_BLOCKS ? (long long)sm_count * 4 : bg::MAX_EW_BLOCKS;
    a.adam_blocks = (int)(ab < cap ? ab : cap);
  }
  // backward product: gemm1 + bt_grad; JAXENT_BG_FUSED=1 selects the single fused tensor-core kernel (whole 128-replica M tiles only)
  {
    const char* env = getenv("JAXENT_BG_FUSED");
    p->fused_grad = (n_replicas % 128 == 0) && (env && env[0] == '1');  // opt-in: measured slower than gemm1 + bt_grad so far
    a.dot_blocks = a.ew_blocks;
    if (p->fused_grad) {
      const int halves = n_replicas / 128;
      const int n_tiles = ((a.n_fb + 1) / 2) * halves;
      int grid = n_tiles < sm_count ? n_tiles : sm_count;
      if (halves == 2) grid &= ~1;  // a CTA keeps one replica half
      if (grid < halves) grid = halves;
      a.dot_blocks = grid;
    }
  }
  {
    int nnz_tr = 1, nnz_va = 1;
    for (int i = 0; i < pb->n_slots; ++i)
      if (pb->slots[i].kind <= JAXENT_LOSS_PF_L2) {
        if (pb->slots[i].train.nnz > nnz_tr) nnz_tr = pb->slots[i].train.nnz;
        if (pb->slots[i].val.nnz > nnz_va) nnz_va = pb->slots[i].val.nnz;
      }
    const int T = pb->uptake ? pb->T : 0;
    a.blob_words_max = blob_words(pb->R, T, p_tr, nnz_tr, p_va, nnz_va);
    p->tail_smem = tail_smem_bytes(pb->R, T, p_tr, p_va, nnz_tr, nnz_va) + (size_t)2 * a.Rp * 4 + 64;
  }
  if (p->tail_smem > 220 * 1024) { delete p; set_error("batched plan: tail needs %zu bytes of shared memory", p->tail_smem); return JAXENT_E_UNSUPPORTED; }
  p->general = general;
  p->sm_count = sm_count;
  p->side = nullptr;
  p->ev_tail = p->ev_kl = nullptr;
  {  // side stream + fork / join events of jaxent_bg_steps (set-up time; JAXENT_BG_OVERLAP=0 keeps a step on one stream)
    const char* env = getenv("JAXENT_BG_OVERLAP");
    if (!(env && env[0] == '0')) {
      cudaError_t e = cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_tail, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_kl, cudaEventDisableTiming);
      if (e != cudaSuccess) {  // no device (CPU-side tests build plans without one) or out of resources: single-stream steps
        (void)cudaGetLastError();
        if (p->side) cudaStreamDestroy(p->side);
        if (p->ev_tail) cudaEventDestroy(p->ev_tail);
        p->side = nullptr;
        p->ev_tail = p->ev_kl = nullptr;
      }
    }
  }
  *out = p;
  return JAXENT_OK;
}

extern "C" void jaxent_bg_plan_destroy(JaxentBatchPlan* p) {
  if (!p) return;
  jaxent_handle_release_ptr(p);
  if (p->side) cudaStreamDestroy(p->side);
  if (p->ev_tail) cudaEventDestroy(p->ev_tail);
  if (p->ev_kl) cudaEventDestroy(p->ev_kl);
  delete p;
}

extern "C" int jaxent_bg_plan_set_row_offsets(JaxentBatchPlan* p, const float* row_offsets) {
  if (!p) { set_error("bg_plan_set_row_offsets: plan is NULL"); return JAXENT_E_INVALID; }
  p->a.row_off = row_offsets;
  p->initialised = 0;  // the forward of jaxent_bg_state_init must follow
  return JAXENT_OK;
}

extern "C" int jaxent_bg_plan_workspace(const JaxentBatchPlan* p, void** ptr, size_t* bytes) {
  if (!p || !ptr || !bytes) { set_error("bg_plan_workspace: NULL argument"); return JAXENT_E_INVALID; }
  *ptr = p->ws;
  *bytes = p->ws_bytes;
  return JAXENT_OK;
}

static int bg_err(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s launch: %s", what, cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

static int bg_gemm2_phase(JaxentBatchPlan* p, cudaStream_t s) {
  const bg::BatchArgs& a = p->a;
  return jaxent_bg_gemm2(a.prob.packed, a.e3, a.P, a.n_fb, a.n_rg, a.Bn, a.NS, a.ksplit, s);
}

// reduce2 + per-replica tails on `s`; the MaxEnt kernel and its fold on `skl` (== s, or the plan's side stream between two
// steps of one jaxent_bg_steps call: forked after the tails, joined by the caller on ev_kl)
static int bg_tail_phase_split(JaxentBatchPlan* p, cudaStream_t s, cudaStream_t skl, bool fork) {
  const bg::BatchArgs& a = p->a;
  int rc;
  dim3 grid((2 * a.Rp + 31) / 32, (a.Bn + 31) / 32);
  bg::bt_reduce2_kernel<<<grid, 1024, 0, s>>>(a);
  if ((rc = bg_err("bt_reduce2"))) return rc;
  static size_t configured[2] = {0, 0};
  if (p->tail_smem > configured[p->general]) {
    cudaError_t e = p->general ? cudaFuncSetAttribute(bg::bt_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->tail_smem)
                               : cudaFuncSetAttribute(bg::bt_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->tail_smem);
    if (e != cudaSuccess) { set_error("bt_tail: %zu bytes of shared memory: %s", p->tail_smem, cudaGetErrorString(e)); return JAXENT_E_UNSUPPORTED; }
    configured[p->general] = p->tail_smem;
  }
  if (p->general) bg::bt_tail_kernel<true><<<a.Bn, TAIL_THREADS, p->tail_smem, s>>>(a);
  else bg::bt_tail_kernel<false><<<a.Bn, TAIL_THREADS, p->tail_smem, s>>>(a);
  if ((rc = bg_err("bt_tail"))) return rc;
  {
    static bool kl_configured = false;
    if (!kl_configured) {
      cudaError_t e = cudaFuncSetAttribute(bg::bt_kl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) { set_error("bt_kl: shared memory: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
      kl_configured = true;
    }
  }
  if (fork) {
    cudaError_t e = cudaEventRecord(p->ev_tail, s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(skl, p->ev_tail, 0);
    if (e != cudaSuccess) { set_error("bg_steps (fork): %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  }
  // beside the tensor-core kernel (197 KB of shared memory per SM) only ONE of these blocks fits on an SM: one block per SM then
  const int kl_blocks = (fork && a.ew_blocks > p->sm_count) ? p->sm_count : a.ew_blocks;
  bg::bt_kl_kernel<<<kl_blocks, (a.Bn / 4) * a.fy, (size_t)a.fy * KL_N * a.Bn * 8, skl>>>(a);
  if ((rc = bg_err("bt_kl"))) return rc;
  bg::bt_klfold_kernel<<<(a.Bn * 32 + 255) / 256, 256, 0, skl>>>(a, kl_blocks);
  if ((rc = bg_err("bt_klfold"))) return rc;
  if (fork) {
    cudaError_t e = cudaEventRecord(p->ev_kl, skl);
    if (e != cudaSuccess) { set_error("bg_steps (join): %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  }
  return JAXENT_OK;
}

static int bg_tail_phase(JaxentBatchPlan* p, cudaStream_t s) { return bg_tail_phase_split(p, s, s, false); }

static int bg_forward(JaxentBatchPlan* p, cudaStream_t s) {
  int rc = bg_gemm2_phase(p, s);
  if (rc) return rc;
  return bg_tail_phase(p, s);
}

static size_t bg_fused_smem(int NS) {
  return (size_t)bg::FG_STAGES * (bg::FG_B_BYTES + (size_t)NS * bg::FG_A_BYTES) + 256 + (size_t)bg::FG_EPI_WARPS * 32 * 8;
}

// phase 0: the backward product.  Bn a multiple of 128: one tensor-core kernel that also forms the masked gradients and the
// grad-dot partials in its epilogue (H stays in tensor memory).  Otherwise: gemm1 -> H, and bt_grad in phase 1.
static int bg_gemm1_phase(JaxentBatchPlan* p, cudaStream_t s) {
  bg::BatchArgs& a = p->a;
  if (p->fused_grad) {
    const size_t smem = bg_fused_smem(a.NS);
    static size_t configured = 0;
    if (smem > configured) {
      cudaError_t e = cudaFuncSetAttribute(bg::bt_fused_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("bt_fused_grad: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return JAXENT_E_UNSUPPORTED; }
      configured = smem;
    }
    bg::bt_fused_grad_kernel<<<a.dot_blocks, bg::FG_THREADS, smem, s>>>(a);
    return bg_err("bt_fused_grad");
  }
  return jaxent_bg_gemm1(a.prob.packed, a.cc3, a.H, a.n_fb, a.n_rg, a.Bn, a.NS, p->sm_count, s);
}

static int bg_update_phase(JaxentBatchPlan* p, cudaStream_t s) {
  const bg::BatchArgs& a = p->a;
  int rc;
  if (!p->fused_grad) {
    bg::bt_grad_kernel<<<a.ew_blocks, (a.Bn / 4) * a.fy, (size_t)a.fy * a.Bn * 8, s>>>(a);
    if ((rc = bg_err("bt_grad"))) return rc;
  }
  bg::bt_decide_kernel<<<(a.Bn * 32 + 255) / 256, 256, 0, s>>>(a);
  if ((rc = bg_err("bt_decide"))) return rc;
  bg::bt_adam_kernel<<<a.adam_blocks, a.adam_threads, (size_t)2 * a.adam_threads * 8, s>>>(a, 1);
  return bg_err("bt_adam");
}

extern "C" int jaxent_bg_state_init(JaxentBatchPlan* p, const float* z0, float zmax, float bc, float bh, const float* fw_host,
                                    const float* fs_host, const float* lr_host, jaxent_stream_t stream) {
  if (!p || !z0 || !fw_host || !fs_host) { set_error("bg_state_init: NULL argument"); return JAXENT_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  bg::BatchArgs& a = p->a;
  std::vector<bg::BCtl> init((size_t)a.Bn);
  for (int b = 0; b < a.Bn; ++b) {
    bg::BCtl c;
    memset(&c, 0, sizeof(c));
    c.lr = a.cfg.initial_learning_rate;
    c.model_lr = a.cfg.initial_learning_rate * a.cfg.model_parameters_lr_scale;
    c.cfg_lr = lr_host ? lr_host[b] : a.cfg.learning_rate;
    c.lse = zmax;
    c.bc = bc;
    c.bh = bh;
    c.kl_slot = -1;
    c.lr_eff = c.lr;
    c.bc1 = c.bc2 = 1.f;
    c.mask_frame = 1.f;
    c.S = 1.0;
    c.snap_now = -1;
    for (int i = 0; i < a.trk_n_thr; ++i) c.thr[i] = p->conv[i] * c.cfg_lr;  // create_convergence_thresholds, opt/track.py:12-21
    for (int i = 0; i < a.prob.n_slots; ++i) {
      c.fw[i] = fw_host[(size_t)b * JAXENT_MAX_SLOTS + i];
      c.fs[i] = fs_host[(size_t)b * JAXENT_MAX_SLOTS + i];
      if (a.prob.slots[i].kind == JAXENT_LOSS_MAXENT_CONVEX_KL) {
        c.kl_slot = i;
        c.kl_scale = c.fw[i] * c.fs[i];
      }
    }
    init[b] = c;
  }
  bg::BCtl* dinit = reinterpret_cast<bg::BCtl*>(p->ws + p->off_ctl_init);
  cudaError_t e = cudaMemcpyAsync(dinit, init.data(), init.size() * sizeof(bg::BCtl), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // `init` is pageable host memory: set-up time only
  if (e != cudaSuccess) { set_error("bg_state_init: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  bg::bt_init_kernel<<<592, 256, 0, s>>>(a, z0, dinit);
  int rc = bg_err("bt_init");
  if (rc) return rc;
  if (a.prob.prior != nullptr) {
    bg::bt_prior_kernel<<<1, 1024, 0, s>>>(a);
    if ((rc = bg_err("bt_prior"))) return rc;
  }
  if ((rc = launch_build_blobs(a.prob, a.blobs, s))) return rc;
  bg::bt_adam_kernel<<<a.adam_blocks, a.adam_threads, (size_t)2 * a.adam_threads * 8, s>>>(a, 0);
  if ((rc = bg_err("bt_adam(init)"))) return rc;
  rc = bg_forward(p, s);
  if (rc) return rc;
  p->initialised = 1;
  p->steps = 0;
  return JAXENT_OK;
}

// phases of one batched step: 0 gemm1 (backward product), 1 grad + decide + adam, 2 gemm2 (forward product),
// 3 reduce + per-replica tails + MaxEnt terms.  jaxent_bg_step runs all four.
extern "C" int jaxent_bg_step_phase(JaxentBatchPlan* p, int32_t phase, jaxent_stream_t stream) {
  if (!p || !p->initialised) { set_error("bg_step: jaxent_bg_state_init has not been called"); return JAXENT_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  switch (phase) {
    case 0: return bg_gemm1_phase(p, s);
    case 1: return bg_update_phase(p, s);
    case 2: return bg_gemm2_phase(p, s);
    case 3: return bg_tail_phase(p, s);
    default: set_error("bg_step_phase: phase must be 0..3"); return JAXENT_E_INVALID;
  }
}

extern "C" int jaxent_bg_step(JaxentBatchPlan* p, jaxent_stream_t stream) { return jaxent_bg_steps(p, 1, stream); }

// n steps in one call.  The per-(frame, replica) MaxEnt kernel of step k (bt_kl + its fold: memory / FMA bound) does not touch
// anything the backward product of step k + 1 reads (features, c pieces), so between two steps of the same call it is
// enqueued on the plan's side stream and joins the main stream just before bt_grad, which needs its sums: it then runs
// under the tensor-core kernel instead of after it.  The last step of a call keeps everything on `stream`, so the state a
// caller can observe after the call (loss records, control blocks, tracker snapshots) is complete -- same contract as n
// calls of jaxent_bg_step.  Fork / join are events: no host synchronisation, capturable in a CUDA graph.
extern "C" int jaxent_bg_steps(JaxentBatchPlan* p, int32_t n, jaxent_stream_t stream) {
  if (!p || !p->initialised) { set_error("bg_step: jaxent_bg_state_init has not been called"); return JAXENT_E_INVALID; }
  if (n < 1) { set_error("bg_steps: n must be >= 1"); return JAXENT_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  bool pending = false;
  int rc;
  for (int i = 0; i < n; ++i) {
    if ((rc = bg_gemm1_phase(p, s))) return rc;
    if (pending) {
      cudaError_t e = cudaStreamWaitEvent(s, p->ev_kl, 0);
      if (e != cudaSuccess) { set_error("bg_steps: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
      pending = false;
    }
    if ((rc = bg_update_phase(p, s))) return rc;
    if ((rc = bg_gemm2_phase(p, s))) return rc;
    const bool overlap = p->side != nullptr && i + 1 < n;
    if ((rc = bg_tail_phase_split(p, s, overlap ? p->side : s, overlap))) return rc;
    pending = overlap;
    p->steps += 1;
  }
  return JAXENT_OK;
}

extern "C" int jaxent_bg_plan_buffer(const JaxentBatchPlan* p, int32_t which, void** ptr, size_t* bytes) {
  if (!p || !ptr || !bytes) { set_error("bg_plan_buffer: NULL argument"); return JAXENT_E_INVALID; }
  const bg::BatchArgs& a = p->a;
  const size_t fb = (size_t)a.F * a.Bn * 4;
  switch (which) {
    case 0: *ptr = a.z; *bytes = fb; break;                                             /* logits  [F][Bn]      */
    case 1: *ptr = a.e; *bytes = fb; break;                                             /* exp(z - lse) [F][Bn] */
    case 2: *ptr = a.records; *bytes = (size_t)bg::RB * a.Bn * sizeof(JaxentLossRecord); break;
    case 3: *ptr = a.ctl; *bytes = (size_t)a.Bn * sizeof(bg::BCtl); break;
    case 4: *ptr = a.logpf; *bytes = (size_t)a.Bn * a.prob.R * 4; break;
    case 5: *ptr = a.uptake; *bytes = (size_t)a.Bn * (a.prob.uptake ? a.prob.T : 1) * a.prob.R * 4; break;
    case 6: *ptr = a.H; *bytes = (size_t)a.n_fb * 128 * a.Bn * 4; break;
    case 7: *ptr = a.red; *bytes = (size_t)a.Bn * 2 * a.Rp * 8; break;
    case 8: *ptr = a.g[0]; *bytes = fb; break;                                          /* masked gradients, buffer 0 */
    case 9: *ptr = a.g[1]; *bytes = fb; break;                                          /* masked gradients, buffer 1 */
    case 10: *ptr = a.gsel; *bytes = 4; break;                                          /* which of 8 / 9 is current  */
    default: set_error("bg_plan_buffer: unknown buffer id %d", which); return JAXENT_E_INVALID;
  }
  return JAXENT_OK;
}

extern "C" int jaxent_bg_ctl_sizeof(void) { return (int)sizeof(bg::BCtl); }

extern "C" int jaxent_bg_ctl_offset(int32_t field) {
  switch (field) {
    case 0: return (int)offsetof(bg::BCtl, S);
    case 1: return (int)offsetof(bg::BCtl, bc);
    case 2: return (int)offsetof(bg::BCtl, bh);
    case 3: return (int)offsetof(bg::BCtl, step);
    case 4: return (int)offsetof(bg::BCtl, lse);
    case 5: return (int)offsetof(bg::BCtl, trk_steps_since);
    case 6: return (int)offsetof(bg::BCtl, trk_best_step);
    case 7: return (int)offsetof(bg::BCtl, trk_best_val);
    default: return -1;
  }
}

// ---- on-device loop of every replica (reference _optimise_pure / ConvergenceTracker, opt/run.py:141-232, opt/track.py:128-197)
extern "C" int jaxent_bg_track_init(JaxentBatchPlan* p, const float* convergence, int32_t n, float ema_alpha, int32_t min_steps,
                                    float tolerance, float* ema_w, float* snap_w, float* snap_ema, JaxentTrackSnapshot* snaps) {
  if (!p) { set_error("bg_track_init: plan is NULL"); return JAXENT_E_INVALID; }
  if (!convergence || n < 1 || n > JAXENT_TRACK_MAX_THRESHOLDS || !ema_w || !snap_w || !snap_ema || !snaps) {
    set_error("bg_track_init: need 1..%d thresholds and the EMA / snapshot buffers", JAXENT_TRACK_MAX_THRESHOLDS);
    return JAXENT_E_INVALID;
  }
  if (!(ema_alpha >= 0.f && ema_alpha <= 1.f) || min_steps < 0) { set_error("bg_track_init: invalid ema_alpha / min_steps"); return JAXENT_E_INVALID; }
  for (int i = 1; i < n; ++i)
    if (convergence[i] > convergence[i - 1]) { set_error("bg_track_init: thresholds must be in descending order"); return JAXENT_E_INVALID; }
  bg::BatchArgs& a = p->a;
  a.trk_on = 1;
  a.trk_n_thr = n;
  a.trk_min_steps = min_steps;
  a.trk_alpha = ema_alpha;
  a.trk_tolerance = tolerance;
  a.ema_w = ema_w;
  a.snap_w = snap_w;
  a.snap_ema = snap_ema;
  a.snaps = snaps;
  for (int i = 0; i < n; ++i) p->conv[i] = convergence[i];
  p->initialised = 0;  // the tracker state lives in the control blocks: jaxent_bg_state_init must follow
  return JAXENT_OK;
}

// which loop the tracker follows: 0 (default) the Python loop `_optimise` (opt/run.py:235-373: ConvergenceTracker, oscillation
// reset, a snapshot per threshold), 1 `_optimise_pure` (opt/run.py:141-232, what batch_optimise vmaps: functional tracker,
// EMAs restart with every threshold, every step is a history entry -> snapshot slot 0 follows the best validation loss)
extern "C" int jaxent_bg_track_mode(JaxentBatchPlan* p, int32_t mode) {
  if (!p) { set_error("bg_track_mode: plan is NULL"); return JAXENT_E_INVALID; }
  if (mode != 0 && mode != 1) { set_error("bg_track_mode: mode must be 0 (Python loop) or 1 (pure loop)"); return JAXENT_E_INVALID; }
  p->a.trk_pure = mode;
  p->initialised = 0;  // the tracker state lives in the control blocks: jaxent_bg_state_init must follow
  return JAXENT_OK;
}

namespace jx {
namespace bg {
__global__ void bt_status_kernel(const BatchArgs a, JaxentTrackStatus* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.Bn) return;
  const BCtl* c = a.ctl + b;
  out[b].stop = c->stop;
  out[b].reason = c->trk_reason;
  out[b].stop_step = c->trk_stop_step;
  out[b].n_snapshots = c->trk_nsnap;
  out[b].steps_done = c->step;
  out[b].threshold_idx = c->trk_idx;
  out[b].last_loss = (float)c->trk_prev_loss;
  out[b].ema_loss_delta = (float)c->trk_ema;
}
}  // namespace bg
}  // namespace jx

// one JaxentTrackStatus per replica written to `status` (device memory, n_replicas entries)
extern "C" int jaxent_bg_track_status(JaxentBatchPlan* p, JaxentTrackStatus* status, jaxent_stream_t stream) {
  if (!p || !p->initialised || !status) { set_error("bg_track_status: invalid argument"); return JAXENT_E_INVALID; }
  bg::bt_status_kernel<<<(p->a.Bn + 63) / 64, 64, 0, (cudaStream_t)stream>>>(p->a, status);
  return bg_err("bt_status");
}
