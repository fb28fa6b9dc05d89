// The streaming pass of the BV MaxEnt optimise step for sm_100a, plus its small helpers.
//
// One pass over the packed (frame-quad major) feature arrays does, per tile of FT=8 frames:
//   phase A  column sums  H_f = sum_r cc_r*Nc[r,f] + ch_r*Nh[r,f]        (backward of step k;
//            reference: the transpose products jax.grad emits for utils/jax_fn.py:30-38)
//   Adam     g_f = mask*w_f*(H_f + kappa_f - hbar), hbar = sum_j w_j (H_j + kappa_j); grad-dot partial, optax adam for BOTH plateau
//            learning rates (opt/optimiser.py:396-424), z' -> e' = exp(z' - lse)
//   phase B  row sums     acc_r += e'_f * N[r,f]  for both branches      (forward of step k+1;
//            reference: frame_average_features, utils/jax_fn.py:15-38)
// so the features cross HBM once per optimise step (8*R*F bytes).  Tiles are staged into shared
// memory by cp.async.bulk (TMA bulk copy, mbarrier completion) in a ring; each thread owns one
// residue row (RPT rows when R > 512) and keeps its tile elements in registers between phases.
#include <cstdio>

#include "jx_internal.cuh"

namespace jx {

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ pack
// (R,F) row-major -> [quad][array][row][4 frames].  One-off re-layout at plan build.
__global__ void pack_kernel(const float* __restrict__ heavy, const float* __restrict__ acceptor, int R, long long F,
                            long long ld, long long frame_offset, long long F_dst, float* __restrict__ packed) {
  __shared__ float tile[2][32][33];
  const long long q0 = frame_offset / 4;
  const long long Fpad = ((F_dst + FT - 1) / FT) * FT;
  const bool last_chunk = (frame_offset + F) >= F_dst;
  const long long f_end = last_chunk ? Fpad - frame_offset : F;  // local frames to emit (incl. zero pad)
  const long long fb = (long long)blockIdx.x * 32;               // local frame block
  const int rb = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = rb + i;
    const long long f = fb + tx;
    float a = 0.f, b = 0.f;
    if (r < R && f < F) {
      a = heavy[(long long)r * ld + f];
      b = acceptor[(long long)r * ld + f];
    }
    tile[0][i][tx] = a;
    tile[1][i][tx] = b;
  }
  __syncthreads();
  // write: for each quad (8 per block), each array, rows rb..rb+31, 4 frames
  // thread (tx,ty): row = tx, quad = ty
  const int r = rb + tx;
  const long long fq = fb + (long long)ty * 4;
  if (r < R && fq < f_end) {
    const long long q = q0 + fq / 4;
    for (int arr = 0; arr < 2; ++arr) {
      float4 o = make_float4(tile[arr][tx][ty * 4 + 0], tile[arr][tx][ty * 4 + 1], tile[arr][tx][ty * 4 + 2],
                             tile[arr][tx][ty * 4 + 3]);
      reinterpret_cast<float4*>(packed)[(q * 2 + arr) * R + r] = o;
    }
  }
}

// (R,F) fp32 row-major -> u8 [tile of 8 frames][array][row][8 bytes]; values must be integers in [0,255]
__global__ void pack_u8_kernel(const float* __restrict__ heavy, const float* __restrict__ acceptor, int R, long long F,
                               long long ld, long long Fpad, unsigned char* __restrict__ packed) {
  __shared__ float tile[2][32][33];
  const long long fb = (long long)blockIdx.x * 32;
  const int rb = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = rb + i;
    const long long f = fb + tx;
    float a = 0.f, b = 0.f;
    if (r < R && f < F) {
      a = heavy[(long long)r * ld + f];
      b = acceptor[(long long)r * ld + f];
    }
    tile[0][i][tx] = a;
    tile[1][i][tx] = b;
  }
  __syncthreads();
  // thread (tx = row, ty = tile-of-8 within the block (4 of them) x array (2))
  const int r = rb + tx;
  const int t8 = ty >> 1, arr = ty & 1;
  const long long f0 = fb + (long long)t8 * 8;
  if (r < R && f0 < Fpad) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      lo |= ((uint32_t)__float2int_rn(tile[arr][tx][t8 * 8 + j]) & 0xffu) << (8 * j);
      hi |= ((uint32_t)__float2int_rn(tile[arr][tx][t8 * 8 + 4 + j]) & 0xffu) << (8 * j);
    }
    const long long tix = f0 / 8;
    reinterpret_cast<uint2*>(packed)[(tix * 2 + arr) * R + r] = make_uint2(lo, hi);
  }
}

// flag[0] <- 0 if any value is not an integer in [0,255] (flag must be initialised to 1 by the caller)
__global__ void fit_u8_kernel(const float* __restrict__ x, long long n_rows, long long F, long long ld, int* flag) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  bool bad = false;
  for (long long i = i0; i < n_rows * F; i += stride) {
    const long long r = i / F, f = i - r * F;
    const float v = x[r * ld + f];
    bad = bad || !(v >= 0.f && v <= 255.f && v == rintf(v));
  }
  if (bad) *flag = 0;
}

// ------------------------------------------------------------------------------------------ small reductions
__global__ void prior_partial_kernel(const float* __restrict__ prior, long long F, double eps, double* out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (long long f = threadIdx.x; f < F; f += blockDim.x) s += (double)fabsf(prior[f]) + eps;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    out[0] = t;
  }
}

__global__ void max_kernel(const float* __restrict__ z, long long F, float* out) {
  __shared__ float sh[32];
  float s = -INFINITY;
  for (long long f = threadIdx.x; f < F; f += blockDim.x) s = fmaxf(s, z[f]);
  for (int o = 16; o; o >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = -INFINITY;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmaxf(t, sh[i]);
    out[0] = t;
  }
}

// ------------------------------------------------------------------------------------------ the pass
// Thread roles: `NTc` consumer threads (one residue row each, RPT rows when R > 512) plus one
// dedicated "Adam warp" (the last warp).  Software pipeline, one tile deep:
//
//   consumers, tile i :  wait TMA -> tile to registers -> phase A(i) -> partials to smem -> bar.arrive PART[i&1]
//                        bar.sync E[(i-1)&1] -> phase B(i-1) on the previous tile's registers
//   Adam warp, tile i :  bar.sync PART[i&1] -> refill the ring slot -> Adam(i) -> e'(i) to smem -> bar.arrive E[i&1]
//
// so the serial per-frame optimiser chain of tile i runs while the consumers stream tile i+1.
constexpr int BAR_PART0 = 1, BAR_E0 = 3;  // named barrier ids (0 is __syncthreads)
constexpr int FLUSH_TILES = 128;          // fp32 tile sums are folded into the fp64 partials every 128 tiles (1024 frames)

__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int RPT>
struct TileRegs {
  float4 nc[RPT][QT], nh[RPT][QT];
};

// FMT 0: fp32 features, sub-tile = [quad][array][row][4 floats];  FMT 1: u8 features (lossless for integer contact
// counts <= 255), sub-tile = [array][row][8 bytes].  `rstride` = bytes between consecutive (quad,array) / array blocks.
__device__ __forceinline__ float4 u8x4_to_float4(uint32_t v) {
  return make_float4((float)(v & 0xffu), (float)((v >> 8) & 0xffu), (float)((v >> 16) & 0xffu), (float)(v >> 24));
}

template <int RPT, int FMT>
__device__ __forceinline__ void load_tile(TileRegs<RPT>& d, const unsigned char* sub, uint32_t rstride, const uint32_t (&rowoff)[RPT]) {
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const unsigned char* p = sub + rowoff[i];
    if (FMT == 0) {
      d.nc[i][0] = *reinterpret_cast<const float4*>(p);
      d.nh[i][0] = *reinterpret_cast<const float4*>(p + rstride);
      d.nc[i][1] = *reinterpret_cast<const float4*>(p + 2 * rstride);
      d.nh[i][1] = *reinterpret_cast<const float4*>(p + 3 * rstride);
    } else {
      const uint2 c = *reinterpret_cast<const uint2*>(p);
      const uint2 h = *reinterpret_cast<const uint2*>(p + rstride);
      d.nc[i][0] = u8x4_to_float4(c.x);
      d.nc[i][1] = u8x4_to_float4(c.y);
      d.nh[i][0] = u8x4_to_float4(h.x);
      d.nh[i][1] = u8x4_to_float4(h.y);
    }
  }
}

// phase A: H partials of this thread's rows, transposing warp reduction, one float per (warp, frame)
// Centred: every row's product carries -k_r = -fl(cc_r <Nc>_r + ch_r <Nh>_r) (nk, from the tail), so the column sum is
// H_f - sum_r k_r = sum_r cc_r (Nc[r,f] - <Nc>_r) + ch_r (Nh[r,f] - <Nh>_r): same instruction count (fma + fma instead of
// mul + fma), and the quantity the softmax Jacobian needs, h_f - hbar, no longer hinges on the last bits of a large fp32 H_f.
// The inner fma holds cc*Nc - k ~ -ch*<Nh>: the acceptor term is normally the smaller of the two, so that is the cheaper rounding.
template <int RPT>
__device__ __forceinline__ void phase_a(const TileRegs<RPT>& d, const float (&cc)[RPT], const float (&ch)[RPT], const float (&nk)[RPT],
                                        float* s_part_warp, int lane, int jred) {
  float p[FT];
  float nk0 = nk[0];  // the constants of all rows of this thread enter with the first row's product
#pragma unroll
  for (int i = 1; i < RPT; ++i) nk0 += nk[i];
#pragma unroll
  for (int q = 0; q < QT; ++q) {
    p[4 * q + 0] = fmaf(ch[0], d.nh[0][q].x, fmaf(cc[0], d.nc[0][q].x, nk0));
    p[4 * q + 1] = fmaf(ch[0], d.nh[0][q].y, fmaf(cc[0], d.nc[0][q].y, nk0));
    p[4 * q + 2] = fmaf(ch[0], d.nh[0][q].z, fmaf(cc[0], d.nc[0][q].z, nk0));
    p[4 * q + 3] = fmaf(ch[0], d.nh[0][q].w, fmaf(cc[0], d.nc[0][q].w, nk0));
  }
#pragma unroll
  for (int i = 1; i < RPT; ++i) {
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      p[4 * q + 0] = fmaf(cc[i], d.nc[i][q].x, fmaf(ch[i], d.nh[i][q].x, p[4 * q + 0]));
      p[4 * q + 1] = fmaf(cc[i], d.nc[i][q].y, fmaf(ch[i], d.nh[i][q].y, p[4 * q + 1]));
      p[4 * q + 2] = fmaf(cc[i], d.nc[i][q].z, fmaf(ch[i], d.nh[i][q].z, p[4 * q + 2]));
      p[4 * q + 3] = fmaf(cc[i], d.nc[i][q].w, fmaf(ch[i], d.nh[i][q].w, p[4 * q + 3]));
    }
  }
  // 8 values x 32 lanes -> each lane ends with the warp total of frame jred (9 shuffles)
  const bool hi = lane & 16;
  float q4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = hi ? p[i] : p[i + 4];
    const float keep = hi ? p[i + 4] : p[i];
    q4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  const bool b3 = lane & 8;
  float r2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b3 ? q4[i] : q4[i + 2];
    const float keep = b3 ? q4[i + 2] : q4[i];
    r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const bool b2_ = lane & 4;
  const float send = b2_ ? r2[0] : r2[1];
  const float keep = b2_ ? r2[1] : r2[0];
  float s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if ((lane & 3) == 0) s_part_warp[jred] = s;
}

// phase B: acc[i][k] += sum_f e'_f * N[row_i, f]  (k: A/Nc, A/Nh, B/Nc, B/Nh)
template <int RPT>
__device__ __forceinline__ void phase_b(const TileRegs<RPT>& d, const float* s_e, float (&acc)[RPT][4]) {
  const float4* e4 = reinterpret_cast<const float4*>(s_e);
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const float4 e0 = e4[2 * br], e1 = e4[2 * br + 1];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      float tc = d.nc[i][0].x * e0.x;
      tc = fmaf(d.nc[i][0].y, e0.y, tc);
      tc = fmaf(d.nc[i][0].z, e0.z, tc);
      tc = fmaf(d.nc[i][0].w, e0.w, tc);
      tc = fmaf(d.nc[i][1].x, e1.x, tc);
      tc = fmaf(d.nc[i][1].y, e1.y, tc);
      tc = fmaf(d.nc[i][1].z, e1.z, tc);
      tc = fmaf(d.nc[i][1].w, e1.w, tc);
      float th = d.nh[i][0].x * e0.x;
      th = fmaf(d.nh[i][0].y, e0.y, th);
      th = fmaf(d.nh[i][0].z, e0.z, th);
      th = fmaf(d.nh[i][0].w, e0.w, th);
      th = fmaf(d.nh[i][1].x, e1.x, th);
      th = fmaf(d.nh[i][1].y, e1.y, th);
      th = fmaf(d.nh[i][1].z, e1.z, th);
      th = fmaf(d.nh[i][1].w, e1.w, th);
      acc[i][2 * br + 0] += tc;
      acc[i][2 * br + 1] += th;
    }
  }
}

// hbar = sum_j w_j h_j of the softmax Jacobian (dL/dz_f = w_f (h_f - hbar)), both parts:
//   data    sum_r c_r lnPF_r                     exact identity, from the tail's control block
//   MaxEnt  sum_j w_j kappa_j = kl[0]            reduced by the tail's frame CTAs over ALL ranks' frames
// The MaxEnt part equals lambda * sum_j p_j eps / (w_j + eps): negligible while every w_j >> eps = 1e-10 / F, but it
// tends to lambda times the prior mass of the collapsed frames once weights fall below eps (logit spread >~ 37 at
// F = 1e6), so it is part of the gradient exactly as jax.value_and_grad forms it (reference opt/optimiser.py:391-393,
// opt/losses.py:304-322).  Every Adam warp gathers the words itself (world * KL_N 16-byte loads from the LOCAL inbox,
// issued while the first tiles are still in flight); frame-sharded, the peers' tails published them before their own
// R x T tail finished, so the spin is normally satisfied by the first load.  CTA 0 also completes the loss record of
// the step those sums belong to.
// (plain pointers instead of the Xchg struct: passing a kernel-parameter struct by reference forces a local-memory copy)
static __device__ __noinline__ double gather_hbar(const unsigned char* kl_words, int world, unsigned char* err_word, unsigned ep,
                                                 const Ctl* __restrict__ ctl, JaxentLossRecord* records, int lane, bool finish) {
  double v = 0.0;
  if (lane < world * KL_N) {
    if (world == 1) v = __ldcg(reinterpret_cast<const double*>(kl_words) + lane);
    else v = ll_load(err_word, kl_words + (size_t)lane * 16, ep);  // [rank][KL_N] 16-byte stamped words
  }
  __syncwarp();
  double kl[KL_N];
#pragma unroll
  for (int k = 0; k < KL_N; ++k) kl[k] = 0.0;
  for (int r = 0; r < world; ++r) {  // rank order: identical on every rank
#pragma unroll
    for (int k = 0; k < KL_N; ++k) kl[k] += __shfl_sync(0xffffffffu, v, r * KL_N + k);
  }
  if (finish && lane == 0) finish_record(ctl, kl[1] + kl[2], records, JAXENT_MAX_SLOTS);
  return ctl->hbar_data + kl[0];
}
__device__ __forceinline__ double pass_hbar(const PassArgs& a, unsigned ep, const Ctl* __restrict__ ctl, int lane, bool finish) {
  const unsigned char* kl_words = a.x.local + (a.x.world == 1 ? xchg_plain_kl_off(a.x) : xchg_kl_off(a.x, (int)(ep & 1u), 0));
  return gather_hbar(kl_words, a.x.world, a.x.local + xchg_err_off(a.x), ep, ctl, a.records, lane, finish);
}


// ------------------------------------------------------------------------------------------ fixed-point row sums
// fx_scale() / fx_add() live in jx_internal.cuh.  The pass only ACCUMULATES: every CTA adds its tile sums into the fixed-point
// accumulators and leaves its three scalars (sum e' of both branches, grad-dot) in cta_scal.  The sums are completed -- and
// the accumulators cleared, each column by its one reader -- after the kernel boundary: by the tail kernel (fx_take_col /
// fold_scal3: all CTAs in parallel, stamped peer stores spread over several CTAs when frame-sharded) or, for the three-phase
// ABI, by jaxent_bv_reduce.  There is no ticket, no last-CTA serial section and no epoch store at the end of the pass (the
// tail advances the epoch).

// ------------------------------------------------------------------------------------------ per-frame uptake mode
// ForwardPass.average_first = False (reference models/core.py:302-304, BV_uptake_ForwardPass models/HDX/forward.py:57-74):
// the uptake is evaluated per (time point, residue, frame) and THEN frame-averaged, so the pass carries the non-linearity:
//   u[t,r,f] = 1 - exp(-k_r tau_t exp(-(bc Nc[r,f] + bh Nh[r,f])))
//   phase A  H_f = sum_{t,r} G[t,r] u[t,r,f]            G = dL/d<uptake> from the tail
//   phase B  acc[branch][t][r] += e'_f (1 - u[t,r,f])  (the tail forms <u> = 1 - acc / S in double precision)
// F*R*(T+1) exponentials per step make this mode special-function bound rather than HBM bound.  With frozen BV parameters the
// uptake values of a unit serve both phases.  With optimised BV parameters (PassArgs::pf_db) their gradient is a frame sum
// of its own (pf_dbeta_kernel, below) and phase B re-evaluates the exponentials at the updated parameters of each plateau
// branch (pf_consumer<TM, true>).
constexpr int PF_TMAX = 8;

// 1 - exp(-x) for x >= 0 on the special-function unit: ex2.approx is accurate to 2 ulp, so the absolute error is below
// 2.4e-7 on a quantity in [0, 1] that is only ever summed with O(1) weights (frame average, peptide map, G-weighted sum)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float uptake_of(float x) { return 1.f - ex2_approx(-1.4426950408889634f * x); }

// 2^x on the FMA pipe: range reduction with the 1.5 * 2^23 trick, degree-6 near-minimax polynomial on [-0.5, 0.5] (Chebyshev
// interpolation: 8e-8 max relative error, bias 1e-9 -- ex2.approx is specified to 2 ulp, and a frame AVERAGE of exponentials
// keeps whatever bias the approximation has), exponent insertion with an integer shift-add.  The Adam warps of the per-frame
// kernel use it for e' (their MUFU instructions would queue behind the consumers'); for the consumers themselves see PF_EXP.
#define JX_E2_C1 0.6931471824645996f
#define JX_E2_C2 0.24022650718688965f
#define JX_E2_C3 0.05550327152013779f
#define JX_E2_C4 0.009618056938052177f
#define JX_E2_C5 0.0013400427997112274f
#define JX_E2_C6 0.00015461444854736328f
__device__ __forceinline__ float exp2_fma(float x) {
  x = fmaxf(x, -126.f);
  const float y = x + 12582912.f;
  const float f = x - (y - 12582912.f);  // [-0.5, 0.5]
  float p = JX_E2_C6;
  p = fmaf(p, f, JX_E2_C5);
  p = fmaf(p, f, JX_E2_C4);
  p = fmaf(p, f, JX_E2_C3);
  p = fmaf(p, f, JX_E2_C2);
  p = fmaf(p, f, JX_E2_C1);
  p = fmaf(p, f, 1.f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(y) << 23));
}
// two values at once: add / fma as packed f32x2 instructions
__device__ __forceinline__ float2 exp2_fma2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f), one = make_float2(1.f, 1.f), mone = make_float2(-1.f, -1.f);
  const float2 y = __fadd2_rn(x, magic);
  const float2 t = __ffma2_rn(magic, mone, y);  // y - magic (exact)
  const float2 f = __ffma2_rn(t, mone, x);      // x - t     (exact)
  float2 p = make_float2(JX_E2_C6, JX_E2_C6);
  p = __ffma2_rn(p, f, make_float2(JX_E2_C5, JX_E2_C5));
  p = __ffma2_rn(p, f, make_float2(JX_E2_C4, JX_E2_C4));
  p = __ffma2_rn(p, f, make_float2(JX_E2_C3, JX_E2_C3));
  p = __ffma2_rn(p, f, make_float2(JX_E2_C2, JX_E2_C2));
  p = __ffma2_rn(p, f, make_float2(JX_E2_C1, JX_E2_C1));
  p = __ffma2_rn(p, f, one);
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(y.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(y.y) << 23)));
}
// which unit evaluates the F*R*(T+1) exponentials of the per-frame mode: 0 MUFU (ex2.approx), 1 the FMA-pipe polynomial,
// 2 both (first value of a time-point pair on the FMA pipe, second on MUFU).  Measured at config 4 (1M x 500 x 8, B200):
// pass 1957 / 2986 / 2632 us for 0 / 1 / 2 -- the polynomial's integer and min/max instructions and its extra live registers
// (72 instead of 56 bytes of spill stack) cost more than the MUFU queue saves -- and the same parity against the fp64 oracle
// in all three (loss 1e-7, gradients 1e-6 relative): MUFU it is.
#ifndef PF_EXP
#define PF_EXP 0
#endif
__device__ __forceinline__ float pf_exp2(float x) { return PF_EXP == 0 ? ex2_approx(x) : exp2_fma(x); }
__device__ __forceinline__ float2 pf_exp2_pair(float2 x) {
  if (PF_EXP == 0) return make_float2(ex2_approx(x.x), ex2_approx(x.y));
  if (PF_EXP == 2) return make_float2(exp2_fma(x.x), ex2_approx(x.y));
  return exp2_fma2(x);
}

constexpr int PF_FU = 4;  // frames per unit of work in this mode: a tile is handled as its two quads (see pf_consumer)

__device__ __forceinline__ float2 bcast2(float x) { return make_float2(x, x); }

// One consumer thread owns one residue row.  A tile (8 frames) is processed as its two quads of 4 frames so that TWO units
// are in flight per thread: phase A of unit h+1 (the exponentials) is issued before phase B of unit h waits for the Adam
// warp, which hides the H_f -> Adam -> e' latency behind special-function work.  The uptake values of a unit stay in
// registers between its two phases (4 frames x TM time points, as packed pairs: fma.rn.f32x2 halves the issue slots of the
// multiply / accumulate instructions around each ex2).  -log2(e) k_r tau_t and G[t][r] live in per-thread slots of shared
// memory (each thread reads back only what it wrote: no barrier).  The special-function unit shares its issue queue with
// shared-memory and shuffle instructions, so the reduction of H_f over the rows is NOT done here: every thread stores its
// four per-frame partials (one 16-byte store) and the Adam warp, which has time to spare, sums them.
// DB (BV parameters optimised, mode 1): the forward half must see the UPDATED parameters, which differ between the two plateau
// branches (bcn[br], bhn[br] from derive_spec).  Phase B then re-evaluates the exponentials from the features of the unit
// (kept in registers instead of the uptake values): three exponential sweeps in this kernel instead of one.
template <int TM, bool DB>
__device__ __forceinline__ void pf_consumer(const PassArgs& a, unsigned char* ring, uint64_t* bars, float4* s_rows, const float* s_e,
                                             float4* s_tab, int n, int mode, float bc, float bh, int NTc, int nbar,
                                             float bcnA = 0.f, float bhnA = 0.f, float bcnB = 0.f, float bhnB = 0.f) {
  constexpr int TP = TM / 2;
  const int R = a.R, T = a.pf_T;
  const int tid = threadIdx.x;
  const bool ok = tid < R;
  const int rc = ok ? tid : R - 1;
  const uint32_t rowoff = (uint32_t)rc * 16u, rstride = (uint32_t)R * 16u;
  float4* my_tab = s_tab + tid;  // [TP][NTc] {kt(2 tp), kt(2 tp + 1), G(2 tp), G(2 tp + 1)}
  // centred: the tail's k_r = fl(sum_t G[t,r] (1 - <u>[t,r])) replaces sum_t G[t,r], so the row partial is sum_t G (u_f - <u>)
  const float Gsum = (ok && mode) ? a.ck[rc] : 0.f;
  {
    const float kr = -1.4426950408889634f * a.pf_k[rc];
#pragma unroll
    for (int tp = 0; tp < TP; ++tp) {
      const int t0_ = 2 * tp, t1_ = 2 * tp + 1;
      const float2 g2 = make_float2((ok && mode && t0_ < T) ? a.pf_G[(size_t)t0_ * R + rc] : 0.f,
                                    (ok && mode && t1_ < T) ? a.pf_G[(size_t)t1_ * R + rc] : 0.f);
      my_tab[tp * NTc] = make_float4(t0_ < T ? kr * a.pf_tp[t0_] : 0.f, t1_ < T ? kr * a.pf_tp[t1_] : 0.f, g2.x, g2.y);  // kt = -log2(e) k_r tau_t
    }
  }
  float2 acc[2][TP];
#pragma unroll
  for (int tp = 0; tp < TP; ++tp) acc[0][tp] = acc[1][tp] = make_float2(0.f, 0.f);
  const float nlog2e_bc = -1.4426950408889634f * bc, nlog2e_bh = -1.4426950408889634f * bh;
  JCLK_DECL(clk_ring);
  JCLK_DECL(clk_a);
  JCLK_DECL(clk_e);
  JCLK_DECL(clk_b);
  int slot = 0;
  uint32_t parity = 0;

  // phase A of one quad: u[j][tp] = exp(-k tau / PF) = 1 - uptake; per-frame partial sums of G.(1 - u) to the Adam warp
  const float nl2e_bcn[2] = {-1.4426950408889634f * bcnA, -1.4426950408889634f * bcnB};
  const float nl2e_bhn[2] = {-1.4426950408889634f * bhnA, -1.4426950408889634f * bhnB};
  auto phase_a_pf = [&](int half, float2 (&u)[PF_FU][TP]) {
    JCLK_TIC(tm0);
    if (half == 0) mbar_wait(smem_u32(&bars[slot]), parity);
    JCLK_TOC(clk_ring, tm0);
    JCLK_TIC(ta0);
    const unsigned char* sub = ring + (size_t)slot * a.stage_stride + rowoff + (half ? 2 * rstride : 0u);
    const float4 c4 = *reinterpret_cast<const float4*>(sub), h4 = *reinterpret_cast<const float4*>(sub + rstride);
    if (half) {
      if (++slot == a.stages) {
        slot = 0;
        parity ^= 1u;
      }
    }
    const float nc[PF_FU] = {c4.x, c4.y, c4.z, c4.w}, nh[PF_FU] = {h4.x, h4.y, h4.z, h4.w};
    float2 pinv[PF_FU], p2[PF_FU];
#pragma unroll
    for (int j = 0; j < PF_FU; ++j) {
      pinv[j] = bcast2(pf_exp2(fmaf(nlog2e_bc, nc[j], nlog2e_bh * nh[j])));  // exp(-lnPF)
      p2[j] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int tp = 0; tp < TP; ++tp) {
      const float4 kg = my_tab[tp * NTc];
      const float2 kt2 = make_float2(kg.x, kg.y), G2 = make_float2(kg.z, kg.w);
#pragma unroll
      for (int j = 0; j < PF_FU; ++j) {
        const float2 x2 = __fmul2_rn(kt2, pinv[j]);
        const float2 e2 = pf_exp2_pair(x2);
        if (!DB) u[j][tp] = e2;
        p2[j] = __ffma2_rn(G2, e2, p2[j]);
      }
    }
    if (DB) {  // the unit's features stand in for its uptake values: u[j][0] = (Nc, Nh) of frame j
#pragma unroll
      for (int j = 0; j < PF_FU; ++j) u[j][0] = make_float2(nc[j], nh[j]);
    }
    // per-row partials of H_f for the 4 frames: the Adam warp of this quad sums them over the rows
    s_rows[half * NTc + tid] = make_float4(Gsum - (p2[0].x + p2[0].y), Gsum - (p2[1].x + p2[1].y), Gsum - (p2[2].x + p2[2].y),
                                           Gsum - (p2[3].x + p2[3].y));
    bar_arrive(BAR_PART0 + half, nbar);
    JCLK_TOC(clk_a, ta0);
  };
  // phase B of one quad: acc[branch][t] += e'_f exp(-k tau / PF)
  auto phase_b_pf = [&](int half, const float2 (&u)[PF_FU][TP]) {
    JCLK_TIC(te0);
    bar_sync(BAR_E0 + half, nbar);  // e' of both branches from the Adam warp
    JCLK_TOC(clk_e, te0);
    JCLK_TIC(tb0);
    const float4* e4 = reinterpret_cast<const float4*>(s_e + half * 2 * PF_FU);
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      const float4 e = e4[br];
      const float2 ev[PF_FU] = {bcast2(e.x), bcast2(e.y), bcast2(e.z), bcast2(e.w)};
      if (DB) {
        float2 pinv[PF_FU];
#pragma unroll
        for (int j = 0; j < PF_FU; ++j) pinv[j] = bcast2(pf_exp2(fmaf(nl2e_bcn[br], u[j][0].x, nl2e_bhn[br] * u[j][0].y)));
#pragma unroll
        for (int tp = 0; tp < TP; ++tp) {
          const float4 kg = my_tab[tp * NTc];
          const float2 kt2 = make_float2(kg.x, kg.y);
#pragma unroll
          for (int j = 0; j < PF_FU; ++j) {
            const float2 x2 = __fmul2_rn(kt2, pinv[j]);
            acc[br][tp] = __ffma2_rn(ev[j], pf_exp2_pair(x2), acc[br][tp]);
          }
        }
      } else {
#pragma unroll
        for (int tp = 0; tp < TP; ++tp)
#pragma unroll
          for (int j = 0; j < PF_FU; ++j) acc[br][tp] = __ffma2_rn(ev[j], u[j][tp], acc[br][tp]);
      }
    }
    JCLK_TOC(clk_b, tb0);
  };
  auto flush = [&]() {
    // the accumulated products are e' * exp(-k tau / PF) with the second factor in [0, 1]: bound = the sum of the weights
    const double fxs = fx_scale(1.f, mode, a.F);
    unsigned long long* fxp = a.fx + (size_t)(blockIdx.x % FX_REP) * (a.part_n - 4);
    if (ok) {
#pragma unroll
      for (int br = 0; br < 2; ++br)
#pragma unroll
        for (int tp = 0; tp < TP; ++tp) {
          // integer atomics: exact in any order
          if (2 * tp < T) fx_add(fxp + (size_t)(br * T + 2 * tp) * R + rc, acc[br][tp].x, fxs);
          if (2 * tp + 1 < T) fx_add(fxp + (size_t)(br * T + 2 * tp + 1) * R + rc, acc[br][tp].y, fxs);
          acc[br][tp] = make_float2(0.f, 0.f);
        }
    }
  };

  float2 U0[PF_FU][TP], U1[PF_FU][TP];
  phase_a_pf(0, U0);
  int since = 0;
  for (int it = 0; it < n; ++it) {
    phase_a_pf(1, U1);
    phase_b_pf(0, U0);
    if (it + 1 < n) phase_a_pf(0, U0);
    phase_b_pf(1, U1);
    if (++since >= FLUSH_TILES) {
      since = 0;
      flush();
    }
  }
  flush();
  if (blockIdx.x == 0 && tid == 0) {
    JCLK_OUT(a.dbg, 56, clk_ring);
    JCLK_OUT(a.dbg, 57, clk_a);
    JCLK_OUT(a.dbg, 58, clk_e);
    JCLK_OUT(a.dbg, 59, clk_b);
  }
}

template <int RPT, int FMT>
__global__ void __launch_bounds__(PASS_THREADS_MAX + 32, RPT == 1 ? 2 : 1) bv_pass_kernel(const PassArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int R = a.R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NTc = blockDim.x - 32, NWc = NTc >> 5;  // consumer threads / warps
  const int nbar = blockDim.x;                        // every named barrier counts all threads
  const bool is_adam = warp == NWc;

  unsigned char* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * a.stage_stride);
  float* s_part = reinterpret_cast<float*>(bars + MAX_STAGES);  // [2][16 warps][8 frames]
  float* s_e = s_part + 2 * 16 * FT;                            // [2][2 branches][8 frames], 16-byte aligned

  const int mode = a.mode;

  // static, deterministic tile partition
#ifdef JX_REVERSE_PARTITION  // experiment: does the straggle follow the CTA index or the feature region?
  const long long G = gridDim.x, cta = gridDim.x - 1 - blockIdx.x;
#else
  const long long G = gridDim.x, cta = blockIdx.x;
#endif
  const int S = a.sub_tiles;  // 8-frame tiles per bulk-copy stage
  const long long s0 = a.n_stage_units * cta / G, s1 = a.n_stage_units * (cta + 1) / G;
  const int ns = (int)(s1 - s0);        // stages of this CTA
  const long long t0 = s0 * S;          // first tile
  const int n = ns * S;                 // tiles of this CTA (frames past F are zero padding)
  const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(a.packed);

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();

  // The first tiles only need the (constant) features: start them before waiting on the previous kernel (PDL).
  if (is_adam && lane == 0) {
    const int pre = ns < a.stages ? ns : a.stages;
    for (int s = 0; s < pre; ++s) {
      const uint32_t bar = smem_u32(&bars[s]);
      mbar_expect_tx(bar, a.tile_bytes);
      bulk_g2s(smem_u32(ring + (size_t)s * a.stage_stride), gsrc + (size_t)(s0 + s) * a.tile_bytes, a.tile_bytes, bar);
    }
  }
  if (blockIdx.x == 0 && tid == blockDim.x - 32) JPROF(a.dbg, 40);
  if (tid == blockDim.x - 32) JPROF(a.dbg, 64 + 4 * blockIdx.x);
  pdl_wait();               // everything the tail kernel wrote (ctl, cc/ch, w, kappa, klsum) is visible from here on
  pdl_launch_dependents();  // lets the tail kernel's launch and constant staging overlap this kernel
  const unsigned ep = *a.epoch;
  if (blockIdx.x == 0 && tid == blockDim.x - 32) JPROF(a.dbg, 41);
  if (tid == blockDim.x - 32) JPROF(a.dbg, 64 + 4 * blockIdx.x + 1);
#ifdef JX_PROFILE
  if (tid == blockDim.x - 32) {  // which SM this CTA runs on (per-CTA timelines: is the end-of-pass straggle tied to the SM?)
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    reinterpret_cast<long long*>(a.dbg)[64 + 4 * blockIdx.x + 3] = (long long)smid;
  }
#endif

  if (a.ctl[ep & 1u].trk_stop) {
    // the on-device loop has ended: drain the prefetched tiles and leave everything untouched
    if (is_adam && lane == 0) {
      const int pre = ns < a.stages ? ns : a.stages;
      for (int s = 0; s < pre; ++s) mbar_wait(smem_u32(&bars[s]), 0);
    }
    return;
  }

  if (is_adam) {
    // ================================================================== Adam warp
    const Ctl* __restrict__ ctl = a.ctl + (ep & 1u);
    // MaxEnt sums of the tail that preceded this pass (all ranks; spins on the peers' words when frame-sharded)
    // hbar as an unevaluated fp32 pair: h_f - hbar is evaluated as (H + (k_hi - hb_hi)) + (k_lo - hb_lo), which keeps the
    // O(lambda) cancellation between kappa_f and hbar exact (Sterbenz) without fp64 instructions or conversions on this warp
    const double hbar_d = mode ? pass_hbar(a, ep, ctl, lane, blockIdx.x == 0) : 0.0;
    const float hb_hi = (float)hbar_d, hb_lo = (float)(hbar_d - (double)hb_hi);
    if (blockIdx.x == 0 && lane == 0) JPROF(a.dbg, 42);
    const int in_set = ep & 1u, out_set = in_set ^ 1;
    const int in_br = ctl->branch;
    const float* __restrict__ zin = a.st.z[in_set][in_br];
    const float* __restrict__ min_ = a.st.m[in_set][in_br];
    const float* __restrict__ vin = a.st.v[in_set][in_br];
    const float* __restrict__ gprev = a.st.g[in_set];
    float* __restrict__ gout = a.st.g[out_set];
    float* __restrict__ zA = a.st.z[out_set][0];
    float* __restrict__ mA = a.st.m[out_set][0];
    float* __restrict__ vA = a.st.v[out_set][0];
    float* __restrict__ zB = a.st.z[out_set][1];
    float* __restrict__ mB = a.st.m[out_set][1];
    float* __restrict__ vB = a.st.v[out_set][1];
    const float lse = ctl->lse;
    const float lrA = ctl->lrA, lrB = ctl->lrB;
    const float maskf = ctl->mask_frame;
    const float bc1 = ctl->bc1, bc2 = ctl->bc2;
    const int have_prev = ctl->have_prev;
    const float b1 = a.b1, b2 = a.b2, om_b1 = a.om_b1, om_b2 = a.om_b2, eps = a.eps, clip = a.clip;
    const int fj = lane & 7, fk = lane >> 3;  // frame within tile, partial-sum group
    double sumA = 0.0, sumB = 0.0, gdot = 0.0;

    int slot = 0, sub = 0, stage_it = 0;
    for (int it = 0; it < n; ++it) {
      const int b = it & 1;
      const long long f = (t0 + it) * FT + fj;
      const bool valid = (fk == 0) && (f < a.F);
      float z_ = 0.f, m_ = 0.f, v_ = 0.f, w_ = 0.f, k_ = 0.f, kl_ = 0.f, gp_ = 0.f;
      if (valid) {  // prefetch before waiting on the consumers
        z_ = zin[f];
        m_ = min_[f];
        v_ = vin[f];
        if (mode) {
          w_ = a.w[f];
          k_ = a.kappa[f];
          kl_ = a.kappa_lo[f];
          gp_ = gprev[f];
        }
      }
      bar_sync(BAR_PART0 + b, nbar);  // partials(it) visible, every consumer has ring[slot] in registers
      if (++sub == S) {  // last sub-tile of the stage is in the consumers' registers: refill the slot
        sub = 0;
        if (lane == 0 && stage_it + a.stages < ns) {
          const uint32_t bar = smem_u32(&bars[slot]);
          mbar_expect_tx(bar, a.tile_bytes);
          bulk_g2s(smem_u32(ring + (size_t)slot * a.stage_stride), gsrc + (size_t)(s0 + stage_it + a.stages) * a.tile_bytes,
                   a.tile_bytes, bar);
        }
        ++stage_it;
        if (++slot == a.stages) slot = 0;
      }
      // H_f: 4 lanes per frame sum a quarter of the warps each, fixed order
      float H = 0.f;
      const float* sp = s_part + b * 16 * FT;
      for (int wv = fk; wv < NWc; wv += 4) H += sp[wv * FT + fj];
      H += __shfl_xor_sync(0xffffffffu, H, 8);
      H += __shfl_xor_sync(0xffffffffu, H, 16);
      float eA = 0.f, eB = 0.f;
      if (valid) {
        if (mode) {
          const float gf = maskf * (w_ * ((H + (k_ - hb_hi)) + (kl_ - hb_lo)));
          const float gpv = have_prev ? gp_ : gf;
          gdot += (double)(gpv * gf);
          gout[f] = gf;
          float sA = gf * lrA, sB = gf * lrB;
          if (clip > 0.f) {
            sA = fminf(fmaxf(sA, -clip), clip);
            sB = fminf(fmaxf(sB, -clip), clip);
          }
          const float m1 = om_b1 * sA + b1 * m_;
          const float m2 = om_b1 * sB + b1 * m_;
          const float v1 = om_b2 * (sA * sA) + b2 * v_;
          const float v2 = om_b2 * (sB * sB) + b2 * v_;
          const float z1 = z_ - (m1 / bc1) / (sqrtf(v1 / bc2) + eps);
          const float z2 = z_ - (m2 / bc1) / (sqrtf(v2 / bc2) + eps);
          eA = expf(z1 - lse);
          eB = expf(z2 - lse);
          zA[f] = z1;
          mA[f] = m1;
          vA[f] = v1;
          zB[f] = z2;
          mB[f] = m2;
          vB[f] = v2;
        } else {
          eA = eB = expf(z_ - lse);
          zA[f] = z_;
          mA[f] = m_;
          vA[f] = v_;
        }
        sumA += (double)eA;
        sumB += (double)eB;
      }
      if (fk == 0) {
        s_e[b * 2 * FT + fj] = eA;
        s_e[b * 2 * FT + FT + fj] = eB;
      }
      bar_arrive(BAR_E0 + b, nbar);  // e'(it) visible to the consumers
    }
    for (int o = 4; o; o >>= 1) {
      sumA += __shfl_xor_sync(0xffffffffu, sumA, o);
      sumB += __shfl_xor_sync(0xffffffffu, sumB, o);
      gdot += __shfl_xor_sync(0xffffffffu, gdot, o);
    }
    if (blockIdx.x == 0 && lane == 0) JPROF(a.dbg, 43);
    if (lane == 0) {
      double* sc = a.cta_scal + (size_t)blockIdx.x * 4;
      sc[0] = sumA;
      sc[1] = sumB;
      sc[2] = gdot;
    }
    return;
  }

  // ==================================================================== consumers
  float ccr[RPT], chr[RPT], nkr[RPT];
  uint32_t rowoff[RPT];
  int row[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    row[i] = tid + i * NTc;
    const bool ok = row[i] < R;
    const int rc = ok ? row[i] : R - 1;  // clamp: finite data, zero coefficients, result never stored
    rowoff[i] = (uint32_t)rc * (FMT == 0 ? 16u : 8u);
    ccr[i] = (ok && mode) ? a.cc[rc] : 0.f;
    chr[i] = (ok && mode) ? a.ch[rc] : 0.f;
    nkr[i] = (ok && mode) ? -a.ck[rc] : 0.f;
  }
  const uint32_t rstride = (uint32_t)R * (FMT == 0 ? 16u : 8u);
  const int jred = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  float acc[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  auto flush = [&]() {  // rare (every FLUSH_TILES tiles): scale and destination are re-derived here rather than kept in registers
    const double fxs = fx_scale(*a.nmax, mode, a.F);
    unsigned long long* fxp = a.fx + (size_t)(blockIdx.x % FX_REP) * (a.part_n - 4);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      if (row[i] < R) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          fx_add(fxp + k * R + row[i], acc[i][k], fxs);  // integer atomics: exact in any order
          acc[i][k] = 0.f;
        }
      }
    }
  };

  int slot = 0, sub = 0;
  uint32_t parity = 0;
  auto fetch = [&](TileRegs<RPT>& D) {
    if (sub == 0) mbar_wait(smem_u32(&bars[slot]), parity);
    load_tile<RPT, FMT>(D, ring + (size_t)slot * a.stage_stride + (size_t)sub * a.subtile_bytes, rstride, rowoff);
    if (++sub == S) {
      sub = 0;
      if (++slot == a.stages) {
        slot = 0;
        parity ^= 1u;
      }
    }
  };
  float* spw0 = s_part + warp * FT;
  float* spw1 = s_part + 16 * FT + warp * FT;
  const float* se0 = s_e;
  const float* se1 = s_e + 2 * FT;

  TileRegs<RPT> D0, D1;
  fetch(D0);
  phase_a<RPT>(D0, ccr, chr, nkr, spw0, lane, jred);
  bar_arrive(BAR_PART0, nbar);
  int it = 1, since = 0;
  while (true) {
    if (it >= n) {
      bar_sync(BAR_E0, nbar);
      phase_b<RPT>(D0, se0, acc);
      break;
    }
    fetch(D1);
    phase_a<RPT>(D1, ccr, chr, nkr, spw1, lane, jred);
    bar_arrive(BAR_PART0 + 1, nbar);
    bar_sync(BAR_E0, nbar);
    phase_b<RPT>(D0, se0, acc);
    ++it;
    if (it >= n) {
      bar_sync(BAR_E0 + 1, nbar);
      phase_b<RPT>(D1, se1, acc);
      break;
    }
    fetch(D0);
    phase_a<RPT>(D0, ccr, chr, nkr, spw0, lane, jred);
    bar_arrive(BAR_PART0, nbar);
    bar_sync(BAR_E0 + 1, nbar);
    phase_b<RPT>(D1, se1, acc);
    ++it;
    since += 2;
    if (since >= FLUSH_TILES) {
      flush();
      since = 0;
    }
  }
  flush();
#ifdef JX_PROFILE
  if (tid == 0) atomicMax(reinterpret_cast<long long*>(a.dbg) + 44, gtime());  // latest CTA end of the run so far
  if (tid == 0) JPROF(a.dbg, 64 + 4 * blockIdx.x + 2);
#endif
}

// ------------------------------------------------------------------------------------------ per-frame mode: Adam warps, kernel
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr int BAR_ADAM = 5;  // the two Adam warps of the per-frame kernel, once, at the end

// One invocation of the Adam code is a ~1500-cycle dependent chain on a single warp whatever the number of frames it carries
// (measured with clock64: it, not the special-function unit, bounded the first version of this mode).  A quad has to be
// turned around in ~1150 cycles, so: two Adam warps (warp q owns quad q of every tile), 2-ulp reciprocal / square root /
// exponential instead of the IEEE sequences, and the per-frame state fetched 8 tiles ahead, one frame per lane.
__device__ __forceinline__ void pf_adam_warp(const PassArgs& a, const Ctl* __restrict__ ctl, unsigned ep, int q, int lane, int NWc, int nbar,
                                             unsigned char* ring, uint64_t* bars, const float4* s_rows, float* s_e, double* s_x,
                                             long long t0, long long s0, int ns, int n) {
  const int mode = a.mode;
  const double hbar_d = mode ? pass_hbar(a, ep, ctl, lane, blockIdx.x == 0 && q == 0) : 0.0;
  const float hb_hi = (float)hbar_d, hb_lo = (float)(hbar_d - (double)hb_hi);  // see bv_pass_kernel's Adam warp
  const int in_set = ep & 1u, out_set = in_set ^ 1;
  const int in_br = ctl->branch;
  const float* __restrict__ zin = a.st.z[in_set][in_br];
  const float* __restrict__ min_ = a.st.m[in_set][in_br];
  const float* __restrict__ vin = a.st.v[in_set][in_br];
  const float* __restrict__ gprev = a.st.g[in_set];
  float* __restrict__ gout = a.st.g[out_set];
  float* __restrict__ zA = a.st.z[out_set][0];
  float* __restrict__ mA = a.st.m[out_set][0];
  float* __restrict__ vA = a.st.v[out_set][0];
  float* __restrict__ zB = a.st.z[out_set][1];
  float* __restrict__ mB = a.st.m[out_set][1];
  float* __restrict__ vB = a.st.v[out_set][1];
  const float nlse2 = -1.4426950408889634f * ctl->lse;
  const float lrA = ctl->lrA, lrB = ctl->lrB;
  const float maskf = ctl->mask_frame;
  const float rbc1 = 1.f / ctl->bc1, rbc2 = 1.f / ctl->bc2;
  const int have_prev = ctl->have_prev;
  const float b1 = a.b1, b2 = a.b2, om_b1 = a.om_b1, om_b2 = a.om_b2, eps = a.eps, clip = a.clip;
  const int fk = lane >> 2;                                   // prefetch batch: lane (fk, lane & 3) holds frame lane & 3 of tile i0 + fk
  const int fj = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);   // frame this lane ends up with after the transposing reduction
  const bool lead = (lane & 7) == 0;                          // lanes 0, 8, 16, 24 carry frames 0..3
  const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(a.packed);
  double sumA = 0.0, sumB = 0.0, gdot = 0.0;
  JCLK_DECL(clk_wait);
  JCLK_DECL(clk_work);
  JCLK_DECL(clk_s0);
  JCLK_DECL(clk_s1);
  JCLK_DECL(clk_s2);
  JCLK_DECL(clk_s3);

  // lane (k, fj) holds frame fj of quad q of tile i0 + k
  float cz, cm, cv, cw, ck, cl, cg, nz, nm, nv, nw, nk, nl, ng;  // (ck, cl): kappa as an unevaluated fp32 pair (frame_terms, bv_tail.cu)
  auto load_batch = [&](int i0, float& z, float& m, float& v, float& w, float& k, float& kl, float& g) {
    const long long f = (t0 + i0 + fk) * FT + q * PF_FU + (lane & 3);
    z = m = v = w = k = kl = g = 0.f;
    if (i0 + fk < n && f < a.F) {
      z = zin[f];
      m = min_[f];
      v = vin[f];
      if (mode) {
        w = a.w[f];
        k = a.kappa[f];
        kl = a.kappa_lo[f];
        g = gprev[f];
      }
    }
  };
  load_batch(0, cz, cm, cv, cw, ck, cl, cg);
  load_batch(8, nz, nm, nv, nw, nk, nl, ng);
  const float4* sr = s_rows + q * (NWc * 32);
  int ring_slot = 0;
  for (int i = 0; i < n; ++i) {
    const int kb = i & 7;
    if (kb == 0 && i > 0) {
      cz = nz, cm = nm, cv = nv, cw = nw, ck = nk, cl = nl, cg = ng;
      load_batch(i + 8, nz, nm, nv, nw, nk, nl, ng);
    }
    const int src = kb * PF_FU + fj;
    const float z_ = __shfl_sync(0xffffffffu, cz, src), m_ = __shfl_sync(0xffffffffu, cm, src), v_ = __shfl_sync(0xffffffffu, cv, src);
    const float w_ = __shfl_sync(0xffffffffu, cw, src), k_ = __shfl_sync(0xffffffffu, ck, src), gp_ = __shfl_sync(0xffffffffu, cg, src);
    const float kl_ = __shfl_sync(0xffffffffu, cl, src);
    const long long f = (t0 + i) * FT + q * PF_FU + fj;
    const bool valid = lead && (f < a.F);
    JCLK_TIC(tw0);
    bar_sync(BAR_PART0 + q, nbar);  // partials of this quad visible, every consumer has it in registers
    JCLK_TOC(clk_wait, tw0);
    JCLK_TIC(tc0);
    if (q == 1) {  // second quad: the tile's slot is free (one tile per stage in this mode)
      const int slot = ring_slot;
      if (++ring_slot == a.stages) ring_slot = 0;
      if (lane == 0 && i + a.stages < ns) {
        const uint32_t bar = smem_u32(&bars[slot]);
        mbar_expect_tx(bar, a.tile_bytes);
        bulk_g2s(smem_u32(ring + (size_t)slot * a.stage_stride), gsrc + (size_t)(s0 + i + a.stages) * a.tile_bytes, a.tile_bytes, bar);
      }
    }
    // H_f = sum over the rows of the consumers' partials: lane-strided, then a transposing butterfly (fixed order)
    float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int w0 = 0; w0 < NWc; w0 += 8) {  // 8 loads in flight: one trip through the shared-memory queue per 256 rows
      float4 t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = (w0 + k < NWc) ? sr[(w0 + k) * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) h4.x += t[k].x, h4.y += t[k].y, h4.z += t[k].z, h4.w += t[k].w;
    }
    JCLK_TOC(clk_s0, tc0);
    float H;
    {
      const bool hi = lane & 16, b3 = lane & 8;
      const float q0 = (hi ? h4.z : h4.x) + __shfl_xor_sync(0xffffffffu, hi ? h4.x : h4.z, 16);
      const float q1 = (hi ? h4.w : h4.y) + __shfl_xor_sync(0xffffffffu, hi ? h4.y : h4.w, 16);
      H = (b3 ? q1 : q0) + __shfl_xor_sync(0xffffffffu, b3 ? q0 : q1, 8);
      H += __shfl_xor_sync(0xffffffffu, H, 4);
      H += __shfl_xor_sync(0xffffffffu, H, 2);
      H += __shfl_xor_sync(0xffffffffu, H, 1);
    }
    JCLK_TOC(clk_s1, tc0);
    // critical path first: e' of both branches -> shared memory -> barrier; the global stores and the fp64 sums follow
    float eA = 0.f, eB = 0.f, gf = 0.f, m1 = 0.f, m2 = 0.f, v1 = 0.f, v2 = 0.f, z1 = 0.f, z2 = 0.f;
    if (valid) {
      if (mode) {
        gf = maskf * (w_ * ((H + (k_ - hb_hi)) + (kl_ - hb_lo)));
        float sA = gf * lrA, sB = gf * lrB;
        if (clip > 0.f) {
          sA = fminf(fmaxf(sA, -clip), clip);
          sB = fminf(fmaxf(sB, -clip), clip);
        }
        m1 = om_b1 * sA + b1 * m_;
        m2 = om_b1 * sB + b1 * m_;
        v1 = om_b2 * (sA * sA) + b2 * v_;
        v2 = om_b2 * (sB * sB) + b2 * v_;
        z1 = z_ - (m1 * rbc1) * rcp_approx(sqrt_approx(v1 * rbc2) + eps);
        z2 = z_ - (m2 * rbc1) * rcp_approx(sqrt_approx(v2 * rbc2) + eps);
        eA = exp2_fma(fmaf(1.4426950408889634f, z1, nlse2));
        eB = exp2_fma(fmaf(1.4426950408889634f, z2, nlse2));
      } else {
        eA = eB = exp2_fma(fmaf(1.4426950408889634f, z_, nlse2));
      }
    }
    JCLK_TOC(clk_s2, tc0);
    if (lead) {
      s_e[q * 2 * PF_FU + fj] = eA;
      s_e[q * 2 * PF_FU + PF_FU + fj] = eB;
    }
    bar_arrive(BAR_E0 + q, nbar);  // e' of this quad visible to the consumers
    JCLK_TOC(clk_s3, tc0);
    if (valid) {
      if (mode) {
        const float gpv = have_prev ? gp_ : gf;
        gdot += (double)(gpv * gf);
        gout[f] = gf;
        zA[f] = z1;
        mA[f] = m1;
        vA[f] = v1;
        zB[f] = z2;
        mB[f] = m2;
        vB[f] = v2;
      } else {
        zA[f] = z_;
        mA[f] = m_;
        vA[f] = v_;
      }
      sumA += (double)eA;
      sumB += (double)eB;
    }
    JCLK_TOC(clk_work, tc0);
  }
  if (blockIdx.x == 0 && lane == 0) {
    JCLK_OUT(a.dbg, 48 + 3 * q, clk_wait);
    JCLK_OUT(a.dbg, 49 + 3 * q, clk_work);
    JCLK_OUT(a.dbg, 50 + 3 * q, (long long)n);
    if (q == 0) {
      JCLK_OUT(a.dbg, 63, clk_s0);
      JCLK_OUT(a.dbg, 60, clk_s1);
      JCLK_OUT(a.dbg, 61, clk_s2);
      JCLK_OUT(a.dbg, 62, clk_s3);
    }
  }
  for (int o = 16; o >= 8; o >>= 1) {  // lanes 0, 8, 16, 24 carry the sums
    sumA += __shfl_xor_sync(0xffffffffu, sumA, o);
    sumB += __shfl_xor_sync(0xffffffffu, sumB, o);
    gdot += __shfl_xor_sync(0xffffffffu, gdot, o);
  }
  if (q == 1 && lane == 0) {
    s_x[0] = sumA;
    s_x[1] = sumB;
    s_x[2] = gdot;
  }
  bar_sync(BAR_ADAM, 64);
  if (q == 0) {
    if (lane == 0) {
      double* sc = a.cta_scal + (size_t)blockIdx.x * 4;
      sc[0] = sumA + s_x[0];
      sc[1] = sumB + s_x[1];
      sc[2] = gdot + s_x[2];
    }
  }
}

// 16 consumer warps + 2 Adam warps: 5 warps on two of the four sub-partitions, i.e. at most 96 registers per thread
__global__ void __launch_bounds__(PASS_THREADS_MAX + 64, 1) bv_pass_kernel_pf(const PassArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NTc = blockDim.x - 64, NWc = NTc >> 5;  // consumer threads / warps
  const int nbar = NTc + 32;                          // a named barrier joins the consumers and ONE Adam warp
  const bool is_adam = warp >= NWc;
  const int q = warp - NWc;

  unsigned char* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * a.stage_stride);
  float* s_part = reinterpret_cast<float*>(bars + MAX_STAGES);            // (main-mode layout kept) here: 3 doubles of Adam warp 1
  double* s_x = reinterpret_cast<double*>(s_part);
  float* s_e = s_part + 2 * 16 * FT;                                      // [2 quads][2 branches][4 frames]
  float4* s_tab = reinterpret_cast<float4*>(s_e + 2 * 2 * FT);            // per-thread slots: [4 pairs of time points][512 threads]
  float4* s_rows = s_tab + (PF_TMAX / 2) * PASS_THREADS_MAX;              // [2 quads][consumer threads] per-row partials of H

  // static, deterministic tile partition (one tile per bulk-copy stage in this mode)
  const long long G = gridDim.x, cta = blockIdx.x;
  const long long s0 = a.n_stage_units * cta / G, s1 = a.n_stage_units * (cta + 1) / G;
  const int ns = (int)(s1 - s0);
  const long long t0 = s0;
  const int n = ns;
  const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(a.packed);

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (is_adam && q == 0 && lane == 0) {  // the first tiles only need the (constant) features: start them before the PDL wait
    const int pre = ns < a.stages ? ns : a.stages;
    for (int s = 0; s < pre; ++s) {
      const uint32_t bar = smem_u32(&bars[s]);
      mbar_expect_tx(bar, a.tile_bytes);
      bulk_g2s(smem_u32(ring + (size_t)s * a.stage_stride), gsrc + (size_t)(s0 + s) * a.tile_bytes, a.tile_bytes, bar);
    }
  }
  if (blockIdx.x == 0 && tid == NTc) JPROF(a.dbg, 40);
  pdl_wait();
  pdl_launch_dependents();
  const unsigned ep = *a.epoch;
  if (blockIdx.x == 0 && tid == NTc) JPROF(a.dbg, 41);

  const Ctl* __restrict__ ctl = a.ctl + (ep & 1u);
  if (ctl->trk_stop) {  // the on-device loop has ended: drain the prefetched tiles and leave everything untouched
    if (is_adam && q == 0 && lane == 0) {
      const int pre = ns < a.stages ? ns : a.stages;
      for (int s = 0; s < pre; ++s) mbar_wait(smem_u32(&bars[s]), 0);
    }
    return;
  }
  const bool db = a.pf_db && a.mode;
  __shared__ SpecScalars s_spec;
  if (db) {  // the BV parameters after the update this sweep speculates on, for both plateau branches (what the tail will select from)
    if (warp == 0) derive_spec(&s_spec, ctl, a.cfg);
    __syncthreads();
  }
  if (is_adam) pf_adam_warp(a, ctl, ep, q, lane, NWc, nbar, ring, bars, s_rows, s_e, s_x, t0, s0, ns, n);
  else if (db) {
    if (a.pf_T <= 4) pf_consumer<4, true>(a, ring, bars, s_rows, s_e, s_tab, n, a.mode, ctl->bc, ctl->bh, NTc, nbar, s_spec.bc[0], s_spec.bh[0], s_spec.bc[1], s_spec.bh[1]);
    else pf_consumer<PF_TMAX, true>(a, ring, bars, s_rows, s_e, s_tab, n, a.mode, ctl->bc, ctl->bh, NTc, nbar, s_spec.bc[0], s_spec.bh[0], s_spec.bc[1], s_spec.bh[1]);
  } else if (a.pf_T <= 4) pf_consumer<4, false>(a, ring, bars, s_rows, s_e, s_tab, n, a.mode, ctl->bc, ctl->bh, NTc, nbar);
  else pf_consumer<PF_TMAX, false>(a, ring, bars, s_rows, s_e, s_tab, n, a.mode, ctl->bc, ctl->bh, NTc, nbar);
}

// ------------------------------------------------------------------------------------------ per-frame mode: d loss / d (bc, bh)
// With average_first = False the BV parameters act inside the frame sum:
//   d<u>[t,r]/d bc = sum_f w_f (-x e^{-x}) Nc[r,f],  x = k_r tau_t exp(-(bc Nc[r,f] + bh Nh[r,f]))      (oracle loss_and_grads)
//   d loss / d bc  = sum_f w_f sum_r Nc[r,f] D[r,f],  D[r,f] = sum_t G[t,r] (-x e^{-x})
// which needs G = dL/d<u> of the CURRENT state (the tail's output) and has to be complete before the update the next sweep
// applies: a sweep of its own over the packed features, between the tail and the pass.  Thread = residue row, quads of 4
// frames read straight from global memory (consecutive rows are consecutive 16-byte words); fp32 per-thread sums are folded
// into fp64 every 32 quads; one (bc, bh) partial per CTA, summed in a fixed order by pf_dbeta_fold_kernel.
constexpr int PF_DB_CTAS = PF_DB_MAX_CTAS;
__device__ __forceinline__ double wsum_d(double v) {  // fixed-shape shuffle tree
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int TM>
__global__ void __launch_bounds__(PASS_THREADS_MAX) pf_dbeta_kernel(const PassArgs a, double* __restrict__ part) {
  __shared__ double s_red[2][PASS_THREADS_MAX / 32];
  const int R = a.R, T = a.pf_T, tid = threadIdx.x;
  const unsigned ep = *a.epoch;
  const Ctl* __restrict__ ctl = a.ctl + (ep & 1u);
  const bool ok = tid < R;
  const int rc = ok ? tid : R - 1;
  float kt[TM], G[TM];
  {
    const float kr = -1.4426950408889634f * a.pf_k[rc];
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      kt[t] = t < T ? kr * a.pf_tp[t] : 0.f;
      G[t] = (ok && t < T) ? a.pf_G[(size_t)t * R + rc] : 0.f;
    }
  }
  const float nl2e_bc = -1.4426950408889634f * ctl->bc, nl2e_bh = -1.4426950408889634f * ctl->bh;
  const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(a.packed);
  const size_t rstride = (size_t)R * 16u, rowoff = (size_t)rc * 16u;
  const long long n_quads = 2 * a.n_tiles;
  double dc = 0.0, dh = 0.0;
  float fc = 0.f, fh = 0.f;
  int since = 0;
  for (long long qd = blockIdx.x; qd < n_quads; qd += gridDim.x) {
    const unsigned char* sub = gsrc + (size_t)(qd >> 1) * a.tile_bytes + ((qd & 1) ? 2 * rstride : 0) + rowoff;
    const float4 c4 = *reinterpret_cast<const float4*>(sub), h4 = *reinterpret_cast<const float4*>(sub + rstride);
    const long long f0 = qd * PF_FU;
    float w[PF_FU];
#pragma unroll
    for (int j = 0; j < PF_FU; ++j) w[j] = (f0 + j < a.F) ? a.w[f0 + j] : 0.f;
    const float nc[PF_FU] = {c4.x, c4.y, c4.z, c4.w}, nh[PF_FU] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
    for (int j = 0; j < PF_FU; ++j) {
      const float pinv = pf_exp2(fmaf(nl2e_bc, nc[j], nl2e_bh * nh[j]));
      float D = 0.f;
#pragma unroll
      for (int t = 0; t < TM; ++t) {
        const float x2 = kt[t] * pinv;        // -log2(e) x
        D = fmaf(G[t] * x2, pf_exp2(x2), D);  // G (-x e^{-x}) log2(e)
      }
      const float wd = w[j] * D;
      fc = fmaf(wd, nc[j], fc);
      fh = fmaf(wd, nh[j], fh);
    }
    if (++since == 32) {
      since = 0;
      dc += (double)fc;
      dh += (double)fh;
      fc = fh = 0.f;
    }
  }
  dc += (double)fc;
  dh += (double)fh;
  if (!ok) dc = dh = 0.0;
  dc = wsum_d(dc);
  dh = wsum_d(dh);
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = dc;
    s_red[1][tid >> 5] = dh;
  }
  __syncthreads();
  if (tid < 2) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_red[tid][k];
    part[(size_t)blockIdx.x * 2 + tid] = t * 0.6931471805599453;  // 1 / log2(e)
  }
}

// fixed-order fold of the CTA partials into the current control block's masked gradients (+ the loss record of this step)
__global__ void pf_dbeta_fold_kernel(const double* __restrict__ part, int n_part, Ctl* ctl, const unsigned* epoch, JaxentLossRecord* records) {
  __shared__ double s_red[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;  // 1024 threads
  double c = 0.0, h = 0.0;
  for (int k = tid; k < n_part; k += blockDim.x) {
    c += part[(size_t)k * 2];
    h += part[(size_t)k * 2 + 1];
  }
  c = wsum_d(c);
  h = wsum_d(h);
  if (lane == 0) {
    s_red[0][warp] = c;
    s_red[1][warp] = h;
  }
  __syncthreads();
  if (tid == 0) {
    Ctl* cb = ctl + (*epoch & 1u);
    if (cb->trk_stop) return;
    double tc = 0.0, th = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      tc += s_red[0][k];
      th += s_red[1][k];
    }
    cb->gb[0] += (float)tc * cb->mask_model;
    cb->gb[1] += (float)th * cb->mask_model;
    JaxentLossRecord* rec = records + cb->rec_idx;
    rec->grad_bc = cb->gb[0];
    rec->grad_bh = cb->gb[1];
  }
}

int launch_pf_dbeta(const PassArgs& a, double* part, Ctl* ctl, JaxentLossRecord* records, cudaStream_t s) {
  const long long n_quads = 2 * a.n_tiles;
  const int grid = (int)(n_quads < PF_DB_CTAS ? n_quads : PF_DB_CTAS);
  const int threads = (a.R + 31) / 32 * 32;
  if (a.pf_T <= 4) pf_dbeta_kernel<4><<<grid, threads, 0, s>>>(a, part);
  else pf_dbeta_kernel<PF_TMAX><<<grid, threads, 0, s>>>(a, part);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) {
    pf_dbeta_fold_kernel<<<1, 1024, 0, s>>>(part, grid, ctl, a.epoch, records);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { set_error("pf_dbeta launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

__global__ void finish_record_kernel(const Ctl* ctl, const unsigned* epoch, const Xchg x, JaxentLossRecord* rec, int n_slots) {
  const unsigned ep = *epoch;
  const Ctl* c = ctl + (ep & 1u);
  double kl[KL_N];
  kl_gather_warp(x, ep, threadIdx.x & 31, kl);
  if (threadIdx.x == 0) finish_record(c, kl[1] + kl[2], rec, n_slots);
}

// ------------------------------------------------------------------------------------------ launchers
int configure_kernels() {
  cudaError_t e;
#define JX_CFG(K, BYTES)                                                                                       \
  e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);                              \
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(" #K "): %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  JX_CFG((bv_pass_kernel<1, 0>), 111 * 1024)
  JX_CFG((bv_pass_kernel<2, 0>), 220 * 1024)
  JX_CFG((bv_pass_kernel<1, 1>), 111 * 1024)
  JX_CFG((bv_pass_kernel<2, 1>), 220 * 1024)
  JX_CFG(bv_pass_kernel_pf, 220 * 1024)
#undef JX_CFG
  return JAXENT_OK;
}

int launch_pass(const PassArgs& a, int threads, int rpt, int ctas, size_t smem, cudaStream_t s) {
  cudaError_t le = cudaSuccess;
  if (a.pf_T > 0) {
    if (rpt != 1 || a.format != 0 || a.pf_T > PF_TMAX) { set_error("per-frame uptake mode needs n_residues <= 512, fp32 features and <= %d time points", PF_TMAX); return JAXENT_E_UNSUPPORTED; }
    le = launch_pdl(bv_pass_kernel_pf, ctas, threads + 32, smem, s, a);  // a second Adam warp
  } else {
    switch (rpt * 2 + (a.format ? 1 : 0)) {
      case 2: le = launch_pdl(bv_pass_kernel<1, 0>, ctas, threads, smem, s, a); break;
      case 3: le = launch_pdl(bv_pass_kernel<1, 1>, ctas, threads, smem, s, a); break;
      case 4: le = launch_pdl(bv_pass_kernel<2, 0>, ctas, threads, smem, s, a); break;
      case 5: le = launch_pdl(bv_pass_kernel<2, 1>, ctas, threads, smem, s, a); break;
      default: set_error("unsupported rows-per-thread %d", rpt); return JAXENT_E_UNSUPPORTED;
    }
  }
  cudaError_t e = (le != cudaSuccess) ? le : cudaGetLastError();
  if (e != cudaSuccess) { set_error("bv_pass_kernel launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

int launch_finish_record(const Ctl* ctl, const unsigned* epoch, const Xchg& x, JaxentLossRecord* rec, int n_slots, cudaStream_t s) {
  finish_record_kernel<<<1, 32, 0, s>>>(ctl, epoch, x, rec, n_slots);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("finish_record launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

}  // namespace jx

// ------------------------------------------------------------------------------------------ C ABI (helpers)
extern "C" size_t jaxent_bv_packed_bytes(int32_t R, int64_t F) {
  if (R <= 0 || F <= 0) return 0;
  const int64_t tiles = (F + jx::FT - 1) / jx::FT;
  return (size_t)tiles * jx::FT * (size_t)R * 2 * sizeof(float);
}

extern "C" int jaxent_bv_pack_features_f32(const float* heavy, const float* acceptor, int32_t R, int64_t F, int64_t ld,
                                           int64_t frame_offset, int64_t F_dst, float* packed, jaxent_stream_t stream) {
  if (!heavy || !acceptor || !packed || R <= 0 || F <= 0 || ld < F || frame_offset < 0 || frame_offset % 4 != 0 ||
      frame_offset + F > F_dst || (frame_offset + F < F_dst && F % 4 != 0)) {
    jx::set_error("pack_features: invalid argument (R=%d F=%lld ld=%lld offset=%lld F_dst=%lld)", R, (long long)F,
                  (long long)ld, (long long)frame_offset, (long long)F_dst);
    return JAXENT_E_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(packed) & 15u) != 0) {
    jx::set_error("pack_features: packed buffer must be 16-byte aligned");
    return JAXENT_E_INVALID;
  }
  const int64_t Fpad = ((F_dst + jx::FT - 1) / jx::FT) * jx::FT;
  const bool last_chunk = frame_offset + F >= F_dst;
  const int64_t f_end = last_chunk ? Fpad - frame_offset : F;
  dim3 grid((unsigned)((f_end + 31) / 32), (unsigned)((R + 31) / 32));
  dim3 block(32, 8);
  jx::pack_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(heavy, acceptor, R, F, ld, frame_offset, F_dst, packed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { jx::set_error("pack_kernel launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" size_t jaxent_bv_packed_bytes_fmt(int32_t R, int64_t F, int32_t format) {
  if (R <= 0 || F <= 0) return 0;
  if (format == 0) return jaxent_bv_packed_bytes(R, F);
  if (format != 1) return 0;
  const int64_t frames_per_stage = (int64_t)jx::FT * jx::U8_SUB_TILES;
  const int64_t stages = (F + frames_per_stage - 1) / frames_per_stage;
  return (size_t)stages * frames_per_stage * (size_t)R * 2;
}

extern "C" int jaxent_bv_features_fit_u8(const float* heavy, const float* acceptor, int32_t R, int64_t F, int64_t ld, int32_t* flag,
                                         jaxent_stream_t stream) {
  if (!heavy || !acceptor || !flag || R <= 0 || F <= 0 || ld < F) { jx::set_error("features_fit_u8: invalid argument"); return JAXENT_E_INVALID; }
  jx::fit_u8_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(heavy, R, F, ld, flag);
  jx::fit_u8_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(acceptor, R, F, ld, flag);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { jx::set_error("fit_u8 launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_pack_features_u8(const float* heavy, const float* acceptor, int32_t R, int64_t F, int64_t ld,
                                          unsigned char* packed, jaxent_stream_t stream) {
  if (!heavy || !acceptor || !packed || R <= 0 || F <= 0 || ld < F) { jx::set_error("pack_features_u8: invalid argument"); return JAXENT_E_INVALID; }
  if ((reinterpret_cast<uintptr_t>(packed) & 15u) != 0) { jx::set_error("pack_features_u8: packed buffer must be 16-byte aligned"); return JAXENT_E_INVALID; }
  const int64_t fps = (int64_t)jx::FT * jx::U8_SUB_TILES;
  const int64_t Fpad = (F + fps - 1) / fps * fps;
  dim3 grid((unsigned)((Fpad + 31) / 32), (unsigned)((R + 31) / 32));
  dim3 block(32, 8);
  jx::pack_u8_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(heavy, acceptor, R, F, ld, Fpad, packed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { jx::set_error("pack_u8_kernel launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_prior_partial(const float* prior, int64_t F, double eps, double* out, jaxent_stream_t stream) {
  if (!prior || !out || F <= 0) { jx::set_error("prior_partial: invalid argument"); return JAXENT_E_INVALID; }
  jx::prior_partial_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(prior, F, eps, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { jx::set_error("prior_partial launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_max(const float* z, int64_t F, float* out, jaxent_stream_t stream) {
  if (!z || !out || F <= 0) { jx::set_error("max: invalid argument"); return JAXENT_E_INVALID; }
  jx::max_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(z, F, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { jx::set_error("max launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}
