// The streaming pass of the BV MaxEnt optimise step for sm_100a, plus its small helpers.
//
// One pass over the packed (frame-quad major) feature arrays does, per tile of FT=8 frames:
//   phase A  column sums  H_f = sum_r cc_r*Nc[r,f] + ch_r*Nh[r,f]        (backward of step k;
//            reference: the transpose products jax.grad emits for utils/jax_fn.py:30-38)
//   Adam     g_f = mask*w_f*(H_f + kappa_f - hbar), grad-dot partial, optax adam for BOTH plateau
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
template <int RPT>
__global__ void __launch_bounds__(PASS_THREADS_MAX, (RPT == 1) ? 2 : 1) bv_pass_kernel(const PassArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int R = a.R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int NT = blockDim.x;

  unsigned char* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * a.stage_stride);
  float* s_part = reinterpret_cast<float*>(bars + MAX_STAGES);  // [16][8]
  float* s_e = s_part + 16 * FT;                                // [2][8], 16-byte aligned

  // ---- per-step scalars (uniform) ----
  const unsigned ep = *a.epoch;
  const Ctl* __restrict__ ctl = a.ctl + (ep & 1u);
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
  const int mode = a.mode;
  const float lse = ctl->lse;
  const float lrA = ctl->lrA, lrB = ctl->lrB;
  const float maskf = ctl->mask_frame;
  const float bc1 = ctl->bc1, bc2 = ctl->bc2;
  const int have_prev = ctl->have_prev;
  const float hbar = (float)(ctl->hbar_data + a.klsum[0]);
  const float b1 = a.b1, b2 = a.b2, om_b1 = a.om_b1, om_b2 = a.om_b2, eps = a.eps, clip = a.clip;

  if (blockIdx.x == 0 && tid == 0 && mode == 1) {
    // fold the (all-reduced) MaxEnt sums into the loss record of the step being differentiated
    finish_record(ctl, a.klsum, a.records, JAXENT_MAX_SLOTS);
  }

  // ---- rows owned by this thread ----
  float ccr[RPT], chr[RPT];
  int row[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    row[i] = tid + i * NT;
    const bool ok = row[i] < R;
    ccr[i] = (ok && mode) ? a.cc[row[i]] : 0.f;
    chr[i] = (ok && mode) ? a.ch[row[i]] : 0.f;
  }

  // ---- static, deterministic tile partition ----
  const long long G = gridDim.x, cta = blockIdx.x;
  const long long t0 = a.n_tiles * cta / G, t1 = a.n_tiles * (cta + 1) / G;
  const int n = (int)(t1 - t0);
  const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(a.packed);

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    const int pre = n < a.stages ? n : a.stages;
    for (int s = 0; s < pre; ++s) {
      const uint32_t bar = smem_u32(&bars[s]);
      mbar_expect_tx(bar, a.tile_bytes);
      bulk_g2s(smem_u32(ring + (size_t)s * a.stage_stride), gsrc + (size_t)(t0 + s) * a.tile_bytes, a.tile_bytes, bar);
    }
  }

  double accAc[RPT], accAh[RPT], accBc[RPT], accBh[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) accAc[i] = accAh[i] = accBc[i] = accBh[i] = 0.0;
  double sumA = 0.0, sumB = 0.0, gdot = 0.0;

  const bool adam_lane = (warp == 0) && (lane < FT);
  // frame slot of this lane after the transposing warp reduction
  const int jred = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);

  int slot = 0;
  uint32_t parity = 0;
  for (int it = 0; it < n; ++it) {
    // -- prefetch the per-frame optimiser state of this tile (consumed after barrier B1)
    const long long f = (t0 + it) * FT + lane;
    const bool valid = adam_lane && (f < a.F);
    float z_ = 0.f, m_ = 0.f, v_ = 0.f, w_ = 0.f, k_ = 0.f, gp_ = 0.f;
    if (valid) {
      z_ = zin[f];
      m_ = min_[f];
      v_ = vin[f];
      if (mode) {
        w_ = a.w[f];
        k_ = a.kappa[f];
        gp_ = gprev[f];
      }
    }

    // -- wait for the tile, move this thread's rows into registers
    mbar_wait(smem_u32(&bars[slot]), parity);
    const float4* __restrict__ tile = reinterpret_cast<const float4*>(ring + (size_t)slot * a.stage_stride);
    float4 nc[RPT][QT], nh[RPT][QT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        if (row[i] < R) {
          nc[i][q] = tile[(q * 2 + 0) * R + row[i]];
          nh[i][q] = tile[(q * 2 + 1) * R + row[i]];
        } else {
          nc[i][q] = make_float4(0.f, 0.f, 0.f, 0.f);
          nh[i][q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }

    // -- phase A: column-sum partials over this thread's rows
    float p[FT];
#pragma unroll
    for (int j = 0; j < FT; ++j) p[j] = 0.f;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        p[4 * q + 0] = fmaf(ccr[i], nc[i][q].x, fmaf(chr[i], nh[i][q].x, p[4 * q + 0]));
        p[4 * q + 1] = fmaf(ccr[i], nc[i][q].y, fmaf(chr[i], nh[i][q].y, p[4 * q + 1]));
        p[4 * q + 2] = fmaf(ccr[i], nc[i][q].z, fmaf(chr[i], nh[i][q].z, p[4 * q + 2]));
        p[4 * q + 3] = fmaf(ccr[i], nc[i][q].w, fmaf(chr[i], nh[i][q].w, p[4 * q + 3]));
      }
    }
    // transposing warp reduction: 8 values x 32 lanes -> lane holds the warp total of frame jred
    {
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
      if ((lane & 3) == 0) s_part[warp * FT + jred] = s;
    }
    __syncthreads();  // B1: partials visible; every thread has consumed ring[slot]

    // -- refill the slot just consumed
    if (tid == 0 && it + a.stages < n) {
      const uint32_t bar = smem_u32(&bars[slot]);
      mbar_expect_tx(bar, a.tile_bytes);
      bulk_g2s(smem_u32(ring + (size_t)slot * a.stage_stride), gsrc + (size_t)(t0 + it + a.stages) * a.tile_bytes,
               a.tile_bytes, bar);
    }

    // -- per-frame optimiser update (8 lanes of warp 0)
    if (adam_lane) {
      float eA = 0.f, eB = 0.f;
      if (valid) {
        if (mode) {
          float H = 0.f;
          for (int wv = 0; wv < NW; ++wv) H += s_part[wv * FT + lane];
          const float gf = maskf * (w_ * ((H + k_) - hbar));
          const float gpv = have_prev ? gp_ : gf;
          gdot += (double)(gpv * gf);
          gout[f] = gf;
          // branch A: base learning rate
          float sA = gf * lrA;
          if (clip > 0.f) sA = fminf(fmaxf(sA, -clip), clip);
          const float m1 = om_b1 * sA + b1 * m_;
          const float v1 = om_b2 * (sA * sA) + b2 * v_;
          const float z1 = z_ - (m1 / bc1) / (sqrtf(v1 / bc2) + eps);
          // branch B: plateau-reduced learning rate
          float sB = gf * lrB;
          if (clip > 0.f) sB = fminf(fmaxf(sB, -clip), clip);
          const float m2 = om_b1 * sB + b1 * m_;
          const float v2 = om_b2 * (sB * sB) + b2 * v_;
          const float z2 = z_ - (m2 / bc1) / (sqrtf(v2 / bc2) + eps);
          zA[f] = z1;
          mA[f] = m1;
          vA[f] = v1;
          zB[f] = z2;
          mB[f] = m2;
          vB[f] = v2;
          eA = expf(z1 - lse);
          eB = expf(z2 - lse);
        } else {
          eA = eB = expf(z_ - lse);
          zA[f] = z_;
          mA[f] = m_;
          vA[f] = v_;
        }
        sumA += (double)eA;
        sumB += (double)eB;
      }
      s_e[lane] = eA;
      s_e[FT + lane] = eB;
    }
    __syncthreads();  // B2: e'_f visible

    // -- phase B: softmax-weighted row sums for both branches
    {
      const float4* e4 = reinterpret_cast<const float4*>(s_e);
      const float4 ea0 = e4[0], ea1 = e4[1], eb0 = e4[2], eb1 = e4[3];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        float tAc = nc[i][0].x * ea0.x;
        tAc = fmaf(nc[i][0].y, ea0.y, tAc);
        tAc = fmaf(nc[i][0].z, ea0.z, tAc);
        tAc = fmaf(nc[i][0].w, ea0.w, tAc);
        tAc = fmaf(nc[i][1].x, ea1.x, tAc);
        tAc = fmaf(nc[i][1].y, ea1.y, tAc);
        tAc = fmaf(nc[i][1].z, ea1.z, tAc);
        tAc = fmaf(nc[i][1].w, ea1.w, tAc);
        float tAh = nh[i][0].x * ea0.x;
        tAh = fmaf(nh[i][0].y, ea0.y, tAh);
        tAh = fmaf(nh[i][0].z, ea0.z, tAh);
        tAh = fmaf(nh[i][0].w, ea0.w, tAh);
        tAh = fmaf(nh[i][1].x, ea1.x, tAh);
        tAh = fmaf(nh[i][1].y, ea1.y, tAh);
        tAh = fmaf(nh[i][1].z, ea1.z, tAh);
        tAh = fmaf(nh[i][1].w, ea1.w, tAh);
        float tBc = nc[i][0].x * eb0.x;
        tBc = fmaf(nc[i][0].y, eb0.y, tBc);
        tBc = fmaf(nc[i][0].z, eb0.z, tBc);
        tBc = fmaf(nc[i][0].w, eb0.w, tBc);
        tBc = fmaf(nc[i][1].x, eb1.x, tBc);
        tBc = fmaf(nc[i][1].y, eb1.y, tBc);
        tBc = fmaf(nc[i][1].z, eb1.z, tBc);
        tBc = fmaf(nc[i][1].w, eb1.w, tBc);
        float tBh = nh[i][0].x * eb0.x;
        tBh = fmaf(nh[i][0].y, eb0.y, tBh);
        tBh = fmaf(nh[i][0].z, eb0.z, tBh);
        tBh = fmaf(nh[i][0].w, eb0.w, tBh);
        tBh = fmaf(nh[i][1].x, eb1.x, tBh);
        tBh = fmaf(nh[i][1].y, eb1.y, tBh);
        tBh = fmaf(nh[i][1].z, eb1.z, tBh);
        tBh = fmaf(nh[i][1].w, eb1.w, tBh);
        accAc[i] += (double)tAc;
        accAh[i] += (double)tAh;
        accBc[i] += (double)tBc;
        accBh[i] += (double)tBh;
      }
    }
    if (++slot == a.stages) {
      slot = 0;
      parity ^= 1u;
    }
  }

  // ---- per-CTA partials (fp64), reduced across CTAs by reduce_kernel in a fixed order ----
  double* part = a.partials + (size_t)blockIdx.x * a.part_n;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    if (row[i] < R) {
      part[0 * R + row[i]] = accAc[i];
      part[1 * R + row[i]] = accAh[i];
      part[2 * R + row[i]] = accBc[i];
      part[3 * R + row[i]] = accBh[i];
    }
  }
  if (warp == 0) {
    for (int o = 4; o; o >>= 1) {
      sumA += __shfl_xor_sync(0xffffffffu, sumA, o);
      sumB += __shfl_xor_sync(0xffffffffu, sumB, o);
      gdot += __shfl_xor_sync(0xffffffffu, gdot, o);
    }
    if (lane == 0) {
      part[4 * R + 0] = sumA;
      part[4 * R + 1] = sumB;
      part[4 * R + 2] = gdot;
      part[4 * R + 3] = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------ reduce
// red[j] = sum_cta partials[cta][j] in CTA order (deterministic); also advances the epoch.
__global__ void __launch_bounds__(1024) reduce_kernel(const ReduceArgs a) {
  // 32 columns x 32 partial-groups per block: every load is independent (latency overlapped) and
  // the summation order is fixed (group-strided, then groups 0..31) -> bitwise reproducible.
  __shared__ double sh[32][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (col < a.part_n) {
#pragma unroll 4
    for (int c = grp; c < a.n_part; c += 32) s += a.partials[(size_t)c * a.part_n + col];
  }
  sh[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && col < a.part_n) {
    double t = 0.0;
#pragma unroll
    for (int g2 = 0; g2 < 32; ++g2) t += sh[g2][lane];
    a.red[col] = t;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.epoch = *a.epoch + 1u;
}

__global__ void finish_record_kernel(const Ctl* ctl, const unsigned* epoch, const double* klsum, JaxentLossRecord* rec,
                                     int n_slots) {
  const Ctl* c = ctl + (*epoch & 1u);
  finish_record(c, klsum, rec, n_slots);
}

// ------------------------------------------------------------------------------------------ launchers
int configure_kernels() {
  cudaError_t e;
  e = cudaFuncSetAttribute(bv_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(pass<1>): %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  e = cudaFuncSetAttribute(bv_pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(pass<2>): %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  e = cudaFuncSetAttribute(bv_pass_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(pass<4>): %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

int launch_pass(const PassArgs& a, int threads, int rpt, int ctas, size_t smem, cudaStream_t s) {
  switch (rpt) {
    case 1: bv_pass_kernel<1><<<ctas, threads, smem, s>>>(a); break;
    case 2: bv_pass_kernel<2><<<ctas, threads, smem, s>>>(a); break;
    case 4: bv_pass_kernel<4><<<ctas, threads, smem, s>>>(a); break;
    default: set_error("unsupported rows-per-thread %d", rpt); return JAXENT_E_UNSUPPORTED;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bv_pass_kernel launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

int launch_reduce(const ReduceArgs& a, cudaStream_t s) {
  const int blocks = (a.part_n + 31) / 32;
  reduce_kernel<<<blocks, 1024, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("reduce_kernel launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

int launch_finish_record(const Ctl* ctl, const unsigned* epoch, const double* klsum, JaxentLossRecord* rec, int n_slots,
                         cudaStream_t s) {
  finish_record_kernel<<<1, 1, 0, s>>>(ctl, epoch, klsum, rec, n_slots);
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
