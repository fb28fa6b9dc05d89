// The R x T tail body shared by the single-ensemble tail kernel (bv_tail.cu) and the batched tail kernel
// (bg_step.cu): ln PF, uptake, peptide maps, data losses and their closed-form backward down to c_r = dL/d lnPF_r.
// Reference symbols: see the header comment of bv_tail.cu.
#pragma once
#include <cmath>

#include "jx_internal.cuh"

namespace jx {

#ifdef JX_PROFILE
#define BPROF(i) do { if (io.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) io.prof[i] = gtime(); } while (0)
#else
#define BPROF(i) do {} while (0)
#endif

// ---- code-size discipline: this kernel runs a long straight-line path once per launch, so its cost is
// dominated by instruction fetch.  Heavy math lives behind single __noinline__ call sites.
static __device__ __noinline__ double wsum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static __device__ __noinline__ double dexp(double x) { return exp(x); }
static __device__ __noinline__ double dlog(double x) { return log(x); }
// fpow: jx_internal.cuh
// Explicit shared-memory loads by 32-bit shared address.  The peptide-map loops select between the train and the validation
// arrays at run time; the compiler then loses the address space and emits GENERIC loads, which cost ~500 cycles per dependent
// iteration on sm_100 (measured: 7600 cycles for a 10-entry row) against ~30 for LDS.
__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_i(uint32_t a) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_d(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_d(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
// floor(x / d) for 0 <= x < 2^21 with inv = 1/d (exact: the fraction is >= 0.5/d away from an integer)
__device__ __forceinline__ int fdiv(int x, float inv) { return (int)(((float)x + 0.5f) * inv); }

// word layout of one data slot's blob == its shared-memory image
__host__ __device__ inline BlobLayout blob_layout(int R, int Tm, int p_tr, int nnz_tr, int p_va, int nnz_va) {
  BlobLayout b;
  int o = 0;
  b.tr_ptr = o;  o += p_tr + 1;
  b.tr_col = o;  o += nnz_tr;
  b.tr_tptr = o; o += R + 1;
  b.tr_trow = o; o += nnz_tr;
  b.va_ptr = o;  o += p_va + 1;
  b.va_col = o;  o += nnz_va;
  b.tr_y = o;    o += p_tr * Tm;
  b.va_y = o;    o += p_va * Tm;
  // the map values are kept as doubles (8-byte aligned): a float -> double conversion is a 16-lane/clk instruction on
  // sm_100 and the loops that read them run on ONE SM
  o = (o + 1) & ~1;
  b.tr_val = o;  o += 2 * nnz_tr;
  b.tr_tval = o; o += 2 * nnz_tr;
  b.va_val = o;  o += 2 * nnz_va;
  b.words = (o + 3) & ~3;
  return b;
}

// The dynamic shared memory of the tail kernels.  Declared at namespace scope and re-derived inside tail_body() so that the
// compiler keeps the shared address space of every pointer (pointers that travel through a struct degrade to generic
// loads, LD.E instead of LDS, in the sparse-map loops).
extern __shared__ __align__(16) unsigned char jx_dyn_smem[];

// shared-memory carve-up of the tail body
struct TailSmem {
  double *slp, *sxe, *sc_rt, *sU, *res_tr, *gp_tr, *res_va, *smean;
  double *s_k, *s_tp;  // intrinsic rates and time points, as doubles (see blob_layout)
  int* s_blob;
};
__device__ __forceinline__ TailSmem carve_tail_smem(unsigned char* smem_raw, int R, int T, int p_max_tr, int p_max_va) {
  const int Tm = T > 0 ? T : 1;
  TailSmem m;
  m.slp = reinterpret_cast<double*>(smem_raw);        // [R] lnPF
  m.sxe = m.slp + R;                                   // [R] exp(-lnPF)
  m.sc_rt = m.sxe + R;                                 // [Tm][R] back-projection items
  m.sU = m.sc_rt + (size_t)Tm * R;                     // [T][R] uptake
  m.res_tr = m.sU + (size_t)T * R;                     // [P_tr * Tm] (GENERAL only)
  m.gp_tr = m.res_tr + (size_t)p_max_tr * Tm;          // [P_tr * Tm] pred -> residual -> dL/dpred
  m.res_va = m.gp_tr + (size_t)p_max_tr * Tm;          // [P_va * Tm]
  m.smean = m.res_va + (size_t)p_max_va * Tm;          // [4 * Tm]
  m.s_k = m.smean + 4 * Tm;                            // [R]
  m.s_tp = m.s_k + R;                                  // [Tm]
  m.s_blob = reinterpret_cast<int*>(smem_raw + ((((size_t)(reinterpret_cast<unsigned char*>(m.s_tp + Tm) - smem_raw)) + 15) & ~(size_t)15));  // 16-byte aligned blob image
  return m;
}

// stages the first data slot's blob, the rates and the time points (constant data: may run before pdl_wait)
__device__ __forceinline__ int stage_tail_constants(const JaxentBvProblem& pb, const BlobSet& blobs, const TailSmem& m) {
  const int R = pb.R, T = pb.uptake ? pb.T : 0;
  const int tid = threadIdx.x, nt = blockDim.x;
  int first_data = -1;
#pragma unroll 1
  for (int i = pb.n_slots - 1; i >= 0; --i)
    if (pb.slots[i].kind <= JAXENT_LOSS_PF_L2) first_data = i;
  if (first_data >= 0) {
    const int4* src = reinterpret_cast<const int4*>(blobs.blob[first_data]);
    int4* dst = reinterpret_cast<int4*>(m.s_blob);
    const int n4 = blobs.words[first_data] >> 2;
#pragma unroll 4
    for (int i = tid; i < n4; i += nt) dst[i] = src[i];
  }
  if (T > 0) {
#pragma unroll 1
    for (int i = tid; i < R; i += nt) m.s_k[i] = (double)pb.k_ints[i];
    if (tid < T) m.s_tp[tid] = (double)pb.timepoints[tid];
  }
  return first_data;
}

struct TailIO {
  float* avg;     // [2R] frame-averaged contacts
  float* logpf;   // [R]
  float* uptake;  // [T*R]
  float* pf_G;    // per-frame uptake mode: [T*R] dL/d<uptake>[t,r] for the next pass (else unused)
  int pf_mode;    // 1: per-frame uptake mode -- the caller has filled sm.sU with the frame-averaged uptake already
  int p_max_tr, p_max_va;  // carve-up parameters (carve_tail_smem)
  float* ck;      // [R] centring constants of the pass (see "centred column sums" in tail_body), or nullptr (batched GEMM path)
  const float* row_off;  // batched tensor-core path: [2][Rp] per-row constants the stored features were centred with (or nullptr)
  int row_off_stride;    // Rp
  long long* prof; // JX_PROFILE builds: %globaltimer stamps of the phases (else unused)
  int c_pieces;   // how the sweep sees the row coefficients: 0 = rounded to fp32, n > 0 = the sum of n bf16 pieces (batched GEMM path)
};

// the value of a row coefficient as the feature sweep will use it (see the hbar comment in tail_body)
__device__ __forceinline__ double coeff_as_used(double x, int pieces) {
  const float xf = (float)x;
  if (pieces <= 0) return (double)xf;
  float r = xf;
  double acc = 0.0;
  for (int i = 0; i < pieces; ++i) {  // round-to-nearest-even bf16 pieces, as split3() forms them (bg_step.cu)
    unsigned u = __float_as_uint(r);
    u += 0x7fffu + ((u >> 16) & 1u);
    const float pc = __uint_as_float(u & 0xffff0000u);
    acc += (double)pc;
    r -= pc;
  }
  return acc;
}

// Whole-CTA function.  In: acc0/acc1 = un-normalised weighted row sums of this thread's residue (tid < R), S = sum of the
// un-normalised weights, (bc, bh), fw/fs = forward_model_weights / scaling of every slot.  Out: s_fin[0..5] train losses,
// [6..11] val losses, [12] sum_r c_r lnPF_r, [13] dL/dbc, [14] dL/dbh (data part); my_c = c_r of this thread's residue;
// gbc_reg / gbh_reg = regulariser gradients.  s_red must be zeroed ([warp][0..15]) by the caller before the call and a
// __syncthreads() must separate that and the staging from the call.
struct NoSide {
  __device__ void operator()() const {}
};

// `side` is called once by the LAST warp at the start of the first data slot's peptide-map phase (a phase that usually keeps
// only the first warps busy): the caller can park latency-tolerant scalar work there.
template <bool GENERAL, class Side = NoSide>
__device__ __forceinline__ void tail_body(const JaxentBvProblem& pb, const BlobSet& blobs, int first_data, const TailSmem& sm,
                                          double acc0, double acc1, double S, double bc, double bh, const float* fw, const float* fs,
                                          const TailIO io, double (*s_red)[16], double* s_fin, double& my_c, double& my_a, double& my_b,
                                          double& my_lp, double& gbc_reg_out, double& gbh_reg_out, Side side = Side()) {
  const int R = pb.R, T = pb.uptake ? pb.T : 0, Tm = T > 0 ? T : 1;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const float invR = 1.0f / (float)R, invTm = 1.0f / (float)Tm;
  const TailSmem smx = carve_tail_smem(jx_dyn_smem, R, T, io.p_max_tr, io.p_max_va);  // == sm, with the address space intact
  (void)sm;
  double *slp = smx.slp, *sxe = smx.sxe, *sc_rt = smx.sc_rt, *sU = smx.sU, *res_tr = smx.res_tr, *gp_tr = smx.gp_tr, *res_va = smx.res_va,
         *smean = smx.smean;
  const double *s_k = smx.s_k, *s_tp = smx.s_tp;
  int* s_blob = smx.s_blob;
  {
    my_a = 0, my_b = 0, my_lp = 0, my_c = 0;
    const bool pfm = io.pf_mode != 0;
    if (!pfm && tid < R) {
      const double invS = __drcp_rn(S);
      my_a = acc0 * invS;
      my_b = acc1 * invS;
      my_lp = bc * my_a + bh * my_b;
      slp[tid] = my_lp;
      io.avg[tid] = (float)my_a;
      io.avg[R + tid] = (float)my_b;
      io.logpf[tid] = (float)my_lp;
      if (T > 0) sxe[tid] = dexp(-my_lp);
    }
    if (T > 0 && !pfm) {
      __syncthreads();  // exp(-lnPF) visible
#pragma unroll 1
      for (int it = tid; it < R * T; it += nt) {  // sU is [r][t] (time fastest): see the gathers of phase (a)
        const int r = fdiv(it, invTm), t = it - r * T;
        const double e = dexp(-(s_k[r] * s_tp[t] * sxe[r]));
        sU[it] = 1.0 - e;
        io.uptake[t * R + r] = (float)(1.0 - e);
      }
    }
    __syncthreads();  // (2) lnPF / uptake visible
    BPROF(4);

    // ---- loss slots
    double gbc_reg = 0.0, gbh_reg = 0.0;
#pragma unroll 1
    for (int i = 0; i < pb.n_slots; ++i) {
      const int kind = pb.slots[i].kind;
      const double scale = (double)fw[i] * (double)fs[i];
      if (kind == JAXENT_LOSS_PARAMS_L1 || kind == JAXENT_LOSS_PARAMS_L2) {
        const bool l1 = kind == JAXENT_LOSS_PARAMS_L1;
        const double dc = bc - (double)pb.slots[i].prior_bc, dh = bh - (double)pb.slots[i].prior_bh;
        if (tid == 0) s_red[0][i] = s_red[0][JAXENT_MAX_SLOTS + i] = l1 ? fabs(dc) + fabs(dh) : dc * dc + dh * dh;
        gbc_reg += scale * (l1 ? (double)((dc > 0) - (dc < 0)) : 2.0 * dc);
        gbh_reg += scale * (l1 ? (double)((dh > 0) - (dh < 0)) : 2.0 * dh);
        continue;
      }
      if (kind > JAXENT_LOSS_PF_L2) continue;  // MaxEnt: handled per frame
      if (i != first_data) {
        __syncthreads();
        const int4* src = reinterpret_cast<const int4*>(blobs.blob[i]);
        int4* dst = reinterpret_cast<int4*>(s_blob);
        const int n4 = blobs.words[i] >> 2;
#pragma unroll 1
        for (int k = tid; k < n4; k += nt) dst[k] = src[k];
        __syncthreads();
      }
      const int Ptr = pb.slots[i].train.n_pep, Pva = pb.slots[i].val.n_pep;
      const BlobLayout bl = blob_layout(R, Tm, Ptr, pb.slots[i].train.nnz, Pva, pb.slots[i].val.nnz);
      const bool pf = kind == JAXENT_LOSS_PF_L2;
      const bool centre = GENERAL && kind == JAXENT_LOSS_UPTAKE_MC_EYE_MSE;
      const bool sigma = GENERAL && kind == JAXENT_LOSS_UPTAKE_SIGMA_MSE;
      const double half = pf ? 1.0 : 0.5;

      // (a) predictions of train and val peptides, one (peptide, time point) item per thread.  This phase is pure latency
      // (a few hundred short dot products on one SM), so the loop is shaped for it: the 4 (index, value, x) triples of a
      // block are loaded before the first product, and only the fused multiply-adds of one accumulator are dependent.
      // The eye / PF kinds finish residual + loss here.
      {
        const int items = (Ptr + Pva) * Tm;
        double pl_tr = 0.0, pl_va = 0.0;
        const uint32_t blob = sm_addr(s_blob);
        const uint32_t xb = sm_addr(pf ? slp : sU);
        BPROF(16);
#pragma unroll 1
        for (int j = tid; j < items; j += nt) {
          const int row = fdiv(j, invTm), t = j - row * Tm;
          const bool va = row >= Ptr;
          const int p = va ? row - Ptr : row;
          const uint32_t a_ptr = blob + 4u * (uint32_t)(va ? bl.va_ptr : bl.tr_ptr), a_col = blob + 4u * (uint32_t)(va ? bl.va_col : bl.tr_col);
          const uint32_t a_val = blob + 4u * (uint32_t)(va ? bl.va_val : bl.tr_val), a_y = blob + 4u * (uint32_t)(va ? bl.va_y : bl.tr_y);
          const uint32_t a_out = sm_addr(va ? res_va : ((centre || sigma) ? res_tr : gp_tr));
          // x[c] = lnPF[c] (PF kind) or uptake[c][t]: the 8 lanes of a peptide read 8 consecutive doubles per entry
          const uint32_t x = xb + (pf ? 0u : 8u * (uint32_t)t), xs = pf ? 8u : 8u * (uint32_t)T;
          const int qb = lds_i(a_ptr + 4u * p), qe = lds_i(a_ptr + 4u * p + 4u);
          const float yv = lds_f(a_y + 4u * (uint32_t)j - (va ? 4u * (uint32_t)(Ptr * Tm) : 0u));  // y[p * Tm + t]
          double pred = 0.0;
          int q = qb;
#pragma unroll 1
          for (; q + 4 <= qe; q += 4) {
            const uint32_t c0 = lds_i(a_col + 4u * q), c1 = lds_i(a_col + 4u * q + 4u), c2 = lds_i(a_col + 4u * q + 8u), c3 = lds_i(a_col + 4u * q + 12u);
            const double v0 = lds_d(a_val + 8u * q), v1 = lds_d(a_val + 8u * q + 8u), v2 = lds_d(a_val + 8u * q + 16u), v3 = lds_d(a_val + 8u * q + 24u);
            const double x0 = lds_d(x + xs * c0), x1 = lds_d(x + xs * c1), x2 = lds_d(x + xs * c2), x3 = lds_d(x + xs * c3);
            pred += v0 * x0;  // same order as a one-by-one loop
            pred += v1 * x1;
            pred += v2 * x2;
            pred += v3 * x3;
          }
#pragma unroll 1
          for (; q < qe; ++q) pred += lds_d(a_val + 8u * q) * lds_d(x + xs * (uint32_t)lds_i(a_col + 4u * q));
          if (!(centre || sigma)) {
            pred -= (double)yv;
            const double pl = half * pred * pred;
            if (va) pl_va += pl; else pl_tr += pl;
          }
          sts_d(a_out + 8u * (uint32_t)(p * Tm + t), pred);
        }
        BPROF(17);
        if (!(centre || sigma)) {
          const double nrm_tr = (Ptr > 0) ? __drcp_rn(pf ? (double)Ptr : (double)T * (double)Ptr) : 0.0;
          const double nrm_va = (Pva > 0) ? __drcp_rn(pf ? (double)Pva : (double)T * (double)Pva) : 0.0;
          pl_tr = wsum(pl_tr * nrm_tr);
          pl_va = wsum(pl_va * nrm_va);
          if (lane == 0) {
            s_red[warp][i] = pl_tr;
            s_red[warp][JAXENT_MAX_SLOTS + i] = pl_va;
          }
        }
        BPROF(18);
      }
      if (GENERAL && (centre || sigma)) {
        const float* tr_y = reinterpret_cast<const float*>(s_blob + bl.tr_y);
        const float* va_y = reinterpret_cast<const float*>(s_blob + bl.va_y);
        __syncthreads();
        if (centre) {
          // per-timepoint means of prediction and target (opt/losses.py:1814-1819): one warp per (set, t)
#pragma unroll 1
          for (int job = warp; job < 2 * T; job += nw) {
            const bool tr = job < T;
            const int t = tr ? job : job - T;
            const int P = tr ? Ptr : Pva;
            const double* rs = tr ? res_tr : res_va;
            const float* yy = tr ? tr_y : va_y;
            double s0 = 0.0;
#pragma unroll 1
            for (int p = lane; p < P; p += 32) s0 += rs[p * T + t] - (double)yy[p * T + t];
            s0 = wsum(s0);
            if (lane == 0) smean[job] = (P > 0) ? s0 / P : 0.0;  // mean(pred) - mean(y)
          }
          __syncthreads();
        }
        double pl2[2] = {0.0, 0.0};
#pragma unroll 1
        for (int set = 0; set < 2; ++set) {
          const int n = (set ? Pva : Ptr) * T;
          double* buf = set ? res_va : res_tr;
          const float* y = set ? va_y : tr_y;
#pragma unroll 1
          for (int j = tid; j < n; j += nt) {
            double r = buf[j] - (double)y[j];
            if (centre) r -= smean[set * T + (j - fdiv(j, invTm) * T)];
            buf[j] = r;
            if (!sigma) {
              pl2[set] += 0.5 * r * r;
              if (set == 0) gp_tr[j] = r;
            }
          }
        }
        if (sigma) {
          __syncthreads();
#pragma unroll 1
          for (int set = 0; set < 2; ++set) {
            const int P = set ? Pva : Ptr;
            const float* W = set ? pb.slots[i].val.W : pb.slots[i].train.W;
            const double* rs = set ? res_va : res_tr;
#pragma unroll 1
            for (int j = tid; j < P * T; j += nt) {
              const int p = fdiv(j, invTm), t = j - p * T;
              double a1 = 0.0, a2 = 0.0;
#pragma unroll 1
              for (int p2 = 0; p2 < P; ++p2) {
                const double rr = rs[p2 * T + t];
                a1 += (double)W[(size_t)p * P + p2] * rr;
                a2 += (double)W[(size_t)p2 * P + p] * rr;
              }
              pl2[set] += 0.5 * rs[j] * a1;
              if (set == 0) gp_tr[j] = 0.5 * (a1 + a2);
            }
          }
        }
#pragma unroll 1
        for (int set = 0; set < 2; ++set) {
          const int P = set ? Pva : Ptr;
          const double nrm = (P > 0) ? __drcp_rn((double)T * (double)P) : 0.0;
          const double t = wsum((set ? pl2[1] : pl2[0]) * nrm);
          if (lane == 0) s_red[warp][set * JAXENT_MAX_SLOTS + i] = t;
        }
        if (centre) {
          __syncthreads();
#pragma unroll 1
          for (int t = warp; t < T; t += nw) {
            double s0 = 0.0;
#pragma unroll 1
            for (int p = lane; p < Ptr; p += 32) s0 += gp_tr[p * T + t];
            s0 = wsum(s0);
            if (lane == 0) smean[2 * T + t] = s0 / Ptr;
          }
        }
      }
      __syncthreads();  // (3) dL/dpred visible
      BPROF(5);

      // (b) back-projection to dL/d lnPF, one (residue, time point) item per thread, shaped like (a)
      {
        const uint32_t blob = sm_addr(s_blob);
        const uint32_t a_tptr = blob + 4u * (uint32_t)bl.tr_tptr, a_trow = blob + 4u * (uint32_t)bl.tr_trow, a_tval = blob + 4u * (uint32_t)bl.tr_tval;
        const uint32_t a_gp = sm_addr(gp_tr);
        const double g = scale * (pf ? 2.0 : 1.0) * __drcp_rn(pf ? (double)Ptr : (double)T * (double)Ptr);
        const int items = R * Tm;
        BPROF(19);
#pragma unroll 1
        for (int it = tid; it < items; it += nt) {
          const int r = fdiv(it, invTm), t = it - r * Tm;  // time fastest: the 8 lanes of a residue read 8 consecutive doubles per entry
          const uint32_t gpt = a_gp + 8u * (uint32_t)t;
          const uint32_t tm8 = 8u * (uint32_t)Tm;
          const int qb = lds_i(a_tptr + 4u * r), qe = lds_i(a_tptr + 4u * r + 4u);
          // (loads issued before the loop: they only depend on the item)
          double xk = 0.0, uu = 0.0;
          if (!pf && !pfm) {
            xk = s_k[r] * s_tp[t] * sxe[r];
            uu = sU[it];
          }
          double back = 0.0, colsum = 0.0;
          int q = qb;
#pragma unroll 1
          for (; q + 4 <= qe; q += 4) {
            const uint32_t r0 = lds_i(a_trow + 4u * q), r1 = lds_i(a_trow + 4u * q + 4u), r2 = lds_i(a_trow + 4u * q + 8u), r3 = lds_i(a_trow + 4u * q + 12u);
            const double v0 = lds_d(a_tval + 8u * q), v1 = lds_d(a_tval + 8u * q + 8u), v2 = lds_d(a_tval + 8u * q + 16u), v3 = lds_d(a_tval + 8u * q + 24u);
            const double g0 = lds_d(gpt + tm8 * r0), g1 = lds_d(gpt + tm8 * r1), g2 = lds_d(gpt + tm8 * r2), g3 = lds_d(gpt + tm8 * r3);
            back += v0 * g0;
            back += v1 * g1;
            back += v2 * g2;
            back += v3 * g3;
            if (GENERAL) colsum += v0 + v1 + v2 + v3;
          }
#pragma unroll 1
          for (; q < qe; ++q) {
            const double v = lds_d(a_tval + 8u * q);
            back += v * lds_d(gpt + tm8 * (uint32_t)lds_i(a_trow + 4u * q));
            if (GENERAL) colsum += v;
          }
          if (centre) back -= colsum * smean[2 * T + t];  // centred gradient: minus mean_p(gp) through the column
          if (!pf && !pfm) back *= -xk * (1.0 - uu);      // du/dlnPF = -x exp(-x)
          sc_rt[(size_t)t * R + r] = g * back;
        }
        BPROF(20);
      }
      __syncthreads();  // (4) items visible
      BPROF(7);
      if (tid < R) {
        if (pfm) {
          // per-frame mode: the items ARE G[t,r] = dL/d<uptake>[t,r]; accumulate them over the data slots in global memory
          // (this thread owns residue tid); the data part of hbar = sum_{t,r} G <u> is formed after the slot loop
#pragma unroll 1
          for (int t = 0; t < Tm; ++t) {
            const double gv = sc_rt[(size_t)t * R + tid];
            float* gdst = io.pf_G + (size_t)t * R + tid;
            *gdst = (i == first_data) ? (float)gv : *gdst + (float)gv;
          }
        } else {
#pragma unroll 1
          for (int t = 0; t < Tm; ++t) my_c += sc_rt[(size_t)t * R + tid];
        }
      }
    }

    // ---- row coefficients for the next pass + one multi-value block reduction
    {
      double v12 = 0.0, v13 = 0.0, v14 = 0.0;
      if (tid < R) {
        // hbar (data part) = sum_j w_j H_j = sum_r k_r,  k_r = cc_r <Nc>_r + ch_r <Nh>_r  (per-frame mode: sum_t G[t,r] <u>[t,r]),
        // with the row coefficients AS THE SWEEP WILL USE THEM (rounded to fp32 / bf16 pieces): with exact c_r the identity
        // would be off by 2^-24 |hbar|, which the near-cancellation h_f - hbar of a dominant frame amplifies.
        // Centred column sums (io.ck != nullptr): the pass subtracts fl(k_r) inside every row's product, i.e. it forms
        // H_f - sum_r fl(k_r) = sum_r cc_r (Nc[r,f] - <Nc>_r) + ch_r (Nh[r,f] - <Nh>_r) + rounding, whose terms are small exactly
        // where the cancellation is worst (a frame that carries most of the weight looks like the average).  Only the
        // residual sum_r (k_r - fl(k_r)) is left for the control block.  An uncentred fp32 H_f cannot even REPRESENT the
        // column sum well enough there (measured: 2.5e-5 of the gradient scale at w_max = 0.997, |H| = 9e3).
        double kd;
        if (pfm) {  // G as the pass reads it (fp32, accumulated over the data slots by this thread)
          double gsum = 0.0, gu = 0.0;
#pragma unroll 1
          for (int t = 0; t < Tm; ++t) {
            const double gv = (double)io.pf_G[(size_t)t * R + tid];
            gsum += gv;
            gu += gv * sU[(size_t)tid * Tm + t];
          }
          kd = gsum - gu;  // the pass forms fl(kd) - sum_t G exp(-k tau / PF_f) = sum_t G (u_f - <u>) per row
          v12 = (io.ck != nullptr) ? -(kd - (double)(float)kd) : gu;
        } else {
          // (the batched GEMM sweeps centred features N - off_r: its column sums, and hence hbar, are about <N>_r - off_r)
          const double oa = io.row_off ? (double)io.row_off[tid] : 0.0, ob = io.row_off ? (double)io.row_off[io.row_off_stride + tid] : 0.0;
          kd = coeff_as_used(my_c * bc, io.c_pieces) * (my_a - oa) + coeff_as_used(my_c * bh, io.c_pieces) * (my_b - ob);
          v12 = (io.ck != nullptr) ? kd - (double)(float)kd : kd;
        }
        if (io.ck != nullptr) io.ck[tid] = (float)kd;
        v13 = my_c * my_a;   // dL/dbc
        v14 = my_c * my_b;   // dL/dbh
      }
      if (warp * 32 < R) {
        v12 = wsum(v12);
        v13 = wsum(v13);
        v14 = wsum(v14);
        if (lane == 0) {
          s_red[warp][12] = v12;
          s_red[warp][13] = v13;
          s_red[warp][14] = v14;
        }
      }
    }
    __syncthreads();  // (5)
    BPROF(8);
    if (warp < 15) {  // value `warp` over the per-warp partials: one load per lane + a fixed-shape shuffle tree
      const double t = wsum(lane < nw ? s_red[lane][warp] : 0.0);
      if (lane == 0) s_fin[warp] = t;
    }
    __syncthreads();  // (6)

    gbc_reg_out = gbc_reg;
    gbh_reg_out = gbh_reg;
  }
}

}  // namespace jx
