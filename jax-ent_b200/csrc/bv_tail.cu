// The "R x T tail" of the BV MaxEnt optimise step and the per-frame MaxEnt terms, sm_100a.
//
// Runs between two streaming passes.  Every CTA resolves the plateau-LR branch from the reduced
// partial sums; CTA 0 then evaluates
//   ln PF = bc*<Nc> + bh*<Nh>                   BV_ForwardPass        models/HDX/forward.py:18-37
//   uptake = 1 - exp(-k t / PF)                 BV_uptake_ForwardPass models/HDX/forward.py:45-76
//   pred = S u  (CSR)                           apply_sparse_mapping  data/splitting/sparse_map.py:220-235
//   MSE / mean-centred MSE / sigma MSE / PF-L2  opt/losses.py:1541-1657,1790-1841,17-50
//   L1/L2 parameter penalties                   opt/losses.py:476-513
// and their closed-form backward down to c_r = dL/d lnPF_r, the BV-parameter Adam update
// (opt/optimiser.py:130-134,416-424) and the control block for the next pass, while all CTAs
// compute the per-frame MaxEnt terms of maxent_convexKL_loss (opt/losses.py:304-322):
// softmax weights, dKL/dw and the three KL sums.  Sums are carried in fp64.
#include <cmath>

#include "jx_internal.cuh"

namespace jx {

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block-wide sum; result valid in all threads.  scratch: >= 33 doubles.
__device__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = (lane < nw) ? scratch[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

struct TailShared {
  int d;            // branch chosen for the update that produced the current parameters
  int pending;
  double S;         // softmax denominator of the current z relative to the old shift
  float lse_old;
  float bc, bh;
  int off;          // offset of the chosen accumulators inside red
};

// One peptide map (train or val) of one data slot.  Returns the unscaled loss; when `grad`,
// accumulates dL_total/d lnPF_r into sc[].
__device__ double map_loss(const JaxentPeptideMap& mp, int kind, int R, int T, bool uptake, const double* sE,
                           const double* slp, const double* sxe, const float* k_ints, const float* tps, double* sres,
                           double* sgp, double* smean, double* scratch, bool grad, double scale, double* sc) {
  const int P = mp.n_pep;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (P <= 0) return 0.0;
  if (kind == JAXENT_LOSS_PF_L2) {
    double part = 0.0;
    for (int p = tid; p < P; p += nt) {
      double pred = 0.0;
      for (int j = mp.row_ptr[p]; j < mp.row_ptr[p + 1]; ++j) pred += (double)mp.val[j] * slp[mp.col[j]];
      const double r = pred - (double)mp.y[p];
      sres[p] = r;
      part += r * r;
    }
    const double loss = block_sum(part, scratch) / (double)P;
    if (grad) {
      const double g = scale * 2.0 / (double)P;
      for (int r = tid; r < R; r += nt) {
        double acc = 0.0;
        for (int j = mp.t_ptr[r]; j < mp.t_ptr[r + 1]; ++j) acc += (double)mp.t_val[j] * sres[mp.t_row[j]];
        sc[r] += g * acc;
      }
    }
    __syncthreads();
    return loss;
  }
  // ---- uptake kinds ----
  const bool centre = kind == JAXENT_LOSS_UPTAKE_MC_EYE_MSE;
  const bool sigma = kind == JAXENT_LOSS_UPTAKE_SIGMA_MSE && mp.W != nullptr;
  const int n = P * T;
  for (int i = tid; i < n; i += nt) {
    const int p = i / T, t = i - p * T;
    double pred = 0.0;
    for (int j = mp.row_ptr[p]; j < mp.row_ptr[p + 1]; ++j) pred += (double)mp.val[j] * (1.0 - sE[t * R + mp.col[j]]);
    sres[i] = pred;  // pred[p,t]
  }
  __syncthreads();
  if (centre) {
    // per-timepoint means of prediction and target (opt/losses.py:1814-1819)
    for (int t = tid; t < T; t += nt) {
      double mp_ = 0.0, my = 0.0;
      for (int p = 0; p < P; ++p) {
        mp_ += sres[p * T + t];
        my += (double)mp.y[p * T + t];
      }
      smean[t] = mp_ / P;
      smean[T + t] = my / P;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) {
    const int t = i % T;
    double r = sres[i] - (double)mp.y[i];
    if (centre) r -= smean[t] - smean[T + t];
    sres[i] = r;  // residual[p,t]
  }
  __syncthreads();
  double part = 0.0;
  for (int i = tid; i < n; i += nt) {
    const int p = i / T, t = i - p * T;
    double wr, sym;
    if (sigma) {
      double a1 = 0.0, a2 = 0.0;
      for (int p2 = 0; p2 < P; ++p2) {
        const double rr = sres[p2 * T + t];
        a1 += (double)mp.W[(size_t)p * P + p2] * rr;
        a2 += (double)mp.W[(size_t)p2 * P + p] * rr;
      }
      wr = a1;
      sym = 0.5 * (a1 + a2);
    } else {
      wr = sres[i];
      sym = sres[i];
    }
    part += 0.5 * sres[i] * wr;
    sgp[i] = sym;
  }
  const double loss = block_sum(part, scratch) / ((double)T * (double)P);
  if (grad) {
    const double g = scale / ((double)T * (double)P);
    if (centre) {
      for (int t = tid; t < T; t += nt) {
        double mg = 0.0;
        for (int p = 0; p < P; ++p) mg += sgp[p * T + t];
        smean[t] = mg / P;
      }
      __syncthreads();
      for (int i = tid; i < n; i += nt) sgp[i] -= smean[i % T];
      __syncthreads();
    }
    for (int r = tid; r < R; r += nt) {
      const double kr = (double)k_ints[r], xe = sxe[r];
      double acc = 0.0;
      for (int t = 0; t < T; ++t) {
        double back = 0.0;
        for (int j = mp.t_ptr[r]; j < mp.t_ptr[r + 1]; ++j) back += (double)mp.t_val[j] * sgp[mp.t_row[j] * T + t];
        const double x = kr * (double)tps[t] * xe;
        acc += back * (-x * sE[t * R + r]);  // du/dlnPF = -x exp(-x)
      }
      sc[r] += g * acc;
    }
  }
  __syncthreads();
  return loss;
}

__global__ void __launch_bounds__(TAIL_THREADS) bv_tail_kernel(const TailArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ TailShared sh;
  __shared__ double scratch[40];
  __shared__ int s_last;

  const JaxentBvProblem& pb = a.prob;
  const int R = pb.R, T = pb.uptake ? pb.T : 0;
  const int tid = threadIdx.x, nt = blockDim.x;
  const unsigned ep = *a.epoch;
  const Ctl* __restrict__ old = a.ctl + ((ep - 1u) & 1u);
  Ctl* __restrict__ neu = a.ctl + (ep & 1u);
  const int cur_set = ep & 1u;

  // ------------------------------------------------------------------ A. branch decision (all CTAs)
  if (tid == 0) {
    int d = 0;
    if (old->pending) {
      const double dot = a.red[4 * R + 2] + (double)(old->have_prev ? old->gb_prev[0] : old->gb[0]) * (double)old->gb[0] +
                         (double)(old->have_prev ? old->gb_prev[1] : old->gb[1]) * (double)old->gb[1];
      d = ((float)dot < 0.f && old->step > 1) ? 1 : 0;  // opt/optimiser.py:403
      scratch[34] = dot;
    }
    sh.d = d;
    sh.pending = old->pending;
    sh.S = a.red[4 * R + d];
    sh.lse_old = old->lse;
    sh.off = d * 2 * R;
  }
  __syncthreads();
  const int d = sh.d;
  const double S = sh.S;

  // ------------------------------------------------------------------ B. the R x T tail (CTA 0)
  if (blockIdx.x == 0) {
    double* sa = reinterpret_cast<double*>(smem_raw);
    double* sb = sa + R;
    double* slp = sb + R;
    double* sc = slp + R;
    double* sxe = sc + R;
    double* sE = sxe + R;                 // [T][R] exp(-x)
    double* sres = sE + (size_t)T * R;    // [p_max * max(T,1)]
    double* sgp = sres + (size_t)a.p_max * (T > 0 ? T : 1);
    double* smean = sgp + (size_t)a.p_max * (T > 0 ? T : 1);  // [2*max(T,1)]

    // B1. BV-parameter Adam for the update that led to the current parameters
    if (tid == 0) {
      float bc = old->bc, bh = old->bh;
      float mb0 = old->mb[0], mb1 = old->mb[1], vb0 = old->vb[0], vb1 = old->vb[1];
      int cm = old->count_model, cf = old->count_frame, step = old->step, mask_idx = old->mask_idx;
      float lr = old->lr, mlr = old->model_lr;
      if (old->pending) {
        lr = d ? old->lrB : old->lrA;
        mlr = d ? old->mlrB : old->mlrA;
        const float b1 = (float)a.cfg.b1, b2 = (float)a.cfg.b2, eps = (float)a.cfg.eps, clip = a.cfg.clip_value;
        const float om_b1 = (float)(1.0 - a.cfg.b1), om_b2 = (float)(1.0 - a.cfg.b2);
        cm += 1;
        cf += 1;
        const float c1 = 1.0f - powf(b1, (float)cm), c2 = 1.0f - powf(b2, (float)cm);
        float par[2] = {bc, bh};
        float mm[2] = {mb0, mb1}, vv[2] = {vb0, vb1};
        for (int i = 0; i < 2; ++i) {
          float sg = old->gb[i] * mlr;
          if (clip > 0.f) sg = fminf(fmaxf(sg, -clip), clip);
          mm[i] = om_b1 * sg + b1 * mm[i];
          vv[i] = om_b2 * (sg * sg) + b2 * vv[i];
          float u = -((mm[i] / c1) / (sqrtf(vv[i] / c2) + eps));
          if (par[i] + u < 0.f) u = -par[i];  // optax.keep_params_nonnegative
          par[i] = par[i] + u;
        }
        bc = par[0]; bh = par[1];
        mb0 = mm[0]; mb1 = mm[1]; vb0 = vv[0]; vb1 = vv[1];
        mask_idx = (step >= a.cfg.initial_steps) ? 1 : mask_idx;
        step += 1;
      }
      sh.bc = bc;
      sh.bh = bh;
      neu->step = step;
      neu->mask_idx = mask_idx;
      neu->branch = d;
      neu->pending = 1;
      neu->have_prev = old->pending ? 1 : old->have_prev;
      neu->count_frame = cf;
      neu->count_model = cm;
      neu->kl_slot = old->kl_slot;
      neu->lr = lr;
      neu->model_lr = mlr;
      neu->bc = bc;
      neu->bh = bh;
      neu->mb[0] = mb0; neu->mb[1] = mb1; neu->vb[0] = vb0; neu->vb[1] = vb1;
      neu->gb_prev[0] = old->gb[0];
      neu->gb_prev[1] = old->gb[1];
      neu->last_dot = old->pending ? (float)scratch[34] : 0.f;
      neu->last_lr = lr;
      neu->last_branch = d;
      neu->kl_scale = old->kl_scale;
      for (int i = 0; i < JAXENT_MAX_SLOTS; ++i) { neu->fw[i] = old->fw[i]; neu->fs[i] = old->fs[i]; }
      // speculative rates of the NEXT update (state.step == step): opt/optimiser.py:365-381,403-413
      const bool switched = step >= a.cfg.initial_steps;
      const int mi = switched ? 1 : mask_idx;
      const float base_lr = switched ? a.cfg.learning_rate : lr;
      const float base_mlr = switched ? a.cfg.learning_rate * a.cfg.model_parameters_lr_scale : mlr;
      neu->lrA = base_lr;
      neu->lrB = base_lr / a.cfg.plateau_denominator;
      neu->mlrA = base_mlr;
      neu->mlrB = base_mlr / a.cfg.plateau_denominator;
      neu->mask_frame = (mi == 0) ? 1.f : (a.cfg.optimise_frame_weights ? 1.f : 0.f);
      neu->mask_model = (mi == 0) ? 0.f : (a.cfg.optimise_model_parameters ? 1.f : 0.f);
      neu->bc1 = 1.0f - powf((float)a.cfg.b1, (float)(cf + 1));
      neu->bc2 = 1.0f - powf((float)a.cfg.b2, (float)(cf + 1));
      neu->lse = (float)((double)old->lse + log(S));
      neu->rec_idx = step % JAXENT_RECORD_RING;
    }
    __syncthreads();
    const double bc = (double)sh.bc, bh = (double)sh.bh;

    // B2. frame averages, ln PF, uptake
    for (int r = tid; r < R; r += nt) {
      const double av = a.red[sh.off + r] / S, bv = a.red[sh.off + R + r] / S;
      const double lp = bc * av + bh * bv;
      sa[r] = av;
      sb[r] = bv;
      slp[r] = lp;
      sc[r] = 0.0;
      a.avg[r] = (float)av;
      a.avg[R + r] = (float)bv;
      a.logpf[r] = (float)lp;
      if (T > 0) {
        const double xe = exp(-lp), kr = (double)pb.k_ints[r];
        sxe[r] = xe;
        for (int t = 0; t < T; ++t) {
          const double E = exp(-(kr * (double)pb.timepoints[t] * xe));
          sE[t * R + r] = E;
          a.uptake[t * R + r] = (float)(1.0 - E);
        }
      }
    }
    __syncthreads();

    // B3. loss slots
    JaxentLossRecord* rec = a.records + (neu->step % JAXENT_RECORD_RING);
    double gbc_reg = 0.0, gbh_reg = 0.0;
    for (int i = 0; i < pb.n_slots; ++i) {
      const JaxentLossSlot& sl = pb.slots[i];
      const double scale = (double)old->fw[i] * (double)old->fs[i];
      double ltr = 0.0, lva = 0.0;
      if (sl.kind <= JAXENT_LOSS_PF_L2) {
        ltr = map_loss(sl.train, sl.kind, R, T, pb.uptake, sE, slp, sxe, pb.k_ints, pb.timepoints, sres, sgp, smean,
                       scratch, true, scale, sc);
        lva = map_loss(sl.val, sl.kind, R, T, pb.uptake, sE, slp, sxe, pb.k_ints, pb.timepoints, sres, sgp, smean,
                       scratch, false, scale, sc);
      } else if (sl.kind == JAXENT_LOSS_PARAMS_L1) {
        const double dc = bc - (double)sl.prior_bc, dh = bh - (double)sl.prior_bh;
        ltr = lva = fabs(dc) + fabs(dh);
        gbc_reg += scale * ((dc > 0) - (dc < 0));
        gbh_reg += scale * ((dh > 0) - (dh < 0));
      } else if (sl.kind == JAXENT_LOSS_PARAMS_L2) {
        const double dc = bc - (double)sl.prior_bc, dh = bh - (double)sl.prior_bh;
        ltr = lva = dc * dc + dh * dh;
        gbc_reg += scale * 2.0 * dc;
        gbh_reg += scale * 2.0 * dh;
      }
      if (tid == 0) {
        rec->train[i] = (float)ltr;
        rec->val[i] = (float)lva;
      }
    }
    __syncthreads();

    // B4. row coefficients for the next pass, hbar (data part), BV-parameter gradients
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;
    for (int r = tid; r < R; r += nt) {
      const double c = sc[r];
      a.cc[r] = (float)(c * bc);
      a.ch[r] = (float)(c * bh);
      p0 += c * slp[r];
      p1 += c * sa[r];
      p2 += c * sb[r];
    }
    const double hbar = block_sum(p0, scratch);
    const double gbc = block_sum(p1, scratch) + gbc_reg;
    const double gbh = block_sum(p2, scratch) + gbh_reg;
    if (tid == 0) {
      neu->hbar_data = hbar;
      neu->gb[0] = (float)gbc * neu->mask_model;
      neu->gb[1] = (float)gbh * neu->mask_model;
      rec->step = neu->step;
      rec->branch = d;
      rec->lr = neu->last_lr;
      rec->grad_dot = neu->last_dot;
      rec->bc = (float)bc;
      rec->bh = (float)bh;
      rec->grad_bc = neu->gb[0];
      rec->grad_bh = neu->gb[1];
      rec->complete = 0;
      for (int i = pb.n_slots; i < JAXENT_MAX_SLOTS; ++i) { rec->train[i] = 0.f; rec->val[i] = 0.f; }
      if (old->kl_slot < 0) finish_record(neu, a.klsum, a.records, pb.n_slots);
    }
  }

  // ------------------------------------------------------------------ C. per-frame MaxEnt terms (all CTAs)
  {
    const float* __restrict__ z = a.st.z[cur_set][d];
    const long long G = gridDim.x, cta = blockIdx.x;
    const long long f0 = pb.F * cta / G, f1 = pb.F * (cta + 1) / G;
    const float lse_old = sh.lse_old;
    const float Sf = (float)S;
    const bool has_kl = old->kl_slot >= 0 && pb.prior != nullptr;
    const float eps = a.eps_kl;
    const float lam = old->kl_scale;
    const double pnorm = pb.prior_norm;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (long long f = f0 + tid; f < f1; f += nt) {
      const float w = expf(z[f] - lse_old) / Sf;  // jax.nn.softmax: exp(x - shift) / sum
      a.w[f] = w;
      float kap = 0.f;
      if (has_kl) {
        // q = (|w|+eps)/sum(|w|+eps); the denominator is 1 + F*eps = 1 + 1e-10 == 1 in fp32
        const float q = fabsf(w) + eps;
        const float p = (float)(((double)fabsf(pb.prior[f]) + (double)eps) / pnorm);
        // p/q and log(p/q) in fp64: differencing two fp32 logs leaves a rounding error that is
        // identical on every frame of a uniform prior and so does not average out over F
        const double ratio = (double)p / (double)q;
        const float G_ = (float)(1.0 - ratio);
        kap = (w > 0.f) ? lam * G_ : 0.f;  // d|w|/dw = sign(w); softmax weights are >= 0
        s0 += (double)w * (double)kap;
        if (p > 0.f) s1 += (double)p * log(ratio);
        s2 += (double)q - (double)p;
      }
      a.kappa[f] = kap;
      s3 += (double)w;
    }
    s0 = block_sum(s0, scratch);
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    s3 = block_sum(s3, scratch);
    if (tid == 0) {
      double* kp = a.klpart + (size_t)blockIdx.x * KL_N;
      kp[0] = s0; kp[1] = s1; kp[2] = s2; kp[3] = s3;
      __threadfence();
      const unsigned ticket = atomicAdd(a.counter, 1u);
      s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && tid < KL_N) {
      __threadfence();
      double t = 0.0;
      for (unsigned c = 0; c < gridDim.x; ++c) t += a.klpart[(size_t)c * KL_N + tid];  // fixed order
      a.klsum[tid] = t;
      if (tid == 0) *a.counter = 0u;
    }
  }
}

size_t tail_smem_bytes(int R, int T, int p_max) {
  const size_t t = T > 0 ? T : 1;
  return sizeof(double) * ((size_t)5 * R + (size_t)T * R + 2 * (size_t)p_max * t + 2 * t + 8);
}

int launch_tail(const TailArgs& a, int ctas, size_t smem, cudaStream_t s) {
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(bv_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("tail kernel needs %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return JAXENT_E_UNSUPPORTED; }
    configured = smem;
  }
  bv_tail_kernel<<<ctas, TAIL_THREADS, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bv_tail_kernel launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

}  // namespace jx
