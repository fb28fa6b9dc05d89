// The "R x T tail" of the BV MaxEnt optimise step and the per-frame MaxEnt terms, sm_100a.
//
// Runs between two streaming passes.  CTA 0 evaluates
//   ln PF = bc*<Nc> + bh*<Nh>                   BV_ForwardPass        models/HDX/forward.py:18-37
//   uptake = 1 - exp(-k t / PF)                 BV_uptake_ForwardPass models/HDX/forward.py:45-76
//   pred = S u  (CSR)                           apply_sparse_mapping  data/splitting/sparse_map.py:220-235
//   MSE / mean-centred MSE / sigma MSE / PF-L2  opt/losses.py:1541-1657,1790-1841,17-50
//   L1/L2 parameter penalties                   opt/losses.py:476-513
// their closed-form backward down to c_r = dL/d lnPF_r, the plateau decision, the BV-parameter Adam
// update (opt/optimiser.py:130-134,403-424) and the control block of the next pass; the other CTAs
// compute the per-frame terms of maxent_convexKL_loss (opt/losses.py:304-322): softmax weights,
// dKL/dw and the KL sums.  All sums are carried in fp64.
//
// This kernel sits on the critical path of every step, so it is organised around latency:
//  * each data slot's peptide maps, transposes and targets live in ONE contiguous device blob
//    (gathered once by build_blobs_kernel) whose image is the shared-memory layout, so staging is a
//    single vectorised copy with every load in flight at once, overlapped with the control-block,
//    k_int / time-point and partial-sum reads;
//  * every thread derives the step scalars (branch, BV-parameter Adam, next learning rates, bias
//    corrections, new log-sum-exp) redundantly from shared memory -- no leader-thread sections;
//  * the back-projection is parallel over (residue, time point) and the loss sums are reduced once.
#include <cmath>

#include "bv_tail_body.cuh"
#include "jx_internal.cuh"

namespace jx {

#ifdef JX_PROFILE
#define TPROF(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<long long*>(a.cl_scratch)[i] = gtime(); } while (0)
#define TPROF_F(i) do { if (blockIdx.x == 1 && threadIdx.x == 0) reinterpret_cast<long long*>(a.cl_scratch)[32 + i] = gtime(); } while (0)
#else
#define TPROF(i) do {} while (0)
#define TPROF_F(i) do {} while (0)
#endif

// step scalars, derived once by warp 0 and broadcast through shared memory
struct StepScalars {
  int d;  // plateau branch taken by the update that produced the current parameters
  int pending, step, mask_idx, cm, cf, have_prev;
  float lr, mlr, bc, bh, mb0, mb1, vb0, vb1;
  float lrA, lrB, mlrA, mlrB, mask_frame, mask_model, bc1, bc2, lse_new, lse_old;
  double S, dot;
  // on-device loop (ConvergenceTracker state after this step + what this launch has to do)
  int trk_stop, trk_reason, trk_stop_step, trk_have_prev, trk_have_ema, trk_steps_since, trk_idx, trk_nsnap;
  int snap_now;             // snapshot slot to fill in this launch, or -1
  int ema_update, ema_first; // update the EMA parameters in this launch / first EMA value
  int loop_step;            // loop index k of the step being resolved
  float trk_ema_bc, trk_ema_bh;
  double trk_prev_loss, trk_ema;
};

// The step scalars are derived in three parts:
//   derive_spec     : BEFORE the pass has finished (the tail is launched early, PDL): everything that only needs the previous
//                     control block -- the BV-parameter Adam update for BOTH plateau branches (lane b takes branch b), bias
//                     corrections, counters, masks.
//   derive_critical : after the pass: the plateau decision from the reduced grad-dot, a select between the two prepared
//                     updates, S.  A handful of instructions between the arrival of the sums and the block-wide barrier (1).
//   derive_rest     : the learning rates / bias corrections / log-sum-exp of the COMING update and the convergence tracker.
//                     Never on the R x T tail's path: a frame CTA (block 1) owns these fields of the new control block.
// SpecScalars / derive_spec live in jx_internal.cuh: the per-frame pass derives the same two BV-parameter updates (bv_pass.cu)

// one thread
__device__ __forceinline__ void derive_critical(StepScalars* out, const SpecScalars* __restrict__ sp, const Ctl* __restrict__ old, double SA,
                                                double SB, double gdot, const JaxentOptConfig& cfg) {
  StepScalars s;
  s.pending = old->pending;
  s.step = old->step;
  s.mask_idx = old->mask_idx;
  s.cm = old->count_model;
  s.cf = old->count_frame;
  s.have_prev = old->have_prev;
  s.lr = old->lr;
  s.mlr = old->model_lr;
  s.bc = old->bc;
  s.bh = old->bh;
  s.mb0 = old->mb[0];
  s.mb1 = old->mb[1];
  s.vb0 = old->vb[0];
  s.vb1 = old->vb[1];
  s.lse_old = old->lse;
  s.d = 0;
  s.dot = 0.0;
  if (s.pending) {
    s.dot = gdot + sp->dot_model;
    const int d = ((float)s.dot < 0.f && s.step > 1) ? 1 : 0;  // opt/optimiser.py:403
    s.d = d;
    s.lr = sp->lr[d];
    s.mlr = sp->mlr[d];
    s.bc = sp->bc[d];
    s.bh = sp->bh[d];
    s.mb0 = sp->mb0[d];
    s.mb1 = sp->mb1[d];
    s.vb0 = sp->vb0[d];
    s.vb1 = sp->vb1[d];
    s.cm += 1;
    s.cf += 1;
    s.have_prev = 1;
    s.mask_idx = (s.step >= cfg.initial_steps) ? 1 : s.mask_idx;
    s.step += 1;
  }
  s.S = s.d ? SB : SA;
  s.loop_step = old->step;
  // masks of the COMING update (state.step == step): opt/optimiser.py:365-381
  {
    const int mi = (s.step >= cfg.initial_steps) ? 1 : s.mask_idx;
    s.mask_frame = (mi == 0) ? 1.f : (cfg.optimise_frame_weights ? 1.f : 0.f);
    s.mask_model = (mi == 0) ? 0.f : (cfg.optimise_model_parameters ? 1.f : 0.f);
  }
  // the fields of derive_rest start from their carried values
  s.trk_stop = old->trk_stop;
  s.snap_now = -1;
  s.ema_update = 0;
  s.ema_first = 0;
  *out = s;
}

// second half: reads the critical part from *sc (shared memory), completes it in place.  One whole warp.
__device__ __noinline__ void derive_rest(StepScalars* sc, const Ctl* __restrict__ old, const JaxentOptConfig cfg, const TrackConfig trk,
                                         const JaxentLossRecord* __restrict__ records) {
  const int lane = threadIdx.x & 31;
  StepScalars s = *sc;
  // lanes 0,1: b1^(cf+1), b2^(cf+1) (bias corrections of the coming frame Adam update);  lane 2: log S
  float pw = 0.f;
  double lg = 0.0;
  if (lane < 2) pw = fpow((lane & 1) ? (float)cfg.b2 : (float)cfg.b1, (float)(s.cf + 1));
  if (lane == 2) lg = dlog(s.S);
  const float pw2 = __shfl_sync(0xffffffffu, pw, 0), pw3 = __shfl_sync(0xffffffffu, pw, 1);
  const double lgS = __shfl_sync(0xffffffffu, lg, 2);
  if (lane != 0) return;
  // speculative rates of the NEXT update (state.step == step): opt/optimiser.py:365-381,403-413
  const bool switched = s.step >= cfg.initial_steps;
  const float base_lr = switched ? cfg.learning_rate : s.lr;
  const float base_mlr = switched ? cfg.learning_rate * cfg.model_parameters_lr_scale : s.mlr;
  s.lrA = base_lr;
  s.lrB = base_lr / cfg.plateau_denominator;
  s.mlrA = base_mlr;
  s.mlrB = base_mlr / cfg.plateau_denominator;
  s.bc1 = 1.0f - pw2;
  s.bc2 = 1.0f - pw3;
  s.lse_new = (float)((double)s.lse_old + lgS);

  // ---- ConvergenceTracker for loop step k = old->step (reference opt/run.py:285-348, opt/track.py:150-197)
  s.trk_stop = old->trk_stop;
  s.trk_reason = old->trk_reason;
  s.trk_stop_step = old->trk_stop_step;
  s.trk_have_prev = old->trk_have_prev;
  s.trk_have_ema = old->trk_have_ema;
  s.trk_steps_since = old->trk_steps_since;
  s.trk_idx = old->trk_idx;
  s.trk_nsnap = old->trk_nsnap;
  s.trk_prev_loss = old->trk_prev_loss;
  s.trk_ema = old->trk_ema;
  s.trk_ema_bc = old->trk_ema_bc;
  s.trk_ema_bh = old->trk_ema_bh;
  s.snap_now = -1;
  s.ema_update = 0;
  s.ema_first = 0;
  if (trk.on && s.pending && !s.trk_stop) {
    const int k = old->step;
    const double loss = (double)records[k % JAXENT_RECORD_RING].total_train;
    const double alpha = (double)trk.ema_alpha;
    if (s.trk_have_prev) {  // tracker.update
      const double raw = fabs(s.trk_prev_loss - loss);
      if (!s.trk_have_ema) {
        s.trk_ema = raw;
        s.trk_have_ema = 1;
        s.ema_first = 1;
      } else {
        s.trk_ema = alpha * raw + (1.0 - alpha) * s.trk_ema;
      }
      s.ema_update = 1;
      s.trk_ema_bc = s.ema_first ? s.bc : (float)(alpha * (double)s.bc + (1.0 - alpha) * (double)s.trk_ema_bc);
      s.trk_ema_bh = s.ema_first ? s.bh : (float)(alpha * (double)s.bh + (1.0 - alpha) * (double)s.trk_ema_bh);
    }
    s.trk_steps_since += 1;
    s.trk_prev_loss = loss;
    s.trk_have_prev = 1;
    const double gd = (k > 0) ? s.dot : 0.0;  // check_gradient_oscillation: zeros before the first step
    if (gd < 0.0) s.trk_steps_since = 0;
    if (loss < (double)trk.tolerance || isnan(loss) || isinf(loss)) {
      s.trk_stop = 1;
      s.trk_reason = 2;
      s.trk_stop_step = k;
    } else {
      const double rel = (s.trk_have_ema && loss > 0.0) ? s.trk_ema / loss : 0.0;
      const int ti = s.trk_idx < trk.n_thr ? s.trk_idx : trk.n_thr - 1;
      const bool met = s.trk_steps_since >= trk.min_steps && s.trk_have_ema && rel < (double)trk.thr[ti] && k > trk.initial_steps;
      if (k == 0 || met) {
        s.snap_now = s.trk_nsnap;
        s.trk_nsnap += 1;
        if (k > 0) {
          s.trk_idx += 1;
          s.trk_steps_since = 0;
          if (s.trk_idx >= trk.n_thr) {
            s.trk_stop = 1;
            s.trk_reason = 1;
            s.trk_stop_step = k;
          }
        }
      }
    }
  }
  *sc = s;
}

// ------------------------------------------------------------------------------------------ blobs
__global__ void build_blobs_kernel(JaxentBvProblem pb, BlobSet bs) {
  const int R = pb.R, Tm = (pb.uptake && pb.T > 0) ? pb.T : 1;
  const int tid = threadIdx.x, nt = blockDim.x;  // single block
  auto cp = [&](int* dst, const void* src, int n) {
    const int* s = static_cast<const int*>(src);
    for (int i = tid; i < n; i += nt) dst[i] = s[i];
  };
  auto cpd = [&](int* dst, const float* src, int n) {  // map values: float -> double (see blob_layout)
    double* d = reinterpret_cast<double*>(dst);
    for (int i = tid; i < n; i += nt) d[i] = (double)src[i];
  };
  for (int i = 0; i < pb.n_slots; ++i) {
    const JaxentLossSlot& sl = pb.slots[i];
    if (sl.kind > JAXENT_LOSS_PF_L2 || bs.blob[i] == nullptr) continue;
    const BlobLayout b = blob_layout(R, Tm, sl.train.n_pep, sl.train.nnz, sl.val.n_pep, sl.val.nnz);
    int* w = bs.blob[i];
    for (int k = tid; k < b.words; k += nt) w[k] = 0;
  }
  __syncthreads();  // zero fill before the copies
  for (int i = 0; i < pb.n_slots; ++i) {
    const JaxentLossSlot& sl = pb.slots[i];
    if (sl.kind > JAXENT_LOSS_PF_L2 || bs.blob[i] == nullptr) continue;
    const BlobLayout b = blob_layout(R, Tm, sl.train.n_pep, sl.train.nnz, sl.val.n_pep, sl.val.nnz);
    int* w = bs.blob[i];
    cp(w + b.tr_ptr, sl.train.row_ptr, sl.train.n_pep + 1);
    cp(w + b.tr_col, sl.train.col, sl.train.nnz);
    cp(w + b.tr_tptr, sl.train.t_ptr, R + 1);
    cp(w + b.tr_trow, sl.train.t_row, sl.train.nnz);
    cpd(w + b.tr_val, sl.train.val, sl.train.nnz);
    cpd(w + b.tr_tval, sl.train.t_val, sl.train.nnz);
    cp(w + b.tr_y, sl.train.y, sl.train.n_pep * Tm);
    if (sl.val.n_pep > 0) {
      cp(w + b.va_ptr, sl.val.row_ptr, sl.val.n_pep + 1);
      cp(w + b.va_col, sl.val.col, sl.val.nnz);
      cpd(w + b.va_val, sl.val.val, sl.val.nnz);
      cp(w + b.va_y, sl.val.y, sl.val.n_pep * Tm);
    }
  }
}

int launch_build_blobs(const JaxentBvProblem& pb, const BlobSet& bs, cudaStream_t s) {
  build_blobs_kernel<<<1, 1024, 0, s>>>(pb, bs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("build_blobs launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

// ------------------------------------------------------------------------------------------ shared pieces
// The new control block has two writers with disjoint fields.  The R x T tail CTA (block 0) writes what it produces:
// parameters, gradients, counters, masks, the loss record.  One warp; lane 0 the scalars, lanes 0..5 the slots.
__device__ __forceinline__ void write_ctl_critical(Ctl* __restrict__ neu, JaxentLossRecord* records, const StepScalars& sc,
                                                   const Ctl& s_old, const double* s_fin, double gbc_reg, double gbh_reg,
                                                   int n_slots, int lane, unsigned ep) {
  const float gb0 = (float)(s_fin[13] + gbc_reg) * sc.mask_model, gb1 = (float)(s_fin[14] + gbh_reg) * sc.mask_model;
  if (lane == 0) {
    neu->step = sc.step;
    neu->mask_idx = sc.mask_idx;
    neu->branch = sc.d;
    neu->pending = 1;
    neu->have_prev = sc.have_prev;
    neu->count_frame = sc.cf;
    neu->count_model = sc.cm;
    neu->kl_slot = s_old.kl_slot;
    neu->lr = sc.lr;
    neu->model_lr = sc.mlr;
    neu->bc = sc.bc;
    neu->bh = sc.bh;
    neu->mb[0] = sc.mb0; neu->mb[1] = sc.mb1; neu->vb[0] = sc.vb0; neu->vb[1] = sc.vb1;
    neu->gb_prev[0] = s_old.gb[0];
    neu->gb_prev[1] = s_old.gb[1];
    neu->gb[0] = gb0;
    neu->gb[1] = gb1;
    neu->mask_frame = sc.mask_frame;
    neu->mask_model = sc.mask_model;
    neu->kl_scale = s_old.kl_scale;
    neu->last_dot = (float)sc.dot;
    neu->last_lr = sc.lr;
    neu->last_branch = sc.d;
    neu->rec_idx = sc.step % JAXENT_RECORD_RING;
    neu->hbar_data = s_fin[12];
    neu->epoch_id = (int)ep;
  }
  if (lane < JAXENT_MAX_SLOTS) {
    neu->fw[lane] = s_old.fw[lane];
    neu->fs[lane] = s_old.fs[lane];
  }
  JaxentLossRecord* rec = records + (sc.step % JAXENT_RECORD_RING);
  if (lane == 1) {
    rec->step = sc.step;
    rec->branch = sc.d;
    rec->lr = sc.lr;
    rec->grad_dot = (float)sc.dot;
    rec->bc = sc.bc;
    rec->bh = sc.bh;
    rec->grad_bc = gb0;
    rec->grad_bh = gb1;
    rec->complete = 0;
  }
  if (lane < JAXENT_MAX_SLOTS) {
    rec->train[lane] = (lane < n_slots) ? (float)s_fin[lane] : 0.f;
    rec->val[lane] = (lane < n_slots) ? (float)s_fin[JAXENT_MAX_SLOTS + lane] : 0.f;
  }
  __syncwarp();
  if (lane == 0 && s_old.kl_slot < 0) finish_record(neu, 0.0, records, n_slots);  // no MaxEnt slot: the record is complete here
}

// The first frame CTA (block 1) writes what derive_rest produced: the speculative learning rates, bias corrections and
// log-sum-exp of the coming update, and the convergence tracker (+ the snapshot header).  Lanes 0 and 2 of one warp.
__device__ __forceinline__ void write_ctl_rest(Ctl* __restrict__ neu, const JaxentLossRecord* records, const StepScalars& sc, int lane,
                                               const TrackConfig& trk) {
  if (lane == 0) {
    neu->lrA = sc.lrA;
    neu->lrB = sc.lrB;
    neu->mlrA = sc.mlrA;
    neu->mlrB = sc.mlrB;
    neu->bc1 = sc.bc1;
    neu->bc2 = sc.bc2;
    neu->lse = sc.lse_new;
  }
  if (lane == 2) {
    neu->trk_stop = sc.trk_stop;
    neu->trk_reason = sc.trk_reason;
    neu->trk_stop_step = sc.trk_stop_step;
    neu->trk_have_prev = sc.trk_have_prev;
    neu->trk_have_ema = sc.trk_have_ema;
    neu->trk_steps_since = sc.trk_steps_since;
    neu->trk_idx = sc.trk_idx;
    neu->trk_nsnap = sc.trk_nsnap;
    neu->trk_ema_bc = sc.trk_ema_bc;
    neu->trk_ema_bh = sc.trk_ema_bh;
    neu->trk_prev_loss = sc.trk_prev_loss;
    neu->trk_ema = sc.trk_ema;
    if (sc.snap_now >= 0 && trk.snaps != nullptr) {
      JaxentTrackSnapshot* sn = trk.snaps + sc.snap_now;
      sn->loop_step = sc.loop_step;
      sn->state_step = sc.step;
      sn->have_ema = sc.trk_have_ema;
      sn->threshold_idx = sc.trk_idx;
      sn->bc = sc.bc;
      sn->bh = sc.bh;
      sn->ema_bc = sc.trk_ema_bc;
      sn->ema_bh = sc.trk_ema_bh;
      sn->losses = records[sc.loop_step % JAXENT_RECORD_RING];
    }
  }
}

// per-frame MaxEnt terms of frames [cta*frames_per_cta, ...): softmax weights, dKL/dw and the KL sums of this CTA
__device__ __forceinline__ void frame_terms(const TailArgs& a, const StepScalars& sc, const Ctl& s_old, int cur_set, long long cta,
                                            long long /*n_frame_ctas*/, double (*s_red)[16]) {
  const JaxentBvProblem& pb = a.prob;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const float* __restrict__ z = a.st.z[cur_set][sc.d];
  const long long f0 = cta * a.frames_per_cta;
  const long long f1 = (f0 + a.frames_per_cta < pb.F) ? f0 + a.frames_per_cta : pb.F;
  const float Sf = (float)sc.S;
  const double invS_d = __drcp_rn(sc.S);
  const float lse_old = sc.lse_old;
  const bool has_kl = s_old.kl_slot >= 0 && pb.prior != nullptr;
  const float eps = a.eps_kl;
  const float lam = s_old.kl_scale;
  const double inv_pnorm = __drcp_rn(pb.prior_norm);
  const float trk_alpha = a.trk.ema_alpha;
#pragma unroll 1
  for (long long f = f0 + tid; f < f1; f += nt) {
    const float ef = expf(z[f] - lse_old);
    const float w = ef / Sf;  // jax.nn.softmax: exp(x - shift) / sum
    a.w[f] = w;
    float kap = 0.f, kap_lo = 0.f;
    if (has_kl) {
      // q = (|w|+eps)/sum(|w|+eps); the denominator is 1 + F*eps = 1 + 1e-10 == 1 in fp32
      const float q = fabsf(w) + eps;
      const float p = (float)(((double)fabsf(pb.prior[f]) + (double)eps) * inv_pnorm);
      // p/q and log(p/q) in fp64: differencing two fp32 logs leaves a rounding error that is
      // identical on every frame of a uniform prior and so does not average out over F
      const double ratio = (double)p * __drcp_rn((double)q);
      // d|w|/dw = sign(w); softmax weights are >= 0.  kappa travels to the pass as an unevaluated fp32 pair: once a few
      // frames dominate, g_f = w_f (H_f + kappa_f - hbar) is a difference of O(lambda) numbers and a single float would
      // put lambda * 2^-24 of noise into it (1.4e-5 of the gradient scale at w_max = 0.997, measured)
      const double kd = (w > 0.f) ? (double)lam * (1.0 - ratio) : 0.0;
      kap = (float)kd;
      kap_lo = (float)(kd - (double)kap);
      // hbar's MaxEnt part with the weights the forward sums carry, e_f / S in fp64 (NOT the rounded fp32 w_f): the data
      // part of hbar is the identity sum_j (e_j / S) H_j, and 2^-24 of inconsistency on a dominant frame times
      // kappa = O(lambda) is 5e-6 of the gradient scale at w_max = 0.997 (measured)
      s0 += ((double)ef * invS_d) * kd;
      if (p > 0.f) s1 += (double)p * dlog(ratio);
      s2 += (double)q - (double)p;
    }
    a.kappa[f] = kap;
    a.kappa_lo[f] = kap_lo;
    s3 += (double)w;
    if (sc.ema_update) {  // tracker.ema_params (opt/track.py:160-165): EMA of the normalised weights
      const float e = sc.ema_first ? w : trk_alpha * w + (1.0f - trk_alpha) * a.trk.ema_w[f];
      a.trk.ema_w[f] = e;
      if (sc.snap_now >= 0) a.trk.snap_ema[(size_t)sc.snap_now * pb.F + f] = e;
    }
    if (sc.snap_now >= 0) a.trk.snap_w[(size_t)sc.snap_now * pb.F + f] = w;
  }
  s0 = wsum(s0);
  s1 = wsum(s1);
  s2 = wsum(s2);
  s3 = wsum(s3);
  __syncthreads();
  if (lane == 0) { s_red[warp][0] = s0; s_red[warp][1] = s1; s_red[warp][2] = s2; s_red[warp][3] = s3; }
  __syncthreads();
  if (tid < KL_N) {
    double t = 0.0;
#pragma unroll 1
    for (int w = 0; w < nw; ++w) t += s_red[w][tid];
    a.klpart[(size_t)cta * KL_N + tid] = t;
  }
}

// last-arriving CTA folds the per-CTA KL sums in CTA order (deterministic) and publishes them: a plain store on one
// rank, stamped 16-byte stores into the kl inbox of every rank (slot = this rank) when frame-sharded
__device__ __forceinline__ void fold_kl_partials(const TailArgs& a, unsigned ep, int* s_last) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned n_frame_ctas = gridDim.x - 1;
  if (n_frame_ctas > 1) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();  // after the barrier: covers every thread's stores of this CTA (w, kappa, klpart, cleared accumulators)
      const unsigned ticket = atomicAdd(a.counter, 1u);
      *s_last = (ticket == n_frame_ctas - 1);
    }
  } else if (tid == 0) {
    *s_last = 1;
  }
  __syncthreads();
  if (*s_last) {
    if (warp < KL_N) {
      __threadfence();
      double t = 0.0;
#pragma unroll 1
      for (unsigned cc = lane; cc < n_frame_ctas; cc += 32) t += __ldcg(&a.klpart[(size_t)cc * KL_N + warp]);
      t = wsum(t);
      if (lane == 0) kl_publish(a.x, ep, warp, t);
      if (tid == 0) *a.counter = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------ completing the pass's sums
// Frame-sharded: the first `world` frame CTAs share the columns; each converts its slice of this rank's sums, clears it and
// stores it into EVERY rank's inbox as stamped 16-byte words (slot = this rank; the own inbox is a destination like any
// other, so every rank adds the same words in the same order).  The CTAs work in parallel -- a single publishing CTA was
// limited by the rate at which one SM can post NVLink stores (16k words for 8 ranks: ~10 us at R = 500).
__device__ __forceinline__ void publish_sums(const TailArgs& a, unsigned ep, int part, int n_parts, int n_rows, double inv) {
  const size_t off = xchg_red_off(a.x, (int)(ep & 1u), a.x.rank);
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int per = (n_rows + n_parts - 1) / n_parts, c_begin = part * per, c_end = min(n_rows, c_begin + per);
  constexpr int CP = 2;  // columns per thread and trip: CP * FX_REP loads in flight
  for (int c0 = c_begin + tid; c0 < c_end; c0 += CP * nt) {
    double v[CP];
#pragma unroll
    for (int k = 0; k < CP; ++k) v[k] = (c0 + k * nt < c_end) ? (double)fx_load_col(a.fx, n_rows, c0 + k * nt) * inv : 0.0;
#pragma unroll
    for (int k = 0; k < CP; ++k) {
      const int col = c0 + k * nt;
      if (col < c_end) {
#pragma unroll
        for (int r = 0; r < JAXENT_MAX_RANKS; ++r)  // compile-time indices: no local-memory copy of the parameter struct
          if (r < a.x.world) ll_store(a.x.base[r] + off + (size_t)col * 16, v[k], ep);
        fx_clear_col(a.fx, n_rows, col);
      }
    }
  }
  if (part == 0 && warp == nw - 1) {  // the three scalars: the last warp of the first publisher
    double t[3];
    fold_scal3(a.cta_scal, a.pass_ctas, lane, t);
    if (lane < 4) {
      const double v = lane == 0 ? t[0] : lane == 1 ? t[1] : lane == 2 ? t[2] : 0.0;
#pragma unroll
      for (int r = 0; r < JAXENT_MAX_RANKS; ++r)
        if (r < a.x.world) ll_store(a.x.base[r] + off + (size_t)(n_rows + lane) * 16, v, ep);
    }
  }
}

// Three-phase ABI (jaxent_bv_reduce): the sums as plain doubles in JAXENT_BUF_RED, for a caller that all-reduces them itself
// (NCCL) between the pass and the tail.
__global__ void __launch_bounds__(1024) reduce_fold_kernel(unsigned long long* fx, const double* cta_scal, const float* nmax,
                                                           const Ctl* ctl, const unsigned* epoch, int per_frame_mode, long long F,
                                                           int n_rows, int pass_ctas, double* red) {
  const unsigned ep = *epoch;  // the epoch the pass started in (the tail advances it)
  const Ctl* c = ctl + (ep & 1u);
  if (c->trk_stop) return;
  const double inv = fx_inv_scale(*nmax, per_frame_mode != 0, c->pending, F);
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int col = tid; col < n_rows; col += nt) {
    red[col] = (double)fx_load_col(fx, n_rows, col) * inv;
    fx_clear_col(fx, n_rows, col);
  }
  if (warp == nw - 1) {
    double t[3];
    fold_scal3(cta_scal, pass_ctas, lane, t);
    if (lane < 4) red[n_rows + lane] = lane == 0 ? t[0] : lane == 1 ? t[1] : lane == 2 ? t[2] : 0.0;
  }
}

int launch_reduce_fold(const TailArgs& a, cudaStream_t s) {
  reduce_fold_kernel<<<1, 1024, 0, s>>>(a.fx, a.cta_scal, a.nmax, a.ctl, a.epoch, a.prob.uptake == 2, a.prob.F, a.part_n - 4, a.pass_ctas,
                                        reinterpret_cast<double*>(a.x.local + xchg_plain_red_off(a.x)));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("reduce_fold launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

// ------------------------------------------------------------------------------------------ the tail
// GENERAL = the mean-centred / sigma MSE kinds are compiled in (host picks the instantiation), so the
// common eye-MSE / PF-L2 configuration executes the shortest possible instruction stream.
template <bool GENERAL>
__global__ void __launch_bounds__(TAIL_THREADS, 1) bv_tail_kernel(const TailArgs a) {
  unsigned char* const smem_raw = jx_dyn_smem;
  __shared__ double s_red[32][16];  // per-warp partials of the multi-value reductions
  __shared__ double s_fin[16];
  __shared__ __align__(16) Ctl s_ctl2[2];  // both control blocks, fetched before the epoch is known
  __shared__ StepScalars s_sc;
  __shared__ SpecScalars s_spec;
  __shared__ int s_last;
  // frame-sharded gather of the 4R row sums, spread over all warps (GATHER_ROWS_MAX rows; larger R keeps one thread per row)
  constexpr int GATHER_ROWS_MAX = 512;
  __shared__ double s_cols[4 * GATHER_ROWS_MAX];

  const JaxentBvProblem& pb = a.prob;
  const int R = pb.R, T = pb.uptake ? pb.T : 0;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const bool tail_cta = blockIdx.x == 0;  // block 0: the R x T tail; blocks 1..: per-frame MaxEnt terms (the grid always has >= 2 blocks)

  const TailSmem tsm = carve_tail_smem(smem_raw, R, T, a.p_max_tr, a.p_max_va);

  TPROF(0); TPROF_F(0);
  // ---- Everything that does not depend on the pass runs BEFORE the grid-dependency wait (the tail is launched while the
  // pass streams, PDL): the peptide-map blob, rates and time points, BOTH control blocks (written by the previous tail, which
  // had completed before the pass -- and hence this kernel -- could start), and the BV-parameter update for both plateau branches.
  int first_data = -1;
  if (tail_cta) first_data = stage_tail_constants(pb, a.blobs, tsm);
  {
    constexpr int CW = (int)(sizeof(Ctl) / 16);
    if (tid >= 64 && tid < 64 + 2 * CW) reinterpret_cast<int4*>(s_ctl2)[tid - 64] = __ldcg(reinterpret_cast<const int4*>(a.ctl) + (tid - 64));
    if (tail_cta && lane < 16) s_red[warp][lane] = 0.0;
  }
  __syncthreads();  // (0) control blocks staged
  // the block that carries the larger epoch is the current one; this step's epoch follows from it (no load after the wait)
  const int newest = ((int)(s_ctl2[1].epoch_id - s_ctl2[0].epoch_id) > 0) ? 1 : 0;
  const Ctl& s_old = s_ctl2[newest];
  if (s_old.trk_stop) return;  // the on-device loop has ended (the pass did not advance the epoch either): no-op
  const unsigned ep = (unsigned)s_old.epoch_id + 1u;
  Ctl* __restrict__ neu = a.ctl + (newest ^ 1);
  const int cur_set = ep & 1u;
  if (warp == 0) derive_spec(&s_spec, &s_old, a.cfg);
  TPROF(1);
  pdl_wait();               // the reduced partial sums are visible from here on
  pdl_launch_dependents();  // lets the next pass kernel's launch + TMA prologue overlap this kernel
  TPROF(2); TPROF_F(1);

  // ---- the pass's sums.  Single rank: every reader folds the fixed-point accumulators itself (all loads of a thread in
  // flight at once, one L2 round trip; no publishing step in between).  Frame-sharded: the first `world` frame CTAs publish
  // this rank's sums to the ranks' inboxes, then everybody gathers all ranks' words (red_total spins on late ones).
  const bool pfm = pb.uptake == 2;
  const int n_rows = a.part_n - 4;
  const bool direct = a.fold_here && a.x.world == 1;
  const double fx_inv = a.fold_here ? fx_inv_scale(__ldg(a.nmax), pfm, s_old.pending, pb.F) : 0.0;
  if (a.fold_here && a.x.world > 1 && !tail_cta) {
    const int nf = (int)gridDim.x - 1, n_parts = nf < a.x.world ? nf : a.x.world;
    if ((int)blockIdx.x - 1 < n_parts) publish_sums(a, ep, (int)blockIdx.x - 1, n_parts, n_rows, fx_inv);
  }
  // single rank: the R x T tail CTA is the one reader of every column (it clears what it has read)
  auto sum_col = [&](int j) -> double {
    if (!direct) return red_total(a.x, ep, j);
    const double v = (double)fx_load_col(a.fx, n_rows, j) * fx_inv;
    fx_clear_col(a.fx, n_rows, j);
    return v;
  };
  double* red_plain = reinterpret_cast<double*>(a.x.local + xchg_plain_red_off(a.x));  // JAXENT_BUF_RED (single rank)
  double accA0 = 0, accA1 = 0, accB0 = 0, accB1 = 0;
  const bool spread_gather = tail_cta && !direct && !pfm && R <= GATHER_ROWS_MAX && nt > 32;
  if (tail_cta && tid < R && !pfm) {
    if (direct) {  // all 4 * FX_REP loads in flight, then the clears
      const long long t0 = fx_load_col(a.fx, n_rows, 0 * R + tid), t1 = fx_load_col(a.fx, n_rows, 1 * R + tid);
      const long long t2 = fx_load_col(a.fx, n_rows, 2 * R + tid), t3 = fx_load_col(a.fx, n_rows, 3 * R + tid);
      accA0 = (double)t0 * fx_inv;
      accA1 = (double)t1 * fx_inv;
      accB0 = (double)t2 * fx_inv;
      accB1 = (double)t3 * fx_inv;
#pragma unroll
      for (int k = 0; k < 4; ++k) fx_clear_col(a.fx, n_rows, k * R + tid);
      red_plain[0 * R + tid] = accA0;
      red_plain[1 * R + tid] = accA1;
      red_plain[2 * R + tid] = accB0;
      red_plain[3 * R + tid] = accB1;
    } else if (!spread_gather) {
      accA0 = red_total(a.x, ep, 0 * R + tid);
      accA1 = red_total(a.x, ep, 1 * R + tid);
      accB0 = red_total(a.x, ep, 2 * R + tid);
      accB1 = red_total(a.x, ep, 3 * R + tid);
    }
  }
  if (spread_gather && warp > 0) {
    // one poll of a column = all ranks' words in flight, one L2 round trip once they have landed.  A thread per row did four of
    // them one after the other and warp 0 three more for the scalars (7 round trips on the critical path of every sharded
    // step); here the 4R columns are spread over warps 1..31 (2-3 each) while warp 0 polls the three scalars side by side.
    for (int j = tid - 32; j < 4 * R; j += nt - 32) s_cols[j] = red_total(a.x, ep, j);
  }
  if (warp == 0) {
    double SA, SB, gdot;
    if (direct) {
      double t3[3];
      fold_scal3(a.cta_scal, a.pass_ctas, lane, t3);
      SA = t3[0];
      SB = t3[1];
      gdot = t3[2];
      if (tail_cta && lane == 0) {
        red_plain[n_rows + 0] = SA;
        red_plain[n_rows + 1] = SB;
        red_plain[n_rows + 2] = gdot;
        red_plain[n_rows + 3] = 0.0;
      }
    } else {  // lanes 0..2 poll one scalar each (they spin independently), then the warp shares them
      double v = 0.0;
      if (lane < 3) v = red_total(a.x, ep, n_rows + lane);
      __syncwarp();
      SA = __shfl_sync(0xffffffffu, v, 0);
      SB = __shfl_sync(0xffffffffu, v, 1);
      gdot = __shfl_sync(0xffffffffu, v, 2);
    }
    __syncwarp();
    if (lane == 0) derive_critical(&s_sc, &s_spec, &s_old, SA, SB, gdot, a.cfg);
    // the per-frame CTAs need the tracker's decisions (snapshot slot, EMA update) before their frame loop
    if (!tail_cta) {
      __syncwarp();
      derive_rest(&s_sc, &s_old, a.cfg, a.trk, a.records);
    }
  }
  __syncthreads();  // (1) scalars broadcast (and the gathered columns)
  TPROF(3); TPROF_F(2);
  const int d = s_sc.d;
  const double S = s_sc.S;
  if (spread_gather && tid < R) {
    accA0 = s_cols[0 * R + tid];
    accA1 = s_cols[1 * R + tid];
    accB0 = s_cols[2 * R + tid];
    accB1 = s_cols[3 * R + tid];
  }

  // ================================================================== the R x T tail (CTA 0)
  if (tail_cta) {
    const double bc = (double)s_sc.bc, bh = (double)s_sc.bh;
    double my_c, my_a, my_b, my_lp, gbc_reg, gbh_reg;
    TailIO io;
    io.avg = a.avg;
    io.logpf = a.logpf;
    io.uptake = a.uptake;
    io.pf_G = a.pf_G;
    io.pf_mode = pfm ? 1 : 0;
    io.p_max_tr = a.p_max_tr;
    io.p_max_va = a.p_max_va;
    io.c_pieces = 0;
    io.row_off = nullptr;
    io.row_off_stride = 0;
    io.prof = reinterpret_cast<long long*>(a.cl_scratch);
    io.ck = a.ck;
    if (pfm) {  // per-frame uptake mode: <u>[t,r] = 1 - acc[branch d][t][r] / S, acc = sum_f e'_f exp(-k tau / PF_f) from the pass
      const double invS = __drcp_rn(S);
#pragma unroll 1
      for (int it = tid; it < R * T; it += nt) {
        const double sv = sum_col(d * T * R + it);
        if (direct) {
          red_plain[d * T * R + it] = sv;
          fx_clear_col(a.fx, n_rows, (d ^ 1) * T * R + it);  // the branch not taken is read by nobody: cleared here
        }
        const double uv = 1.0 - sv * invS;
        const int t = it / R, r = it - t * R;
        tsm.sU[r * T + t] = uv;  // [r][t], as tail_body reads it
        a.uptake[it] = (float)uv;
      }
    }
#ifdef JX_ICACHE_TEST
#pragma unroll 1
    for (int rep = 0; rep < 2; ++rep) {
      if (rep == 1) { __syncthreads(); if (lane < 16) s_red[warp][lane] = 0.0; __syncthreads(); TPROF(13); }
#endif
    tail_body<GENERAL>(pb, a.blobs, first_data, tsm, d ? accB0 : accA0, d ? accB1 : accA1, S, bc, bh, s_old.fw, s_old.fs, io, s_red, s_fin,
                       my_c, my_a, my_b, my_lp, gbc_reg, gbh_reg);
#ifdef JX_ICACHE_TEST
      if (rep == 0) TPROF(14);
    }
#endif
    if (tid < R && !pfm) {  // row coefficients of the next pass
      a.cc[tid] = (float)(my_c * bc);
      a.ch[tid] = (float)(my_c * bh);
    }
    TPROF(9);
    if (tid < 32) write_ctl_critical(neu, a.records, s_sc, s_old, s_fin, gbc_reg, gbh_reg, pb.n_slots, lane, ep);
    if (tid == 0) *a.epoch = ep;  // the next pass starts after this kernel has completed and reads the new control block
    TPROF(10);
    return;
  }

  // ================================================================== per-frame MaxEnt terms (blocks 1..)
  // The last frame CTA to finish publishes the MaxEnt sums, typically well before the R x T tail above has finished.
  if (blockIdx.x == 1 && warp == 0) write_ctl_rest(neu, a.records, s_sc, lane, a.trk);
  frame_terms(a, s_sc, s_old, cur_set, (long long)blockIdx.x - 1, 0, s_red);
  TPROF_F(3);
  fold_kl_partials(a, ep, &s_last);
  TPROF_F(4);
}

size_t tail_smem_bytes(int R, int T, int p_tr, int p_va, int nnz_tr, int nnz_va) {
  const size_t Tm = T > 0 ? T : 1;
  const size_t dbl = (size_t)4 * R + (size_t)Tm * R + (size_t)T * R + 2 * (size_t)p_tr * Tm + (size_t)p_va * Tm + 4 * Tm + R + Tm;
  const BlobLayout b = blob_layout(R, (int)Tm, p_tr, nnz_tr, p_va, nnz_va);
  return dbl * 8 + (size_t)b.words * 4 + 64;
}

int blob_words(int R, int T, int p_tr, int nnz_tr, int p_va, int nnz_va) {
  return blob_layout(R, T > 0 ? T : 1, p_tr, nnz_tr, p_va, nnz_va).words;
}

int launch_tail(const TailArgs& a, int ctas, size_t smem, cudaStream_t s) {
  bool general = false;
  for (int i = 0; i < a.prob.n_slots; ++i)
    general = general || a.prob.slots[i].kind == JAXENT_LOSS_UPTAKE_MC_EYE_MSE || a.prob.slots[i].kind == JAXENT_LOSS_UPTAKE_SIGMA_MSE;
  static size_t configured[2] = {0, 0};
  if (smem > configured[general]) {
    cudaError_t e = general ? cudaFuncSetAttribute(bv_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(bv_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("tail kernel needs %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return JAXENT_E_UNSUPPORTED; }
    configured[general] = smem;
  }
  cudaError_t e = general ? launch_pdl(bv_tail_kernel<true>, ctas, TAIL_THREADS, smem, s, a)
                          : launch_pdl(bv_tail_kernel<false>, ctas, TAIL_THREADS, smem, s, a);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bv_tail_kernel launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

}  // namespace jx
