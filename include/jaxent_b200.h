/*
 * jaxent_b200.h -- C ABI of the B200-native BV MaxEnt optimise step.
 *
 * Drop-in boundary for the one hot path of alexisiddiqui/JAX-ENT (paths below are relative to
 * /root/reference/jaxent/src):
 *
 *   OptaxOptimizer._step_with_rates   opt/optimiser.py:337-445   (value_and_grad + mask + plateau LR + optax)
 *   compute_loss                      opt/optimiser.py:608-644
 *   Simulation.forward / forward_pure models/core.py:131-163,270-307
 *   frame_average_features            utils/jax_fn.py:15-38
 *   BV_ForwardPass / BV_uptake_ForwardPass  models/HDX/forward.py:15-76
 *   apply_sparse_mapping              data/splitting/sparse_map.py:220-235
 *   hdx_uptake_{eye,mean_centred_eye,sigma}_MSE_loss, hdx_pf_l2_loss, maxent_convexKL_loss,
 *   model_params_L1/L2_loss           opt/losses.py:1541-1657,1790-1841,17-50,304-322,476-513
 *
 * The reference has no FFI of its own (it is pure Python on jax/XLA); these entry points are what
 * a jax.ffi custom call for the path binds (see INTEGRATION.md and csrc/xla_ffi.cc).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / jax types.  All device buffers are caller owned; the
 *     library never calls cudaMalloc/cudaFree, never synchronises the host and only enqueues work
 *     on the stream it is given.
 *   - every function returns 0 on success or a negative JAXENT_E_* code; jaxent_last_error()
 *     returns a thread-local message for the last failure.
 *   - there is no CPU fallback: without a CUDA device the compute entry points fail.
 */
#ifndef JAXENT_B200_H_
#define JAXENT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* jaxent_stream_t; /* cudaStream_t */

#define JAXENT_OK 0
#define JAXENT_E_INVALID (-1)   /* bad shape / alignment / argument   (reference: ValueError, models/core.py:65-66) */
#define JAXENT_E_UNSUPPORTED (-2) /* unsupported loss / forward combination (raised at plan build) */
#define JAXENT_E_CUDA (-3)      /* launch / runtime error             (reference: RuntimeError, opt/run.py:350-353) */
#define JAXENT_E_WORKSPACE (-4) /* workspace too small */

/* loss slot kinds; one per (loss_function, data_target) pair of run_optimise (opt/run.py:376-406) */
enum {
  JAXENT_LOSS_UPTAKE_EYE_MSE = 0,    /* hdx_uptake_eye_MSE_loss               opt/losses.py:1601 */
  JAXENT_LOSS_UPTAKE_MC_EYE_MSE = 1, /* hdx_uptake_mean_centred_eye_MSE_loss  opt/losses.py:1790 */
  JAXENT_LOSS_UPTAKE_SIGMA_MSE = 2,  /* hdx_uptake_sigma_MSE_loss             opt/losses.py:1541 */
  JAXENT_LOSS_PF_L2 = 3,             /* hdx_pf_l2_loss                        opt/losses.py:17   */
  JAXENT_LOSS_MAXENT_CONVEX_KL = 4,  /* maxent_convexKL_loss                  opt/losses.py:304  */
  JAXENT_LOSS_PARAMS_L1 = 5,         /* model_params_L1_loss                  opt/losses.py:495  */
  JAXENT_LOSS_PARAMS_L2 = 6          /* model_params_L2_loss                  opt/losses.py:476  */
};

#define JAXENT_MAX_SLOTS 6

/* Peptide<-residue map in CSR (rows = peptides) plus its transpose in CSC order (rows = residues),
 * both on the device.  Same linear map as the reference's dense `sparse_map.todense() @ x`. */
typedef struct {
  int32_t n_pep;
  int32_t nnz;            /* number of stored entries (row_ptr[n_pep]) */
  const int32_t* row_ptr; /* (n_pep+1) */
  const int32_t* col;     /* (nnz) residue index   */
  const float* val;       /* (nnz)                 */
  const int32_t* t_ptr;   /* (R+1)   transpose     */
  const int32_t* t_row;   /* (nnz) peptide index   */
  const float* t_val;     /* (nnz)                 */
  const float* y;         /* uptake kinds: (n_pep, T) row-major [Dataset.y_true (P,T,1)]; PF: (n_pep) */
  const float* W;         /* sigma kind: (n_pep,n_pep) trace-normalised precision, else NULL */
} JaxentPeptideMap;

typedef struct {
  int32_t kind;
  JaxentPeptideMap train; /* data kinds only */
  JaxentPeptideMap val;
  float prior_bc, prior_bh; /* PARAMS_L1/L2 */
} JaxentLossSlot;

/* Static description of one rank's shard of the problem. */
typedef struct {
  int32_t R;            /* residues */
  int32_t T;            /* time points (uptake mode) */
  int32_t uptake;       /* 1: BV_uptake_ForwardPass, 0: BV_ForwardPass (log_Pf), 2: BV_uptake_ForwardPass with
                           average_first = False (uptake per frame, then frame-averaged; models/core.py:302-304): fp32 features,
                           R <= 512, T <= 8; optimise_model_parameters only on an unsharded ensemble (F_total == F) */
  int32_t n_slots;
  int64_t F;            /* frames held by this rank */
  int64_t F_total;      /* frames of the whole ensemble (eps = 1e-10 / F_total, opt/losses.py:308) */
  const float* packed;  /* packed features, jaxent_bv_packed_bytes() bytes, 16-byte aligned */
  const float* k_ints;  /* (R)  uptake mode */
  const float* timepoints; /* (T) uptake mode */
  const float* prior;   /* (F) this rank's slice of the MaxEnt prior frame weights, or NULL */
  double prior_norm;    /* sum over ALL ranks of (|prior_f| + eps); see jaxent_bv_prior_partial() */
  JaxentLossSlot slots[JAXENT_MAX_SLOTS];
  int32_t feature_format; /* layout of `packed`: 0 = fp32 (jaxent_bv_pack_features_f32), 1 = u8 (jaxent_bv_pack_features_u8) */
  int32_t pad_;
} JaxentBvProblem;

/* OptaxOptimizer.__init__ arguments (opt/optimiser.py:68-82) */
typedef struct {
  float learning_rate;
  float initial_learning_rate;
  int32_t initial_steps;
  float plateau_denominator;
  float model_parameters_lr_scale;
  float clip_value;             /* <= 0: no clipping (clip_value=None) */
  int32_t optimise_frame_weights;
  int32_t optimise_model_parameters;
  double b1, b2, eps;           /* optax.adam defaults 0.9 / 0.999 / 1e-8 (doubles: optax forms 1-b1 in Python floats) */
} JaxentOptConfig;

/* LossComponents (opt/base.py:61-69) of one step, plus the step's scalars. */
typedef struct {
  int32_t step;      /* state.step the loss was evaluated at */
  int32_t branch;    /* 1 if the plateau rule reduced the LR for the update that led here */
  float lr;          /* learning rate used by that update */
  float grad_dot;    /* sum vdot(prev_grads, grads) of that update (opt/optimiser.py:398-401) */
  float bc, bh;      /* BV parameters the loss was evaluated at */
  float total_train, total_val;
  float train[JAXENT_MAX_SLOTS], val[JAXENT_MAX_SLOTS];
  float scaled_train[JAXENT_MAX_SLOTS], scaled_val[JAXENT_MAX_SLOTS];
  float grad_bc, grad_bh; /* masked d loss / d bc, bh at this step */
  int32_t complete;  /* 1 once the MaxEnt term has been folded in */
  int32_t pad;
} JaxentLossRecord;

/* One history snapshot of the on-device loop: what OptimizationHistory.add_state(save_state) keeps
 * (opt/run.py:318-348), minus the frame-sized arrays, which live in the caller's snapshot buffers. */
typedef struct {
  int32_t loop_step;   /* loop index k of the step that met the threshold (0 for the initial snapshot) */
  int32_t state_step;  /* save_state.step = k + 1 */
  int32_t have_ema;    /* EMA parameters exist (tracker.ema_params is not None) */
  int32_t threshold_idx;
  float bc, bh;        /* BV parameters of the saved state */
  float ema_bc, ema_bh;
  JaxentLossRecord losses; /* LossComponents of the saved state (evaluated at step k) */
} JaxentTrackSnapshot;

/* status of the on-device loop */
typedef struct {
  int32_t stop;        /* 1 once the loop has ended; further step launches are no-ops */
  int32_t reason;      /* 1 all relative thresholds completed, 2 tolerance reached or non-finite loss */
  int32_t stop_step;   /* loop index of the step that ended the loop */
  int32_t n_snapshots;
  int32_t steps_done;  /* state.step: optimiser updates applied */
  int32_t threshold_idx;
  float last_loss;
  float ema_loss_delta;
} JaxentTrackStatus;

#define JAXENT_TRACK_MAX_THRESHOLDS 8

typedef struct JaxentPlan JaxentPlan; /* opaque, host memory */

/* identifiers for jaxent_bv_plan_buffer() */
enum {
  JAXENT_BUF_RED = 0,      /* double[4R+4]  per-rank partial sums -> all-reduce(sum) #1 across ranks   */
  JAXENT_BUF_KLSUM = 1,    /* double[4]     per-rank MaxEnt sums  -> all-reduce(sum) #2 across ranks   */
  JAXENT_BUF_WEIGHTS = 2,  /* float[F]      softmax(z) of the current step (save_state.params.frame_weights) */
  JAXENT_BUF_RECORDS = 4,  /* JaxentLossRecord[JAXENT_RECORD_RING]                                      */
  JAXENT_BUF_LOGPF = 5,    /* float[R]      log_Pf of the current step                                  */
  JAXENT_BUF_UPTAKE = 6,   /* float[T*R]    uptake (T,R) of the current step (uptake mode)              */
  JAXENT_BUF_AVG = 7,      /* float[2R]     frame-averaged heavy / acceptor contacts                    */
  JAXENT_BUF_XCHG_ERR = 9  /* uint64[1]     non-zero once a wait of the NVLink exchange timed out               */
};
#define JAXENT_RECORD_RING 4096

const char* jaxent_last_error(void);
int jaxent_version(void);
/* sizeof of the ABI structs, for binding self-checks: 0 JaxentBvProblem, 1 JaxentOptConfig, 2 JaxentLossRecord,
 * 3 JaxentLossSlot, 4 JaxentPeptideMap; -1 for an unknown id */
int jaxent_abi_sizeof(int which);

/* ---- feature packing (one-off; replaces Input_Features.cast_to_jax, custom_types/features.py:82-89) ---- */
/* bytes of the packed layout for (R,F): [ceil(F/8)*2 quads][2 arrays][R rows][4 frames] fp32 */
size_t jaxent_bv_packed_bytes(int32_t R, int64_t F);
/* heavy / acceptor: (R, F) row-major fp32 on the device with leading dimension ld (>= F). */
int jaxent_bv_pack_features_f32(const float* heavy, const float* acceptor, int32_t R, int64_t F, int64_t ld,
                                int64_t frame_offset, int64_t F_dst, float* packed, jaxent_stream_t stream);

/* Lossless u8 storage of the features (hard-cutoff BV contacts are integer counts, reference
 * jaxent/src/models/func/contacts.py:236-240): a quarter of the HBM traffic per step, bit-identical results.
 * layout [ceil(F/32)*4 tiles of 8 frames][2 arrays][R rows][8 bytes].  features_fit_u8 clears *flag (device int,
 * preset to 1 by the caller) if any value is not an integer in [0,255]. */
size_t jaxent_bv_packed_bytes_fmt(int32_t R, int64_t F, int32_t format);
int jaxent_bv_features_fit_u8(const float* heavy, const float* acceptor, int32_t R, int64_t F, int64_t ld, int32_t* flag,
                              jaxent_stream_t stream);
int jaxent_bv_pack_features_u8(const float* heavy, const float* acceptor, int32_t R, int64_t F, int64_t ld,
                               unsigned char* packed, jaxent_stream_t stream);

/* per-rank partial of sum_f (|prior_f| + eps) written to out[0] (device double) */
int jaxent_bv_prior_partial(const float* prior, int64_t F, double eps, double* out, jaxent_stream_t stream);
/* max_f z_f written to out[0] (device float); used once for the initial log-sum-exp shift */
int jaxent_bv_max(const float* z, int64_t F, float* out, jaxent_stream_t stream);

/* ---- plan ---- */
size_t jaxent_bv_workspace_bytes(const JaxentBvProblem* problem);
int jaxent_bv_plan_create(const JaxentBvProblem* problem, const JaxentOptConfig* cfg, void* workspace,
                          size_t workspace_bytes, int32_t device_sm_count, JaxentPlan** plan);
void jaxent_bv_plan_destroy(JaxentPlan* plan);
int jaxent_bv_plan_buffer(const JaxentPlan* plan, int32_t which, void** ptr, size_t* bytes);

/* ---- frame-sharded exchange over NVLink peer memory ----
 * Without it a sharded run all-reduces JAXENT_BUF_RED / JAXENT_BUF_KLSUM between the phases (NCCL).  With it the
 * two exchanges are fused into the kernels: `reduce` / `tail` store this rank's partial sums straight into every
 * peer's inbox as self-validating 16-byte words (value + epoch stamp, the NCCL "LL" idea: no flags, no fences), the
 * consumers (`tail`, the next `pass`) spin on the words of their local inbox and add the slots in rank order, so
 * every rank computes bit-identical sums and a step is three kernel launches on any number of GPUs.
 * The exchange buffer of each rank (jaxent_bv_exchange_bytes) must be mapped into every process: jaxent_peer_alloc
 * (cudaMalloc + zero fill + cudaIpcGetMemHandle) on the owner, jaxent_peer_open (cudaIpcOpenMemHandle) on the
 * peers -- set-up time only; the step itself never allocates.  peers[world] lists the mapped base pointers in rank
 * order (peers[rank] is the local buffer).  Call before jaxent_bv_state_init; every rank must then complete
 * jaxent_bv_state_init (which clears the local inbox) before any rank launches the first pass. */
#define JAXENT_MAX_RANKS 8
#define JAXENT_IPC_HANDLE_BYTES 64
size_t jaxent_bv_exchange_bytes(int32_t R, int32_t world);
/* same, for any forward mode of `problem` (the per-frame uptake mode exchanges 2*T*R + 4 doubles instead of 4*R + 4) */
size_t jaxent_bv_exchange_bytes_for(const JaxentBvProblem* problem, int32_t world);
int jaxent_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle);
int jaxent_peer_open(const unsigned char* handle, void** dev_ptr);
int jaxent_peer_close(void* dev_ptr);
int jaxent_peer_free(void* dev_ptr);
int jaxent_bv_plan_set_exchange(JaxentPlan* plan, int32_t rank, int32_t world, void* const* peers);
/* copies the exchange error word (JAXENT_BUF_XCHG_ERR) to out[0] (device uint64) on the stream */
int jaxent_bv_exchange_error(JaxentPlan* plan, uint64_t* out, jaxent_stream_t stream);

/* ---- state ----
 * OptaxOptimizer.initialise (opt/optimiser.py:221-304): z0 are the logits (w_init * F_total), zmax
 * the max of z0 over ALL ranks, bc/bh the BV parameters, fw/fs the (already normalised)
 * forward_model_weights / forward_model_scaling.  Zeroes the Adam moments. */
int jaxent_bv_state_init(JaxentPlan* plan, const float* z0, float zmax, float bc, float bh, const float* fw_host,
                         const float* fs_host, jaxent_stream_t stream);

/* ---- the step, in the three phases a frame-sharded run interleaves with its two all-reduces ----
 *   pass   : one streaming pass over the packed features: backward of step k (column sums),
 *            mask, Adam on the frame logits for both plateau-LR branches, forward of step k+1
 *            (softmax-weighted row sums).  mode 0 = forward only (used once after state_init).
 *   reduce : per-CTA partials -> JAXENT_BUF_RED (deterministic order)
 *            [all-reduce(sum) JAXENT_BUF_RED across ranks here]
 *   tail   : plateau decision, BV-parameter Adam, ln PF / uptake / peptide map / losses and their
 *            backward (the R x T tail), MaxEnt per-frame terms -> JAXENT_BUF_KLSUM
 *            [all-reduce(sum) JAXENT_BUF_KLSUM across ranks here]
 */
int jaxent_bv_pass(JaxentPlan* plan, int32_t mode, jaxent_stream_t stream);
int jaxent_bv_reduce(JaxentPlan* plan, jaxent_stream_t stream);
int jaxent_bv_tail(JaxentPlan* plan, jaxent_stream_t stream);
/* folds the (all-reduced) MaxEnt sums into the newest loss record without running a pass */
int jaxent_bv_finish_record(JaxentPlan* plan, jaxent_stream_t stream);

/* single-rank conveniences: init = pass(0)+reduce+tail ; step = pass(1)+reduce+tail */
int jaxent_bv_forward_init(JaxentPlan* plan, jaxent_stream_t stream);
int jaxent_bv_step(JaxentPlan* plan, jaxent_stream_t stream);

/* ---- on-device optimise loop (replaces the host-synchronising loop of _optimise, opt/run.py:235-373) ----
 * Call before jaxent_bv_state_init.  thresholds: n (<= 8) relative-convergence thresholds, descending and already
 * multiplied by the learning rate (create_convergence_thresholds, opt/track.py:12-21).  ema_w: float[F];
 * snap_w, snap_ema: float[(n+1)*F]; snaps: JaxentTrackSnapshot[n+1] -- all caller-owned device memory.
 * Every tail launch then updates the ConvergenceTracker state, the EMA parameters, takes the history snapshots
 * and raises the stop flag exactly where the reference loop would `break`. */
int jaxent_bv_track_init(JaxentPlan* plan, const float* thresholds, int32_t n, float ema_alpha, int32_t min_steps_per_threshold,
                         float tolerance, float* ema_w, float* snap_w, float* snap_ema, JaxentTrackSnapshot* snaps);
/* writes a JaxentTrackStatus to `status` (device memory) */
int jaxent_bv_track_status(JaxentPlan* plan, JaxentTrackStatus* status, jaxent_stream_t stream);

/* copy the current optimiser state (chosen plateau branch) to caller buffers; any pointer may be NULL.
 * z/m/v: float[F] logits and Adam moments; g: float[F] masked d loss / d z of the last completed update;
 * scalars: float[8] = {step, lr, model_lr, mask_idx, bc, bh, count_frame, count_model} */
int jaxent_bv_export_state(JaxentPlan* plan, float* z, float* m, float* v, float* g, float* scalars,
                           jaxent_stream_t stream);

/* ---- BV featuriser (calc_BV_contacts_universe, models/func/contacts.py:125-243; BV_model.featurise,
 * models/HDX/BV/forwardmodel.py:214-300): per frame, for every target atom the number of contact atoms within `radius`
 * (use_switch = 0) or the sum of the rational switch 1/(1 + (r/r0)^6) over the atoms within the radius (use_switch = 1),
 * skipping contact atoms whose residue id is in [target_resid + ignore_lo, target_resid + ignore_hi] (default -2..2).
 * coords: float[F][A][3] positions (Angstrom); target_idx / contact_idx: atom indices into a frame; box: float[F][3]
 * orthorhombic box lengths for the minimum-image convention, or NULL; out: float[Nt][F] (frame axis last).  All device memory. */
int jaxent_bv_contacts(const float* coords, int64_t F, int32_t A, const int32_t* target_idx, const int32_t* target_resid, int32_t Nt,
                       const int32_t* contact_idx, const int32_t* contact_resid, int32_t Nc, int32_t ignore_lo, int32_t ignore_hi,
                       float radius, int32_t use_switch, const float* box, float* out, jaxent_stream_t stream);
/* The same with a general (triclinic) cell -- the reference passes `box=universe.dimensions` to MDAnalysis' distance_array
 * (contacts.py:210-215), which applies the minimum image in whatever cell the trajectory has.  box_stride = 3: box is
 * float[F][3] orthorhombic lengths (as above); box_stride = 6: float[F][6] = {a_x, b_x, b_y, c_x, c_y, c_z}, the non-zero
 * elements of the lower-triangular cell matrix (MDAnalysis triclinic_vectors).  radius must be below half the smallest
 * diagonal element (one image at most within the radius). */
int jaxent_bv_contacts_cell(const float* coords, int64_t F, int32_t A, const int32_t* target_idx, const int32_t* target_resid, int32_t Nt,
                            const int32_t* contact_idx, const int32_t* contact_resid, int32_t Nc, int32_t ignore_lo, int32_t ignore_hi,
                            float radius, int32_t use_switch, const float* box, int32_t box_stride, float* out, jaxent_stream_t stream);

/* ---- per-frame forward pass (Simulation.predict, models/core.py:181-208): the forward pass on the un-averaged features.
 * packed: fp32 tile layout (jaxent_bv_pack_features_f32).  T = 0: out = log_Pf float[R][F] (BV_ForwardPass);
 * T > 0: out = uptake float[T][R][F] (BV_uptake_ForwardPass).  Row-major, frame axis last, as in the reference. */
int jaxent_bv_predict_frames(const float* packed, int32_t R, int64_t F, int32_t T, const float* k_ints, const float* timepoints,
                             float bc, float bh, float* out, jaxent_stream_t stream);

/* ---- batched (replicate / hyper-parameter sweep) mode: tensor-core building blocks (csrc/bg_gemm.cu) ----
 * With B replicas sharing the features (reference opt/batch.py:51-192 vmaps _optimise_pure over the hyper-parameter
 * batch) the two sweeps of a step are dense GEMMs on the tcgen05 tensor cores:
 *   gemm1:  H[f, b]            = sum_{a, r} N_a[r, f] * c_a[r, b]     (backward: transpose product)
 *   gemm2:  P[ks][a*Rp + r, b] = sum_{f in split ks} N_a[r, f] * e[f, b]   (forward: frame_average_features)
 * Layouts (all bf16, byte images of the UMMA no-swizzle canonical layout; Rp = R padded to 128, Fp = F padded to 128,
 * Bn = replicas padded to 16, NS = number of bf16 pieces the fp32 operand is split into):
 *   features  [Fp/128][2][Rp/8][16][8 rows][8 frames]              jaxent_bg_pack_features
 *   cc3       [NS][2][Rp/8][Bn][8 residues]
 *   e3        [Fp/128][NS][16][Bn][8 frames]
 * exact_flag (device int preset to 1, may be NULL) is cleared when a feature value is not exactly representable in bf16. */
size_t jaxent_bg_features_bytes(int32_t R, int64_t F);
int jaxent_bg_pack_features(const float* heavy, const float* acceptor, int32_t R, int64_t F, int64_t ld, void* out,
                            int32_t* exact_flag, jaxent_stream_t stream);
/* The same, storing N[r,f] - row_offsets[a][r] (row_offsets: device float[2][Rp], Rp = n_residues padded to 128, integer-valued;
 * typically the rounded row means).  Hard-cutoff counts minus an integer stay exact in bf16, and the column sums H_f of the
 * backward GEMM shrink from R x mean to the fluctuation scale, which is what their fp32 accumulation can carry through the
 * cancellation H_f - hbar at R = 500 (measured: gradient error 3e-4 -> below 2e-5).  Pass the same array to
 * jaxent_bg_plan_set_row_offsets: the tail adds row_offsets * S back to the forward sums and centres hbar. */
int jaxent_bg_pack_features_centred(const float* heavy, const float* acceptor, const float* row_offsets, int32_t n_residues, int64_t n_frames,
                                    int64_t ld, void* out, int32_t* exact_flag, jaxent_stream_t stream);
int jaxent_bg_gemm1(const void* features, const void* cc3, float* H, int32_t n_fb, int32_t n_rg, int32_t Bn, int32_t NS,
                    int32_t sm_count, jaxent_stream_t stream);
int jaxent_bg_gemm2(const void* features, const void* e3, float* partials, int32_t n_fb, int32_t n_rg, int32_t Bn, int32_t NS,
                    int32_t ksplit, jaxent_stream_t stream);

/* ---- batched optimise step (csrc/bg_step.cu): n_replicas runs of OptaxOptimizer._step_with_rates that share the
 * features, the data and the starting point and differ in (forward_model_weights, forward_model_scaling, learning_rate)
 * -- what batch_optimise (opt/batch.py:51-192) vmaps.  problem->packed must hold the bf16 GEMM layout
 * (jaxent_bg_pack_features) and problem->feature_format must be 2; n_replicas is a multiple of 16 (<= 256; pad by
 * repeating the last replica as the reference does, batch.py:25-29); n_pieces in 1..3.
 * fw / fs: host float[n_replicas][JAXENT_MAX_SLOTS]; lr: host float[n_replicas] or NULL (cfg->learning_rate).
 * jaxent_bg_state_init = OptaxOptimizer.initialise + the step-0 forward; jaxent_bg_step = one optimise step of every replica.
 * Buffers (jaxent_bg_plan_buffer): 0 logits float[F][Bn], 1 exp(z - lse) float[F][Bn] (softmax weights = column / S),
 * 2 JaxentLossRecord[256][Bn] ring (index step % 256), 3 control blocks (jaxent_bg_ctl_sizeof() bytes each; double S at
 * jaxent_bg_ctl_offset(0)), 4 log_Pf float[Bn][R], 5 uptake float[Bn][T*R]; 8 / 9 the two masked-gradient buffers float[F][Bn]
 * (state.gradients of the update applied last is the one selected by the uint32 at buffer 10). */
typedef struct JaxentBatchPlan JaxentBatchPlan;
size_t jaxent_bg_workspace_bytes(const JaxentBvProblem* problem, int32_t n_replicas, int32_t n_pieces, int32_t sm_count);
int jaxent_bg_plan_create(const JaxentBvProblem* problem, const JaxentOptConfig* cfg, int32_t n_replicas, int32_t n_pieces,
                          void* workspace, size_t workspace_bytes, int32_t sm_count, JaxentBatchPlan** plan);
void jaxent_bg_plan_destroy(JaxentBatchPlan* plan);
/* features packed by jaxent_bg_pack_features_centred: the same row_offsets (device float[2][Rp], kept alive by the caller); before state_init */
int jaxent_bg_plan_set_row_offsets(JaxentBatchPlan* plan, const float* row_offsets);
int jaxent_bg_state_init(JaxentBatchPlan* plan, const float* z0, float zmax, float bc, float bh, const float* fw_host,
                         const float* fs_host, const float* lr_host, jaxent_stream_t stream);
int jaxent_bg_step(JaxentBatchPlan* plan, jaxent_stream_t stream);
/* n steps in one call: between two of them the per-(frame, replica) MaxEnt kernel of step k runs on the plan's side stream
 * beside the backward tensor-core product of step k + 1 (event fork / join, no host synchronisation); the last step stays on
 * `stream`, so everything observable after the call is complete.  jaxent_bg_step(p, s) == jaxent_bg_steps(p, 1, s). */
int jaxent_bg_steps(JaxentBatchPlan* plan, int32_t n, jaxent_stream_t stream);
/* the four phases of jaxent_bg_step, for timing: 0 gemm1, 1 grad + decide + adam, 2 gemm2, 3 reduce + tails + MaxEnt terms */
int jaxent_bg_step_phase(JaxentBatchPlan* plan, int32_t phase, jaxent_stream_t stream);
int jaxent_bg_plan_buffer(const JaxentBatchPlan* plan, int32_t which, void** ptr, size_t* bytes);
/* on-device loop of every replica (the ConvergenceTracker of _optimise_pure, opt/run.py:141-232): call before
 * jaxent_bg_state_init.  convergence: n (<= 8) relative thresholds in descending order, NOT yet multiplied by the learning
 * rate (each replica uses convergence[i] * its own learning rate, opt/track.py:12-21).  ema_w: float[F][Bn]; snap_w, snap_ema:
 * float[n+1][F][Bn]; snaps: JaxentTrackSnapshot[n+1][Bn].  A replica whose loop has ended is frozen; the others continue. */
int jaxent_bg_track_init(JaxentBatchPlan* plan, const float* convergence, int32_t n, float ema_alpha, int32_t min_steps_per_threshold,
                         float tolerance, float* ema_w, float* snap_w, float* snap_ema, JaxentTrackSnapshot* snaps);
/* which loop the per-replica tracker follows; call between jaxent_bg_track_init and jaxent_bg_state_init.
 *   0 (default)  the Python loop _optimise (opt/run.py:235-373): ConvergenceTracker (opt/track.py:128-197), gradient-oscillation
 *                reset, one history snapshot per completed threshold (slots 0..n);
 *   1            _optimise_pure (opt/run.py:141-232) = lax.while_loop over OptaxOptimizer._pure_step (opt/optimiser.py:494-579),
 *                the loop batch_optimise vmaps (opt/batch.py:107-146): update_convergence / check_and_advance_threshold
 *                (opt/track.py:40-125) in fp32 -- the EMAs restart with every threshold, no oscillation reset, the threshold test
 *                uses the 1-based state.step, cond_fn stops on the loss of the previous iteration.  Every iteration is an entry of
 *                the reference's dense history; here the per-step losses stay in the record ring (256 deep: drain it at least
 *                that often) and snapshot slot 0 follows the entry get_best_state would pick (smallest val_losses[0], first
 *                occurrence; JaxentTrackSnapshot.loop_step = its index, snap_w slot 0 = its post-update frame weights). */
int jaxent_bg_track_mode(JaxentBatchPlan* plan, int32_t mode);
int jaxent_bg_track_status(JaxentBatchPlan* plan, JaxentTrackStatus* status, jaxent_stream_t stream);
int jaxent_bg_ctl_sizeof(void);
int jaxent_bg_ctl_offset(int32_t field); /* 0: double S, 1: float bc, 2: float bh, 3: int step, 4: float lse, 5: int steps_since_threshold_start,
                                           6: int best_step, 7: float best_val (pure-loop tracker) */

/* ---- plan handles and workspace identity, for foreign-function layers (csrc/xla_ffi.cc, the jax.ffi custom calls of
 * INTEGRATION.md; the reference has no FFI of its own).  A custom call cannot carry a typed pointer, and a raw
 * pointer attribute would outlive its plan silently: handles are (generation, slot) pairs that jaxent_handle_lookup
 * rejects once the plan has been destroyed.  The workspace a plan was built on is reported by *_plan_workspace so that a
 * handler can compare it with the buffer the runtime hands it; jaxent_bv_plan_rebind_workspace moves a plan to a new copy
 * of the same workspace contents (XLA copies an operand it cannot alias in place). */
#define JAXENT_HANDLE_BV_PLAN 1
#define JAXENT_HANDLE_BATCH_PLAN 2
int64_t jaxent_handle_register(void* plan, int32_t kind); /* > 0, or 0 on error; idempotent per plan */
void* jaxent_handle_lookup(int64_t handle, int32_t kind);  /* NULL (and jaxent_last_error) for stale / foreign handles */
void jaxent_handle_release_ptr(void* plan);                /* called by the *_plan_destroy functions */
int jaxent_bv_plan_workspace(const JaxentPlan* plan, void** ptr, size_t* bytes);
int jaxent_bv_plan_rebind_workspace(JaxentPlan* plan, void* workspace, size_t workspace_bytes);
int jaxent_bg_plan_workspace(const JaxentBatchPlan* plan, void** ptr, size_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* JAXENT_B200_H_ */
