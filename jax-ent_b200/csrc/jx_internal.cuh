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
  double hbar_data;            // sum_r c_r * log_Pf_r  (= sum_j w_j h_j, data part)
  float fw[JAXENT_MAX_SLOTS], fs[JAXENT_MAX_SLOTS];
  // ---- on-device convergence tracking (reference opt/run.py:272-348, opt/track.py:128-197)
  int trk_stop;                // 1: the loop has ended; pass / reduce / tail launches become no-ops
  int trk_reason;              // 1 all thresholds completed, 2 tolerance reached / non-finite loss
  int trk_stop_step;           // loop index of the step that ended the loop
  int trk_have_prev, trk_have_ema;
  int trk_steps_since, trk_idx, trk_nsnap;
  float trk_ema_bc, trk_ema_bh;
  int trk_pad[2];
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
  const float* cc;
  const float* ch;
  double* partials;
  const double* klsum;
  const Ctl* ctl;
  const unsigned* epoch;
  JaxentLossRecord* records;
  float b1, b2, om_b1, om_b2, eps, clip;
};

struct ReduceArgs {
  const Ctl* ctl;
  const double* partials;
  double* red;
  unsigned* epoch;
  int n_part;
  int part_n;
};

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
  const double* red;
  Ctl* ctl;
  const unsigned* epoch;
  unsigned* counter;
  float* w;
  float* kappa;
  float* cc;
  float* ch;
  float* logpf;
  float* uptake;
  float* avg;
  double* klpart;
  double* klsum;
  JaxentLossRecord* records;
  BlobSet blobs;
  double* cl_scratch;  // L2 exchange buffer of the clustered tail (tail_cluster_scratch_doubles())
  int p_max_tr, p_max_va, nnz_max_tr, nnz_max_va;
  long long frames_per_cta;  // frames handled by each per-frame CTA of the tail kernel
  float eps_kl;   // 1e-10 / F_total
};

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
int launch_reduce(const ReduceArgs& a, cudaStream_t s);
int launch_tail(const TailArgs& a, int ctas, size_t smem, cudaStream_t s);
int launch_finish_record(const Ctl* ctl, const unsigned* epoch, const double* klsum, JaxentLossRecord* rec, int n_slots,
                         cudaStream_t s);
size_t tail_smem_bytes(int R, int T, int p_tr, int p_va, int nnz_tr, int nnz_va);
int blob_words(int R, int T, int p_tr, int nnz_tr, int p_va, int nnz_va);
size_t tail_cluster_smem_bytes(int R, int T, int p_tr, int p_va, int nnz_tr, int nnz_va);
size_t tail_cluster_scratch_doubles(int R, int T, int p_tr);
int launch_tail_cluster(const TailArgs& a, int ctas, size_t smem, cudaStream_t s);
constexpr int TAIL_CLUSTER = 8;
int launch_build_blobs(const JaxentBvProblem& pb, const BlobSet& bs, cudaStream_t s);
int configure_kernels();

__device__ __forceinline__ void finish_record(const Ctl* c, const double* klsum, JaxentLossRecord* records, int n_slots) {
  JaxentLossRecord* r = records + c->rec_idx;
  if (c->kl_slot >= 0) {
    const float kl = (float)(klsum[1] + klsum[2]);
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
