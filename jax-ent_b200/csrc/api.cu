// C ABI of libjaxent_b200: plan / workspace management and the step entry points.
// See include/jaxent_b200.h for the contract and the reference symbols each call replaces.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "jx_internal.cuh"

namespace jx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Layout {
  size_t off_z[2][2], off_m[2][2], off_v[2][2], off_g[2];
  size_t off_w, off_kappa, off_kappa_lo, off_cc, off_ch, off_ck, off_logpf, off_uptake, off_avg, off_pfG;
  size_t off_cta_scal, off_fx, off_nmax, off_xchg, off_klpart, off_ctl, off_epoch, off_counter, off_records;
  size_t off_cl_scratch, off_pf_db;
  size_t off_blob[JAXENT_MAX_SLOTS];
  int blob_words[JAXENT_MAX_SLOTS];
  size_t total;
};

constexpr int MAX_TAIL_CTAS = 176;

static Layout make_layout(const JaxentBvProblem* pb) {
  const int R = pb->R, T = pb->uptake ? pb->T : 0;
  const long long F = pb->F;
  Layout L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes, 256);
    return r;
  };
  const size_t fb = (size_t)F * sizeof(float);
  for (int s = 0; s < 2; ++s)
    for (int b = 0; b < 2; ++b) {
      L.off_z[s][b] = take(fb);
      L.off_m[s][b] = take(fb);
      L.off_v[s][b] = take(fb);
    }
  L.off_g[0] = take(fb);
  L.off_g[1] = take(fb);
  L.off_w = take(fb);
  L.off_kappa = take(fb);
  L.off_kappa_lo = take(fb);
  L.off_cc = take((size_t)R * 4);
  L.off_ch = take((size_t)R * 4);
  L.off_ck = take((size_t)R * 4);
  L.off_logpf = take((size_t)R * 4);
  L.off_uptake = take((size_t)(T > 0 ? T : 1) * R * 4);
  L.off_avg = take((size_t)2 * R * 4);
  L.off_pfG = take((size_t)(T > 0 ? T : 1) * R * 4);
  const int part_n = pass_part_n(R, T, pb->uptake);
  L.off_cta_scal = take((size_t)MAX_PASS_CTAS * 4 * sizeof(double));
  L.off_fx = take((size_t)FX_REP * (part_n - 4) * sizeof(unsigned long long));
  L.off_nmax = take(sizeof(float));
  L.off_xchg = take(xchg_bytes(part_n, 1));  // single-rank exchange buffer (red / kl inbox of world = 1)
  L.off_klpart = take((size_t)MAX_TAIL_CTAS * KL_N * sizeof(double));
  L.off_ctl = take(2 * sizeof(Ctl));
  L.off_epoch = take(sizeof(unsigned));
  L.off_counter = take(sizeof(unsigned));
  L.off_records = take((size_t)JAXENT_RECORD_RING * sizeof(JaxentLossRecord));
  L.off_cl_scratch = take((64 + 4 * (size_t)MAX_PASS_CTAS) * sizeof(double));  // + per-CTA pass timestamps (JX_PROFILE builds)
  L.off_pf_db = take((size_t)PF_DB_MAX_CTAS * 2 * sizeof(double));  // per-frame mode: CTA partials of d loss / d (bc, bh)
  for (int i = 0; i < JAXENT_MAX_SLOTS; ++i) {
    L.off_blob[i] = 0;
    L.blob_words[i] = 0;
    if (i < pb->n_slots && pb->slots[i].kind <= JAXENT_LOSS_PF_L2) {
      const JaxentLossSlot& sl = pb->slots[i];
      L.blob_words[i] = blob_words(R, T, sl.train.n_pep, sl.train.nnz, sl.val.n_pep, sl.val.nnz);
      L.off_blob[i] = take((size_t)L.blob_words[i] * 4);
    }
  }
  L.total = o;
  return L;
}

__global__ void state_init_kernel(const float* __restrict__ z0, long long F, StateSets st, Ctl* ctl, Ctl init,
                                  unsigned* epoch, unsigned* counter, unsigned long long* xchg_local,
                                  long long xchg_words, JaxentLossRecord* records, unsigned long long* fx, long long fx_words,
                                  unsigned* nmax_bits, unsigned nmax_init) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long f = i0; f < F; f += stride) {
    st.z[0][0][f] = z0[f];
    st.m[0][0][f] = 0.f;
    st.v[0][0][f] = 0.f;
    st.g[0][f] = 0.f;
    st.g[1][f] = 0.f;
  }
  if (i0 == 0) {
    ctl[0] = init;          // epoch_id 0: the current block
    init.epoch_id = -1;
    ctl[1] = init;
    *epoch = 0u;
    *counter = 0u;
    *nmax_bits = nmax_init;
  }
  for (long long i = i0; i < fx_words; i += stride) fx[i] = 0ull;
  // the local exchange buffer: inboxes, flags and the error word (peers only write here after every rank's state_init)
  for (long long i = i0; i < xchg_words; i += stride) xchg_local[i] = 0ull;
  for (long long i = i0; i < (long long)(JAXENT_RECORD_RING * sizeof(JaxentLossRecord) / 4); i += stride)
    reinterpret_cast<int*>(records)[i] = 0;
}

// max |x| over the packed fp32 features as a bit pattern (|x| bits compare like unsigned integers; inf and NaN rank above
// every finite value, so a non-finite feature is seen): bounds the fixed-point row sums of the pass (fx_scale)
__global__ void absmax_kernel(const uint4* __restrict__ x, long long n16, unsigned* out_bits) {
  unsigned m = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = x[i];
    m = max(max(m, v.x & 0x7fffffffu), max(max(v.y & 0x7fffffffu, v.z & 0x7fffffffu), v.w & 0x7fffffffu));
  }
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out_bits, m);
}

__global__ void export_kernel(StateSets st, const Ctl* ctl, const unsigned* epoch, long long F, float* z, float* m,
                              float* v, float* g, float* scalars) {
  const unsigned ep = *epoch;
  const Ctl* c = ctl + (ep & 1u);
  const int set = ep & 1u, br = c->branch;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long f = i0; f < F; f += stride) {
    if (z) z[f] = st.z[set][br][f];
    if (m) m[f] = st.m[set][br][f];
    if (v) v[f] = st.v[set][br][f];
    if (g) g[f] = st.g[set][f];
  }
  if (i0 == 0 && scalars) {
    scalars[0] = (float)c->step;
    scalars[1] = c->lr;
    scalars[2] = c->model_lr;
    scalars[3] = (float)c->mask_idx;
    scalars[4] = c->bc;
    scalars[5] = c->bh;
    scalars[6] = (float)c->count_frame;
    scalars[7] = (float)c->count_model;
  }
}

__global__ void track_status_kernel(const Ctl* ctl, const unsigned* epoch, const JaxentLossRecord* records, JaxentTrackStatus* out) {
  const Ctl* c = ctl + (*epoch & 1u);
  out->stop = c->trk_stop;
  out->reason = c->trk_reason;
  out->stop_step = c->trk_stop_step;
  out->n_snapshots = c->trk_nsnap;
  out->steps_done = c->step;
  out->threshold_idx = c->trk_idx;
  out->last_loss = (float)c->trk_prev_loss;
  out->ema_loss_delta = (float)c->trk_ema;
}

}  // namespace jx

struct JaxentPlan {
  JaxentBvProblem prob;
  JaxentOptConfig cfg;
  jx::Layout lay;
  unsigned char* ws;
  jx::StateSets st;
  int pass_threads, rpt, pass_ctas, stages, tail_ctas, p_max_tr, p_max_va, nnz_max_tr, nnz_max_va;
  int sums_folded;  // jaxent_bv_reduce ran after the last pass: the tail reads JAXENT_BUF_RED instead of completing the sums itself
  jx::Xchg x;
  uint32_t tile_bytes, stage_stride, subtile_bytes;
  int sub_tiles;
  long long n_stage_units;
  size_t pass_smem, tail_smem;
  long long n_tiles;
  int initialised;
  jx::TrackConfig trk;
};

using namespace jx;

static int validate_problem(const JaxentBvProblem* p) {
  if (!p) { set_error("problem is NULL"); return JAXENT_E_INVALID; }
  if (p->R <= 0 || p->F <= 0 || p->F_total < p->F) {
    set_error("invalid shape: R=%d F=%lld F_total=%lld", p->R, (long long)p->F, (long long)p->F_total);
    return JAXENT_E_INVALID;
  }
  if (p->R > 2 * PASS_THREADS_MAX) {
    set_error("n_residues=%d exceeds the supported maximum %d", p->R, 2 * PASS_THREADS_MAX);
    return JAXENT_E_UNSUPPORTED;
  }
  if (!p->packed || (reinterpret_cast<uintptr_t>(p->packed) & 15u)) {
    set_error("packed features must be a 16-byte aligned device pointer");
    return JAXENT_E_INVALID;
  }
  if (p->n_slots <= 0 || p->n_slots > JAXENT_MAX_SLOTS) {
    set_error("n_slots=%d outside 1..%d", p->n_slots, JAXENT_MAX_SLOTS);
    return JAXENT_E_INVALID;
  }
  if (p->uptake && (p->T <= 0 || !p->k_ints || !p->timepoints)) {
    set_error("uptake mode needs T>0, k_ints and timepoints");
    return JAXENT_E_INVALID;
  }
  if (p->uptake != 0 && p->uptake != 1 && p->uptake != 2) { set_error("uptake must be 0 (log_Pf), 1 (uptake) or 2 (per-frame uptake)"); return JAXENT_E_INVALID; }
  if (p->uptake == 2 && (p->R > PASS_THREADS_MAX || p->T > 8 || p->feature_format != 0)) {
    set_error("per-frame uptake mode (average_first=False) needs n_residues <= %d, <= 8 time points and fp32 features", PASS_THREADS_MAX);
    return JAXENT_E_UNSUPPORTED;
  }
  int n_kl = 0;
  for (int i = 0; i < p->n_slots; ++i) {
    const JaxentLossSlot& s = p->slots[i];
    switch (s.kind) {
      case JAXENT_LOSS_UPTAKE_EYE_MSE:
      case JAXENT_LOSS_UPTAKE_MC_EYE_MSE:
      case JAXENT_LOSS_UPTAKE_SIGMA_MSE:
        if (!p->uptake) { set_error("slot %d: uptake loss needs BV_uptake_ForwardPass (uptake=1)", i); return JAXENT_E_UNSUPPORTED; }
        break;
      case JAXENT_LOSS_PF_L2:
        if (p->uptake) { set_error("slot %d: hdx_pf_l2_loss needs BV_ForwardPass (uptake=0)", i); return JAXENT_E_UNSUPPORTED; }
        break;
      case JAXENT_LOSS_MAXENT_CONVEX_KL:
        if (!p->prior || !(p->prior_norm > 0.0)) { set_error("slot %d: maxent_convexKL_loss needs prior and prior_norm", i); return JAXENT_E_INVALID; }
        if (++n_kl > 1) { set_error("at most one maxent_convexKL_loss slot is supported"); return JAXENT_E_UNSUPPORTED; }
        break;
      case JAXENT_LOSS_PARAMS_L1:
      case JAXENT_LOSS_PARAMS_L2:
        break;
      default:
        set_error("slot %d: unsupported loss kind %d (this path covers the losses of SURVEY.md section 8a)", i, s.kind);
        return JAXENT_E_UNSUPPORTED;
    }
    if (s.kind <= JAXENT_LOSS_PF_L2) {
      const JaxentPeptideMap* maps[2] = {&s.train, &s.val};
      for (const JaxentPeptideMap* m : maps) {
        if (m->n_pep < 0 || m->nnz < 0 || (m->n_pep > 0 && (!m->row_ptr || !m->col || !m->val || !m->t_ptr || !m->t_row || !m->t_val || !m->y))) {
          set_error("slot %d: incomplete peptide map", i);
          return JAXENT_E_INVALID;
        }
        if (s.kind == JAXENT_LOSS_UPTAKE_SIGMA_MSE && m->n_pep > 0 && !m->W) {
          set_error("slot %d: sigma MSE needs the precision matrix W", i);
          return JAXENT_E_INVALID;
        }
      }
      if (s.train.n_pep <= 0) { set_error("slot %d: empty training set", i); return JAXENT_E_INVALID; }
    }
  }
  return JAXENT_OK;
}

extern "C" const char* jaxent_last_error(void) { return g_err; }
extern "C" int jaxent_version(void) { return 100; }
extern "C" int jaxent_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(JaxentBvProblem);
    case 1: return (int)sizeof(JaxentOptConfig);
    case 2: return (int)sizeof(JaxentLossRecord);
    case 3: return (int)sizeof(JaxentLossSlot);
    case 4: return (int)sizeof(JaxentPeptideMap);
    case 5: return (int)sizeof(JaxentTrackSnapshot);
    case 6: return (int)sizeof(JaxentTrackStatus);
    default: return -1;
  }
}

extern "C" size_t jaxent_bv_workspace_bytes(const JaxentBvProblem* p) {
  if (!p || p->R <= 0 || p->F <= 0) return 0;
  if (p->n_slots < 0 || p->n_slots > JAXENT_MAX_SLOTS) return 0;
  return make_layout(p).total;
}

extern "C" int jaxent_bv_plan_create(const JaxentBvProblem* problem, const JaxentOptConfig* cfg, void* workspace,
                                     size_t workspace_bytes, int32_t device_sm_count, JaxentPlan** out) {
  if (!out) { set_error("plan output pointer is NULL"); return JAXENT_E_INVALID; }
  *out = nullptr;
  int rc = validate_problem(problem);
  if (rc) return rc;
  if (!cfg) { set_error("optimiser config is NULL"); return JAXENT_E_INVALID; }
  if (!(cfg->plateau_denominator > 0.f) || !(cfg->b1 >= 0.0 && cfg->b1 < 1.0) || !(cfg->b2 >= 0.0 && cfg->b2 < 1.0)) {
    set_error("invalid optimiser config");
    return JAXENT_E_INVALID;
  }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
    set_error("workspace must be a 256-byte aligned device pointer");
    return JAXENT_E_INVALID;
  }
  const int R = problem->R, T = problem->uptake ? problem->T : 0;
  if (problem->uptake == 2 && cfg->optimise_model_parameters && problem->F_total != problem->F) {
    set_error("per-frame uptake mode (average_first=False) with optimised BV parameters runs on one GPU: their gradient is a frame "
              "sum of its own that the frame-sharded exchange does not carry (set optimise_model_parameters = 0 or do not shard)");
    return JAXENT_E_UNSUPPORTED;
  }
  Layout L = make_layout(problem);
  if (workspace_bytes < L.total) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, L.total);
    return JAXENT_E_WORKSPACE;
  }
  if (device_sm_count <= 0) device_sm_count = 148;
  if (device_sm_count > 160) device_sm_count = 160;

  JaxentPlan* p = new (std::nothrow) JaxentPlan();
  if (!p) { set_error("out of host memory"); return JAXENT_E_INVALID; }
  p->prob = *problem;
  p->cfg = *cfg;
  p->lay = L;
  p->ws = static_cast<unsigned char*>(workspace);
  for (int s = 0; s < 2; ++s)
    for (int b = 0; b < 2; ++b) {
      p->st.z[s][b] = reinterpret_cast<float*>(p->ws + L.off_z[s][b]);
      p->st.m[s][b] = reinterpret_cast<float*>(p->ws + L.off_m[s][b]);
      p->st.v[s][b] = reinterpret_cast<float*>(p->ws + L.off_v[s][b]);
    }
  p->st.g[0] = reinterpret_cast<float*>(p->ws + L.off_g[0]);
  p->st.g[1] = reinterpret_cast<float*>(p->ws + L.off_g[1]);

  // pass geometry: one residue row per thread, RPT rows when R > 512
  p->rpt = (R <= PASS_THREADS_MAX) ? 1 : 2;
  const int rows_per_thread_threads = (R + p->rpt - 1) / p->rpt;
  p->pass_threads = ((rows_per_thread_threads + 31) / 32) * 32;
  if (p->pass_threads < 32) p->pass_threads = 32;
  if (problem->feature_format != 0 && problem->feature_format != 1) { delete p; set_error("unknown feature_format %d", problem->feature_format); return JAXENT_E_INVALID; }
  p->sub_tiles = problem->feature_format == 1 ? U8_SUB_TILES : 1;
  p->subtile_bytes = (uint32_t)(FT * R * 2 * (problem->feature_format == 1 ? 1 : sizeof(float)));
  p->tile_bytes = p->subtile_bytes * p->sub_tiles;
  p->stage_stride = (uint32_t)align_up(p->tile_bytes, 128);
  const size_t smem_budget = ((p->rpt == 1 && problem->uptake != 2) ? 108 : 216) * 1024;
  // per-frame mode adds per-thread slots {k tau, G}: [4 pairs of time points][512] float4, and the row partials [2 quads][512] float4
  const size_t fixed = MAX_STAGES * 8 + 2 * 16 * FT * 4 + 2 * 2 * FT * 4 + 128 + (problem->uptake == 2 ? (4 + 2) * PASS_THREADS_MAX * 16 : 0);
  int stages = (int)((smem_budget - fixed) / p->stage_stride);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) { delete p; set_error("tile of %u bytes does not fit a 2-stage ring", p->tile_bytes); return JAXENT_E_UNSUPPORTED; }
  p->stages = stages;
  p->pass_smem = (size_t)stages * p->stage_stride + fixed;
  p->n_tiles = (problem->F + FT - 1) / FT;
  p->n_stage_units = (p->n_tiles + p->sub_tiles - 1) / p->sub_tiles;
  const int per_sm = (p->rpt == 1 && problem->uptake != 2) ? 2 : 1;
  long long ctas = (long long)device_sm_count * per_sm;
  if (ctas > p->n_stage_units) ctas = p->n_stage_units;
  if (ctas > MAX_PASS_CTAS) ctas = MAX_PASS_CTAS;
  p->pass_ctas = (int)ctas;

  int p_tr = 1, p_va = 1, nnz_tr = 1, nnz_va = 1;
  for (int i = 0; i < problem->n_slots; ++i)
    if (problem->slots[i].kind <= JAXENT_LOSS_PF_L2) {
      const JaxentLossSlot& sl = problem->slots[i];
      if (sl.train.n_pep > p_tr) p_tr = sl.train.n_pep;
      if (sl.val.n_pep > p_va) p_va = sl.val.n_pep;
      if (sl.train.nnz > nnz_tr) nnz_tr = sl.train.nnz;
      if (sl.val.nnz > nnz_va) nnz_va = sl.val.nnz;
    }
  p->p_max_tr = p_tr;
  p->p_max_va = p_va;
  p->nnz_max_tr = nnz_tr;
  p->nnz_max_va = nnz_va;
  p->tail_smem = tail_smem_bytes(R, T, p_tr, p_va, nnz_tr, nnz_va);
  if (p->tail_smem > 200 * 1024) {  // + 22 KB of static shared memory in bv_tail_kernel: 227 KB per CTA
    const size_t need = p->tail_smem;
    delete p;
    set_error("R*T / peptide maps too large for the tail kernel (%zu bytes of shared memory)", need);
    return JAXENT_E_UNSUPPORTED;
  }
  // CTA 0 runs the R x T tail; the per-frame MaxEnt work is spread over the others (one CTA does both when F is small)
  long long fctas = (problem->F + 4095) / 4096;
  if (fctas > device_sm_count - 1) fctas = device_sm_count - 1;
  if (fctas < 1) fctas = 1;
  p->tail_ctas = (int)fctas + 1;
  memset(&p->x, 0, sizeof(p->x));
  p->x.world = 1;
  p->x.rank = 0;
  p->x.part_n = pass_part_n(R, T, problem->uptake);
  p->x.base[0] = p->ws + L.off_xchg;
  p->x.local = p->x.base[0];
  p->initialised = 0;
  memset(&p->trk, 0, sizeof(p->trk));
  *out = p;
  return JAXENT_OK;
}

extern "C" void jaxent_bv_plan_destroy(JaxentPlan* plan) {
  jaxent_handle_release_ptr(plan);
  delete plan;
}

// ---- plan handles for foreign-function layers (csrc/xla_ffi.cc): a custom call carries a small integer, not a raw
// pointer, so a handle that outlives its plan is detected instead of dereferenced.  handle = generation << 16 | slot.
namespace {
constexpr int HANDLE_SLOTS = 1024;
struct HandleSlot { void* ptr; int kind; unsigned gen; };
HandleSlot g_slots[HANDLE_SLOTS];
std::mutex g_slots_mu;
}  // namespace

extern "C" int64_t jaxent_handle_register(void* plan, int32_t kind) {
  if (!plan || (kind != JAXENT_HANDLE_BV_PLAN && kind != JAXENT_HANDLE_BATCH_PLAN)) { set_error("handle_register: invalid argument"); return 0; }
  std::lock_guard<std::mutex> lk(g_slots_mu);
  for (int i = 1; i < HANDLE_SLOTS; ++i)
    if (g_slots[i].ptr == plan) return ((int64_t)g_slots[i].gen << 16) | i;  // idempotent
  for (int i = 1; i < HANDLE_SLOTS; ++i)
    if (!g_slots[i].ptr) {
      g_slots[i].ptr = plan;
      g_slots[i].kind = kind;
      g_slots[i].gen += 1;
      return ((int64_t)g_slots[i].gen << 16) | i;
    }
  set_error("handle_register: more than %d live plans", HANDLE_SLOTS - 1);
  return 0;
}

extern "C" void* jaxent_handle_lookup(int64_t handle, int32_t kind) {
  const int i = (int)(handle & 0xffff);
  const unsigned gen = (unsigned)(handle >> 16);
  std::lock_guard<std::mutex> lk(g_slots_mu);
  if (handle <= 0 || i <= 0 || i >= HANDLE_SLOTS || !g_slots[i].ptr || g_slots[i].gen != gen || g_slots[i].kind != kind) {
    set_error("stale or unknown plan handle %lld", (long long)handle);
    return nullptr;
  }
  return g_slots[i].ptr;
}

extern "C" void jaxent_handle_release_ptr(void* plan) {
  if (!plan) return;
  std::lock_guard<std::mutex> lk(g_slots_mu);
  for (int i = 1; i < HANDLE_SLOTS; ++i)
    if (g_slots[i].ptr == plan) g_slots[i].ptr = nullptr;  // the generation stays: an old handle no longer matches
}

extern "C" int jaxent_bv_plan_workspace(const JaxentPlan* p, void** ptr, size_t* bytes) {
  if (!p || !ptr || !bytes) { set_error("plan_workspace: NULL argument"); return JAXENT_E_INVALID; }
  *ptr = p->ws;
  *bytes = p->lay.total;
  return JAXENT_OK;
}

// The caller's allocator moved the workspace (XLA copies a buffer it could not alias in place): same contents at a new
// address.  Every pointer the plan derived from the old base is re-derived; external buffers (features, prior, maps,
// peer-mapped exchange buffers, tracker buffers) are untouched.
extern "C" int jaxent_bv_plan_rebind_workspace(JaxentPlan* p, void* workspace, size_t workspace_bytes) {
  if (!p) { set_error("rebind_workspace: plan is NULL"); return JAXENT_E_INVALID; }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) { set_error("workspace must be a 256-byte aligned device pointer"); return JAXENT_E_INVALID; }
  if (workspace_bytes < p->lay.total) { set_error("workspace too small: %zu < %zu", workspace_bytes, p->lay.total); return JAXENT_E_WORKSPACE; }
  const Layout& L = p->lay;
  p->ws = static_cast<unsigned char*>(workspace);
  for (int s = 0; s < 2; ++s)
    for (int b = 0; b < 2; ++b) {
      p->st.z[s][b] = reinterpret_cast<float*>(p->ws + L.off_z[s][b]);
      p->st.m[s][b] = reinterpret_cast<float*>(p->ws + L.off_m[s][b]);
      p->st.v[s][b] = reinterpret_cast<float*>(p->ws + L.off_v[s][b]);
    }
  p->st.g[0] = reinterpret_cast<float*>(p->ws + L.off_g[0]);
  p->st.g[1] = reinterpret_cast<float*>(p->ws + L.off_g[1]);
  if (p->x.world == 1) {  // the single-rank exchange buffer lives in the workspace
    p->x.base[0] = p->ws + L.off_xchg;
    p->x.local = p->x.base[0];
  }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_plan_buffer(const JaxentPlan* p, int32_t which, void** ptr, size_t* bytes) {
  if (!p || !ptr || !bytes) { set_error("plan_buffer: NULL argument"); return JAXENT_E_INVALID; }
  const int R = p->prob.R, T = p->prob.uptake ? p->prob.T : 0;
  switch (which) {
    case JAXENT_BUF_RED:  // plain doubles; meaningful when the exchange is not fused (one rank, or NCCL between the phases)
      *ptr = p->x.local + xchg_plain_red_off(p->x); *bytes = (size_t)p->x.part_n * 8; break;
    case JAXENT_BUF_KLSUM: *ptr = p->x.local + xchg_plain_kl_off(p->x); *bytes = KL_N * 8; break;
    case JAXENT_BUF_XCHG_ERR: *ptr = p->x.local + xchg_err_off(p->x); *bytes = 8; break;
    case JAXENT_BUF_WEIGHTS: *ptr = p->ws + p->lay.off_w; *bytes = (size_t)p->prob.F * 4; break;
    case JAXENT_BUF_RECORDS: *ptr = p->ws + p->lay.off_records; *bytes = (size_t)JAXENT_RECORD_RING * sizeof(JaxentLossRecord); break;
    case JAXENT_BUF_LOGPF: *ptr = p->ws + p->lay.off_logpf; *bytes = (size_t)R * 4; break;
    case JAXENT_BUF_UPTAKE: *ptr = p->ws + p->lay.off_uptake; *bytes = (size_t)T * R * 4; break;
    case JAXENT_BUF_AVG: *ptr = p->ws + p->lay.off_avg; *bytes = (size_t)2 * R * 4; break;
    case 8: *ptr = p->ws + p->lay.off_cl_scratch; *bytes = (64 + 4 * (size_t)MAX_PASS_CTAS) * 8; break;  // debug timestamps (JX_PROFILE builds)
    default: set_error("plan_buffer: unknown buffer id %d", which); return JAXENT_E_INVALID;
  }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_state_init(JaxentPlan* p, const float* z0, float zmax, float bc, float bh, const float* fw_host,
                                    const float* fs_host, jaxent_stream_t stream) {
  if (!p || !z0 || !fw_host || !fs_host) { set_error("state_init: NULL argument"); return JAXENT_E_INVALID; }
  int rc = configure_kernels();
  if (rc) return rc;
  Ctl c;
  memset(&c, 0, sizeof(c));
  c.lr = p->cfg.initial_learning_rate;
  c.model_lr = p->cfg.initial_learning_rate * p->cfg.model_parameters_lr_scale;
  c.bc = bc;
  c.bh = bh;
  c.lse = zmax;
  c.kl_slot = -1;
  c.lrA = c.lrB = c.lr;
  c.mlrA = c.mlrB = c.model_lr;
  c.mask_frame = 1.f;
  c.bc1 = c.bc2 = 1.f;
  for (int i = 0; i < p->prob.n_slots; ++i) {
    c.fw[i] = fw_host[i];
    c.fs[i] = fs_host[i];
    if (p->prob.slots[i].kind == JAXENT_LOSS_MAXENT_CONVEX_KL) {
      c.kl_slot = i;
      c.kl_scale = fw_host[i] * fs_host[i];
    }
  }
  Ctl* ctl = reinterpret_cast<Ctl*>(p->ws + p->lay.off_ctl);
  state_init_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(
      z0, p->prob.F, p->st, ctl, c, reinterpret_cast<unsigned*>(p->ws + p->lay.off_epoch),
      reinterpret_cast<unsigned*>(p->ws + p->lay.off_counter),
      reinterpret_cast<unsigned long long*>(p->x.local), (long long)(xchg_bytes(p->x.part_n, p->x.world) / 8),
      reinterpret_cast<JaxentLossRecord*>(p->ws + p->lay.off_records), reinterpret_cast<unsigned long long*>(p->ws + p->lay.off_fx),
      (long long)FX_REP * (p->x.part_n - 4), reinterpret_cast<unsigned*>(p->ws + p->lay.off_nmax),
      p->prob.feature_format == 1 ? 0x437f0000u /* 255.0f: u8 features */ : 0u);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("state_init launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  if (p->prob.feature_format == 0) {
    absmax_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(p->prob.packed),
                                                         (long long)(jaxent_bv_packed_bytes(p->prob.R, p->prob.F) / 16),
                                                         reinterpret_cast<unsigned*>(p->ws + p->lay.off_nmax));
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("absmax launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  }
  BlobSet bs;
  for (int i = 0; i < JAXENT_MAX_SLOTS; ++i) {
    bs.words[i] = p->lay.blob_words[i];
    bs.blob[i] = bs.words[i] ? reinterpret_cast<int*>(p->ws + p->lay.off_blob[i]) : nullptr;
  }
  rc = launch_build_blobs(p->prob, bs, (cudaStream_t)stream);
  if (rc) return rc;
  p->initialised = 1;
  return JAXENT_OK;
}

static int check_ready(const JaxentPlan* p, const char* what) {
  if (!p) { set_error("%s: plan is NULL", what); return JAXENT_E_INVALID; }
  if (!p->initialised) { set_error("%s: jaxent_bv_state_init has not been called", what); return JAXENT_E_INVALID; }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_pass(JaxentPlan* p, int32_t mode, jaxent_stream_t stream) {
  int rc = check_ready(p, "pass");
  if (rc) return rc;
  if (mode != 0 && mode != 1) { set_error("pass: mode must be 0 or 1"); return JAXENT_E_INVALID; }
  PassArgs a;
  memset(&a, 0, sizeof(a));
  a.packed = p->prob.packed;
  a.R = p->prob.R;
  a.mode = mode;
  a.stages = p->stages;
  a.part_n = p->x.part_n;
  a.format = p->prob.feature_format;
  if (p->prob.uptake == 2) {
    a.pf_T = p->prob.T;
    a.pf_G = reinterpret_cast<const float*>(p->ws + p->lay.off_pfG);
    a.pf_k = p->prob.k_ints;
    a.pf_tp = p->prob.timepoints;
  }
  a.sub_tiles = p->sub_tiles;
  a.subtile_bytes = p->subtile_bytes;
  a.n_stage_units = p->n_stage_units;
  a.tile_bytes = p->tile_bytes;
  a.stage_stride = p->stage_stride;
  a.n_tiles = p->n_tiles;
  a.F = p->prob.F;
  a.st = p->st;
  a.w = reinterpret_cast<const float*>(p->ws + p->lay.off_w);
  a.kappa = reinterpret_cast<const float*>(p->ws + p->lay.off_kappa);
  a.kappa_lo = reinterpret_cast<const float*>(p->ws + p->lay.off_kappa_lo);
  a.cc = reinterpret_cast<const float*>(p->ws + p->lay.off_cc);
  a.ch = reinterpret_cast<const float*>(p->ws + p->lay.off_ch);
  a.ck = reinterpret_cast<const float*>(p->ws + p->lay.off_ck);
  a.fx = reinterpret_cast<unsigned long long*>(p->ws + p->lay.off_fx);
  a.cta_scal = reinterpret_cast<double*>(p->ws + p->lay.off_cta_scal);
  a.nmax = reinterpret_cast<const float*>(p->ws + p->lay.off_nmax);
  a.x = p->x;
  a.ctl = reinterpret_cast<const Ctl*>(p->ws + p->lay.off_ctl);
  a.epoch = reinterpret_cast<unsigned*>(p->ws + p->lay.off_epoch);
  a.records = reinterpret_cast<JaxentLossRecord*>(p->ws + p->lay.off_records);
  a.b1 = (float)p->cfg.b1;
  a.b2 = (float)p->cfg.b2;
  a.om_b1 = (float)(1.0 - p->cfg.b1);
  a.om_b2 = (float)(1.0 - p->cfg.b2);
  a.eps = (float)p->cfg.eps;
  a.clip = p->cfg.clip_value;
  a.dbg = reinterpret_cast<double*>(p->ws + p->lay.off_cl_scratch);
  p->sums_folded = 0;
  if (p->prob.uptake == 2 && p->cfg.optimise_model_parameters) {
    // per-frame mode with optimised BV parameters (reference models/core.py:298-305 places no restriction): their gradient is a
    // frame sum of its own (pf_dbeta_kernel) that must be complete before the update this pass applies; the pass then evaluates
    // its forward half at the updated parameters of both plateau branches
    a.pf_db = 1;
    a.cfg = p->cfg;
    if (mode == 1) {
      rc = launch_pf_dbeta(a, reinterpret_cast<double*>(p->ws + p->lay.off_pf_db), reinterpret_cast<Ctl*>(p->ws + p->lay.off_ctl), a.records,
                           (cudaStream_t)stream);
      if (rc) return rc;
    }
  }
  return launch_pass(a, p->pass_threads + 32, p->rpt, p->pass_ctas, p->pass_smem, (cudaStream_t)stream);
}

// The pass leaves raw sums (fixed-point accumulators + per-CTA scalars); the tail kernel normally completes them itself.
// A caller that all-reduces JAXENT_BUF_RED between the phases (NCCL instead of the fused peer-memory exchange) calls
// jaxent_bv_reduce after the pass: it writes the completed sums to JAXENT_BUF_RED and the tail then reads them from there.
static void fill_tail_args(JaxentPlan* p, TailArgs& a);
extern "C" int jaxent_bv_reduce(JaxentPlan* p, jaxent_stream_t stream) {
  int rc = check_ready(p, "reduce");
  if (rc) return rc;
  if (p->x.world != 1) { set_error("reduce: the fused exchange completes the sums in the tail kernel"); return JAXENT_E_INVALID; }
  TailArgs a;
  fill_tail_args(p, a);
  p->sums_folded = 1;
  return launch_reduce_fold(a, (cudaStream_t)stream);
}

static void fill_tail_args(JaxentPlan* p, TailArgs& a) {
  memset(&a, 0, sizeof(a));
  a.prob = p->prob;
  a.trk = p->trk;
  a.cfg = p->cfg;
  a.st = p->st;
  a.x = p->x;
  a.ctl = reinterpret_cast<Ctl*>(p->ws + p->lay.off_ctl);
  a.counter = reinterpret_cast<unsigned*>(p->ws + p->lay.off_counter);
  a.epoch = reinterpret_cast<unsigned*>(p->ws + p->lay.off_epoch);
  a.fx = reinterpret_cast<unsigned long long*>(p->ws + p->lay.off_fx);
  a.cta_scal = reinterpret_cast<const double*>(p->ws + p->lay.off_cta_scal);
  a.nmax = reinterpret_cast<const float*>(p->ws + p->lay.off_nmax);
  a.pass_ctas = p->pass_ctas;
  a.fold_here = p->sums_folded ? 0 : 1;
  a.w = reinterpret_cast<float*>(p->ws + p->lay.off_w);
  a.kappa = reinterpret_cast<float*>(p->ws + p->lay.off_kappa);
  a.kappa_lo = reinterpret_cast<float*>(p->ws + p->lay.off_kappa_lo);
  a.cc = reinterpret_cast<float*>(p->ws + p->lay.off_cc);
  a.ch = reinterpret_cast<float*>(p->ws + p->lay.off_ch);
  a.ck = reinterpret_cast<float*>(p->ws + p->lay.off_ck);
  a.logpf = reinterpret_cast<float*>(p->ws + p->lay.off_logpf);
  a.uptake = reinterpret_cast<float*>(p->ws + p->lay.off_uptake);
  a.avg = reinterpret_cast<float*>(p->ws + p->lay.off_avg);
  a.pf_G = reinterpret_cast<float*>(p->ws + p->lay.off_pfG);
  a.part_n = p->x.part_n;
  a.klpart = reinterpret_cast<double*>(p->ws + p->lay.off_klpart);
  a.records = reinterpret_cast<JaxentLossRecord*>(p->ws + p->lay.off_records);
  for (int i = 0; i < JAXENT_MAX_SLOTS; ++i) {
    a.blobs.words[i] = p->lay.blob_words[i];
    a.blobs.blob[i] = a.blobs.words[i] ? reinterpret_cast<int*>(p->ws + p->lay.off_blob[i]) : nullptr;
  }
  {
    const long long fc = p->tail_ctas - 1;
    a.frames_per_cta = (p->prob.F + fc - 1) / fc;
  }
  a.cl_scratch = reinterpret_cast<double*>(p->ws + p->lay.off_cl_scratch);
  a.p_max_tr = p->p_max_tr;
  a.p_max_va = p->p_max_va;
  a.nnz_max_tr = p->nnz_max_tr;
  a.nnz_max_va = p->nnz_max_va;
  a.eps_kl = (float)(1e-10 / (double)p->prob.F_total);
}

extern "C" int jaxent_bv_tail(JaxentPlan* p, jaxent_stream_t stream) {
  int rc = check_ready(p, "tail");
  if (rc) return rc;
  TailArgs a;
  fill_tail_args(p, a);
  return launch_tail(a, p->tail_ctas, p->tail_smem, (cudaStream_t)stream);
}

extern "C" int jaxent_bv_finish_record(JaxentPlan* p, jaxent_stream_t stream) {
  int rc = check_ready(p, "finish_record");
  if (rc) return rc;
  return launch_finish_record(reinterpret_cast<const Ctl*>(p->ws + p->lay.off_ctl),
                              reinterpret_cast<const unsigned*>(p->ws + p->lay.off_epoch),
                              p->x, reinterpret_cast<JaxentLossRecord*>(p->ws + p->lay.off_records), p->prob.n_slots,
                              (cudaStream_t)stream);
}

extern "C" int jaxent_bv_forward_init(JaxentPlan* p, jaxent_stream_t stream) {
  int rc = jaxent_bv_pass(p, 0, stream);
  if (rc) return rc;
  return jaxent_bv_tail(p, stream);
}

extern "C" int jaxent_bv_step(JaxentPlan* p, jaxent_stream_t stream) {
  int rc = jaxent_bv_pass(p, 1, stream);
  if (rc) return rc;
  return jaxent_bv_tail(p, stream);
}

// ---- NVLink peer exchange (set-up time only)
extern "C" size_t jaxent_bv_exchange_bytes(int32_t R, int32_t world) {
  if (R <= 0 || world < 1 || world > JAXENT_MAX_RANKS) return 0;
  return xchg_bytes(4 * R + 4, world);
}

extern "C" size_t jaxent_bv_exchange_bytes_for(const JaxentBvProblem* pb, int32_t world) {
  if (!pb || pb->R <= 0 || world < 1 || world > JAXENT_MAX_RANKS) return 0;
  return xchg_bytes(pass_part_n(pb->R, pb->uptake ? pb->T : 0, pb->uptake), world);
}

extern "C" int jaxent_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle) {
  if (!dev_ptr || !handle || bytes == 0) { set_error("peer_alloc: invalid argument"); return JAXENT_E_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == JAXENT_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    set_error("peer_alloc: %s", cudaGetErrorString(e));
    return JAXENT_E_CUDA;
  }
  memcpy(handle, &h, sizeof(h));
  *dev_ptr = p;
  return JAXENT_OK;
}

extern "C" int jaxent_peer_open(const unsigned char* handle, void** dev_ptr) {
  if (!dev_ptr || !handle) { set_error("peer_open: invalid argument"); return JAXENT_E_INVALID; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("peer_open: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  *dev_ptr = p;
  return JAXENT_OK;
}

extern "C" int jaxent_peer_close(void* dev_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) { set_error("peer_close: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" int jaxent_peer_free(void* dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) { set_error("peer_free: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_plan_set_exchange(JaxentPlan* p, int32_t rank, int32_t world, void* const* peers) {
  if (!p || !peers || world < 1 || world > JAXENT_MAX_RANKS || rank < 0 || rank >= world) {
    set_error("set_exchange: need 1..%d ranks and one mapped exchange buffer per rank", JAXENT_MAX_RANKS);
    return JAXENT_E_INVALID;
  }
  for (int r = 0; r < world; ++r)
    if (!peers[r] || (reinterpret_cast<uintptr_t>(peers[r]) & 15u)) { set_error("set_exchange: peer %d buffer is NULL or misaligned", r); return JAXENT_E_INVALID; }
  memset(&p->x, 0, sizeof(p->x));
  p->x.world = world;
  p->x.rank = rank;
  p->x.part_n = pass_part_n(p->prob.R, p->prob.uptake ? p->prob.T : 0, p->prob.uptake);
  for (int r = 0; r < world; ++r) p->x.base[r] = static_cast<unsigned char*>(peers[r]);
  p->x.local = p->x.base[rank];
  p->initialised = 0;  // state_init clears the local flags; it must follow
  return JAXENT_OK;
}

extern "C" int jaxent_bv_exchange_error(JaxentPlan* p, uint64_t* out, jaxent_stream_t stream) {
  if (!p || !out) { set_error("exchange_error: NULL argument"); return JAXENT_E_INVALID; }
  cudaError_t e = cudaMemcpyAsync(out, p->x.local + xchg_err_off(p->x), 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("exchange_error: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

// ---- per-frame forward pass (Simulation.predict, reference models/core.py:181-208): no frame average
namespace jx {
// one thread per (quad of 4 frames, residue): reads the packed tile layout, writes row-major (R, F) / (T, R, F)
__global__ void predict_frames_kernel(const float* __restrict__ packed, int R, long long F, int T, const float* __restrict__ k_ints,
                                      const float* __restrict__ tp, float bc, float bh, float* __restrict__ out) {
  const long long n_quads = (F + 3) / 4;
  const long long n = n_quads * R;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < n; it += (long long)gridDim.x * blockDim.x) {
    const long long q = it / R;
    const int r = (int)(it - q * R);
    const float4 c = reinterpret_cast<const float4*>(packed)[(q * 2 + 0) * R + r];
    const float4 h = reinterpret_cast<const float4*>(packed)[(q * 2 + 1) * R + r];
    const float nc[4] = {c.x, c.y, c.z, c.w}, nh[4] = {h.x, h.y, h.z, h.w};
    for (int j = 0; j < 4; ++j) {
      const long long f = 4 * q + j;
      if (f >= F) break;
      const float lp = bc * nc[j] + bh * nh[j];  // BV_ForwardPass, models/HDX/forward.py:30-35
      if (T == 0) {
        out[(long long)r * F + f] = lp;
      } else {
        const float pf = expf(lp);  // BV_uptake_ForwardPass, models/HDX/forward.py:57-74
        for (int t = 0; t < T; ++t) out[((long long)t * R + r) * F + f] = 1.0f - expf(-k_ints[r] * tp[t] / pf);
      }
    }
  }
}
}  // namespace jx

extern "C" int jaxent_bv_predict_frames(const float* packed, int32_t R, int64_t F, int32_t T, const float* k_ints, const float* timepoints,
                                        float bc, float bh, float* out, jaxent_stream_t stream) {
  if (!packed || !out || R <= 0 || F <= 0 || T < 0 || (T > 0 && (!k_ints || !timepoints))) { set_error("predict_frames: invalid argument"); return JAXENT_E_INVALID; }
  if (reinterpret_cast<uintptr_t>(packed) & 15u) { set_error("predict_frames: packed features must be 16-byte aligned"); return JAXENT_E_INVALID; }
  predict_frames_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(packed, R, F, T, k_ints, timepoints, bc, bh, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("predict_frames launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_track_init(JaxentPlan* p, const float* thresholds, int32_t n, float ema_alpha, int32_t min_steps,
                                    float tolerance, float* ema_w, float* snap_w, float* snap_ema, JaxentTrackSnapshot* snaps) {
  if (!p) { set_error("track_init: plan is NULL"); return JAXENT_E_INVALID; }
  if (!thresholds || n < 1 || n > TRK_MAX_THR || !ema_w || !snap_w || !snap_ema || !snaps) {
    set_error("track_init: need 1..%d thresholds and the EMA / snapshot buffers", TRK_MAX_THR);
    return JAXENT_E_INVALID;
  }
  if (!(ema_alpha >= 0.f && ema_alpha <= 1.f) || min_steps < 0) { set_error("track_init: invalid ema_alpha / min_steps"); return JAXENT_E_INVALID; }
  for (int i = 1; i < n; ++i)
    if (thresholds[i] > thresholds[i - 1]) { set_error("track_init: thresholds must be in descending order"); return JAXENT_E_INVALID; }
  TrackConfig t;
  memset(&t, 0, sizeof(t));
  t.on = 1;
  t.n_thr = n;
  t.min_steps = min_steps;
  t.initial_steps = p->cfg.initial_steps;
  for (int i = 0; i < n; ++i) t.thr[i] = thresholds[i];
  t.ema_alpha = ema_alpha;
  t.tolerance = tolerance;
  t.ema_w = ema_w;
  t.snap_w = snap_w;
  t.snap_ema = snap_ema;
  t.snaps = snaps;
  p->trk = t;
  p->initialised = 0;  // the tracker state lives in the control block: state_init must follow
  return JAXENT_OK;
}

extern "C" int jaxent_bv_track_status(JaxentPlan* p, JaxentTrackStatus* status, jaxent_stream_t stream) {
  int rc = check_ready(p, "track_status");
  if (rc) return rc;
  if (!status) { set_error("track_status: NULL output"); return JAXENT_E_INVALID; }
  track_status_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<const Ctl*>(p->ws + p->lay.off_ctl),
                                                         reinterpret_cast<const unsigned*>(p->ws + p->lay.off_epoch),
                                                         reinterpret_cast<const JaxentLossRecord*>(p->ws + p->lay.off_records), status);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("track_status launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" int jaxent_bv_export_state(JaxentPlan* p, float* z, float* m, float* v, float* g, float* scalars,
                                      jaxent_stream_t stream) {
  int rc = check_ready(p, "export_state");
  if (rc) return rc;
  export_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(p->st, reinterpret_cast<const Ctl*>(p->ws + p->lay.off_ctl),
                                                       reinterpret_cast<const unsigned*>(p->ws + p->lay.off_epoch),
                                                       p->prob.F, z, m, v, g, scalars);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("export launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}
