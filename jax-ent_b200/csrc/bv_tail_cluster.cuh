// Clustered variant of the tail kernel (included from bv_tail.cu, inside namespace jx).
//
// The single-CTA tail is bound by one SM's fp64 issue rate (~20 us at R=500, T=8).  Here the R x T tail
// of the eye-MSE / PF-L2 loss kinds is spread over a thread-block cluster of TAILC CTAs on different SMs:
//   phase 1  rows   [c*R/8, (c+1)*R/8)        : <Nc>,<Nh>, lnPF, exp(-lnPF), uptake        -> L2 scratch
//   phase 2  items  [c*n/8, (c+1)*n/8) of (p,t): S u - y, loss partials, dL/dpred           -> L2 scratch
//   phase 3  rows   (same slice as phase 1)    : back-projection, c_r, cc/ch, partial sums   -> L2 scratch
//   CTA rank 0 folds the 8 partial-sum rows in rank order and writes the control block / loss record.
// Phases are separated by cluster barriers (release/acquire at cluster scope); exchanged data stays in L2
// and is read with ld.global.cg.  Every CTA stages the full peptide-map blob (constant, prefetched before
// the PDL wait) and derives the step scalars redundantly, so no broadcast is needed.
constexpr int TAILC = 8;  // cluster size (portable maximum)

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }

__global__ void __launch_bounds__(TAIL_THREADS, 1) bv_tail_cluster_kernel(const TailArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[32][16];
  __shared__ double s_fin[16];
  __shared__ __align__(16) Ctl s_old;
  __shared__ StepScalars s_sc;
  __shared__ int s_last;

  const JaxentBvProblem& pb = a.prob;
  const int R = pb.R, T = pb.uptake ? pb.T : 0, Tm = T > 0 ? T : 1;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const bool tail_cluster = blockIdx.x < TAILC;
  const int c = (int)cluster_rank();
  const bool frame_cta = (gridDim.x == TAILC) || !tail_cluster;
  const float invTm = 1.0f / (float)Tm;

  // rows / shared-memory views of this CTA
  const int r0 = (int)((long long)R * c / TAILC), r1 = (int)((long long)R * (c + 1) / TAILC), nr = r1 - r0;
  double* sxe = reinterpret_cast<double*>(smem_raw);       // [nr_max] exp(-lnPF) of the own rows
  double* slp_l = sxe + (R / TAILC + 1);                   // [nr_max] lnPF of the own rows
  double* sc_rt = slp_l + (R / TAILC + 1);                 // [Tm][nr_max] back-projection items
  double* sX = sc_rt + (size_t)Tm * (R / TAILC + 1);      // [Tm][R] uptake of ALL rows (PF mode: [R] lnPF), copied from L2
  double* sgp = sX + (size_t)Tm * R;                       // [p_max_tr * Tm] dL/dpred of ALL train peptides, copied from L2
  float* s_k = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sgp + (size_t)a.p_max_tr * Tm) + 15) & ~uintptr_t(15));
  float* s_tp = s_k + ((R + 3) & ~3);
  int* s_blob = reinterpret_cast<int*>(s_tp + ((Tm + 3) & ~3));
  const int nrs = R / TAILC + 1;  // row stride of sc_rt

  // L2 scratch shared by the cluster
  double* g_x = a.cl_scratch;                          // [Tm][R] uptake (or [R] lnPF in PF mode: row 0)
  double* g_lp = g_x + (size_t)Tm * R;                 // [R] lnPF
  double* g_gp = g_lp + R;                             // [p_max_tr * Tm] dL/dpred (train)
  double* g_part = g_gp + (size_t)a.p_max_tr * Tm;     // [TAILC][16] partial sums

  int first_data = -1;
  if (tail_cluster) {
#pragma unroll 1
    for (int i = pb.n_slots - 1; i >= 0; --i)
      if (pb.slots[i].kind <= JAXENT_LOSS_PF_L2) first_data = i;
    if (first_data >= 0) {
      const int4* src = reinterpret_cast<const int4*>(a.blobs.blob[first_data]);
      int4* dst = reinterpret_cast<int4*>(s_blob);
      const int n4 = a.blobs.words[first_data] >> 2;
#pragma unroll 4
      for (int i = tid; i < n4; i += nt) dst[i] = src[i];
    }
    if (T > 0) {
#pragma unroll 1
      for (int i = tid; i < R; i += nt) s_k[i] = pb.k_ints[i];
      if (tid < T) s_tp[tid] = pb.timepoints[tid];
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  const unsigned ep = *a.epoch;
  if (a.ctl[ep & 1u].trk_stop) return;  // the on-device loop has ended: no-op (cluster-uniform)
  const Ctl* __restrict__ old_g = a.ctl + ((ep - 1u) & 1u);
  Ctl* __restrict__ neu = a.ctl + (ep & 1u);
  const int cur_set = ep & 1u;

  double accA0 = 0, accA1 = 0, accB0 = 0, accB1 = 0;
  if (warp == 0) {
    derive_scalars(&s_sc, old_g, a.red, R, a.cfg, a.trk, a.records);
  } else if (warp == 1) {
    const int4* src = reinterpret_cast<const int4*>(old_g);
    int4* dst = reinterpret_cast<int4*>(&s_old);
    if (lane < (int)(sizeof(Ctl) / 16)) dst[lane] = src[lane];
  }
  if (tail_cluster) {
    if (tid < nr) {
      accA0 = a.red[0 * R + r0 + tid];
      accA1 = a.red[1 * R + r0 + tid];
      accB0 = a.red[2 * R + r0 + tid];
      accB1 = a.red[3 * R + r0 + tid];
    }
    if (lane < 16) s_red[warp][lane] = 0.0;
  }
  __syncthreads();  // staged, scalars broadcast
  const int d = s_sc.d;
  const double S = s_sc.S;

  if (tail_cluster) {
    const double bc = (double)s_sc.bc, bh = (double)s_sc.bh;
    double my_a = 0, my_b = 0, my_lp = 0, my_c = 0;
    // ---------------------------------------------------------------- phase 1: own rows
    if (tid < nr) {
      const int r = r0 + tid;
      const double invS = __drcp_rn(S);
      my_a = (d ? accB0 : accA0) * invS;
      my_b = (d ? accB1 : accA1) * invS;
      my_lp = bc * my_a + bh * my_b;
      slp_l[tid] = my_lp;
      g_lp[r] = my_lp;
      a.avg[r] = (float)my_a;
      a.avg[R + r] = (float)my_b;
      a.logpf[r] = (float)my_lp;
      if (T > 0) sxe[tid] = dexp(-my_lp);
    }
    if (T > 0) {
      __syncthreads();
#pragma unroll 1
      for (int it = tid; it < nr * T; it += nt) {
        const int t = fdiv(it, 1.0f / (float)nr), rl = it - t * nr;
        const double e = dexp(-((double)s_k[r0 + rl] * (double)s_tp[t] * sxe[rl]));
        g_x[(size_t)t * R + r0 + rl] = 1.0 - e;
        a.uptake[(size_t)t * R + r0 + rl] = (float)(1.0 - e);
      }
    }
    cluster_sync_all();  // lnPF / uptake of all rows visible in L2
    {
      // one coalesced round trip: bring what phase 2/3 read into shared memory
      const double* src = (T > 0) ? g_x : g_lp;
      const int n = (T > 0) ? T * R : R;
#pragma unroll 4
      for (int i = tid; i < n; i += nt) sX[i] = ldcg(src + i);
    }
    __syncthreads();

    double gbc_reg = 0.0, gbh_reg = 0.0;
#pragma unroll 1
    for (int i = 0; i < pb.n_slots; ++i) {
      const int kind = pb.slots[i].kind;
      const double scale = (double)s_old.fw[i] * (double)s_old.fs[i];
      if (kind == JAXENT_LOSS_PARAMS_L1 || kind == JAXENT_LOSS_PARAMS_L2) {
        const bool l1 = kind == JAXENT_LOSS_PARAMS_L1;
        const double dc = bc - (double)pb.slots[i].prior_bc, dh = bh - (double)pb.slots[i].prior_bh;
        if (tid == 0 && c == 0) s_red[0][i] = s_red[0][JAXENT_MAX_SLOTS + i] = l1 ? fabs(dc) + fabs(dh) : dc * dc + dh * dh;
        gbc_reg += scale * (l1 ? (double)((dc > 0) - (dc < 0)) : 2.0 * dc);
        gbh_reg += scale * (l1 ? (double)((dh > 0) - (dh < 0)) : 2.0 * dh);
        continue;
      }
      if (kind > JAXENT_LOSS_PF_L2) continue;
      if (i != first_data) {
        cluster_sync_all();  // everyone is done with the previous slot's g_gp and blob
        const int4* src = reinterpret_cast<const int4*>(a.blobs.blob[i]);
        int4* dst = reinterpret_cast<int4*>(s_blob);
        const int n4 = a.blobs.words[i] >> 2;
#pragma unroll 1
        for (int k = tid; k < n4; k += nt) dst[k] = src[k];
        __syncthreads();
      }
      const int Ptr = pb.slots[i].train.n_pep, Pva = pb.slots[i].val.n_pep;
      const BlobLayout bl = blob_layout(R, Tm, Ptr, pb.slots[i].train.nnz, Pva, pb.slots[i].val.nnz);
      const bool pf = kind == JAXENT_LOSS_PF_L2;
      const double half = pf ? 1.0 : 0.5;
      // -------------------------------------------------------------- phase 2: own (p,t) items of train and val
#pragma unroll 1
      for (int set = 0; set < 2; ++set) {
        const int P = set ? Pva : Ptr;
        const int n = P * Tm;
        const int j0 = (int)((long long)n * c / TAILC), j1 = (int)((long long)n * (c + 1) / TAILC);
        const int* ptr = s_blob + (set ? bl.va_ptr : bl.tr_ptr);
        const int* col = s_blob + (set ? bl.va_col : bl.tr_col);
        const float* val = reinterpret_cast<const float*>(s_blob + (set ? bl.va_val : bl.tr_val));
        const float* y = reinterpret_cast<const float*>(s_blob + (set ? bl.va_y : bl.tr_y));
        double pl = 0.0;
#pragma unroll 1
        for (int j = j0 + tid; j < j1; j += nt) {
          const int p = fdiv(j, invTm), t = j - p * Tm;
          const double* x = pf ? sX : sX + (size_t)t * R;
          double pred = 0.0;
#pragma unroll 1
          for (int q = ptr[p]; q < ptr[p + 1]; ++q) pred += (double)val[q] * x[col[q]];
          pred -= (double)y[j];
          pl += half * pred * pred;
          if (set == 0) g_gp[j] = pred;
        }
        const double nrm = (P > 0) ? __drcp_rn(pf ? (double)P : (double)T * (double)P) : 0.0;
        pl = wsum(pl * nrm);
        if (lane == 0) s_red[warp][set * JAXENT_MAX_SLOTS + i] = pl;
      }
      cluster_sync_all();  // dL/dpred of all train peptides visible in L2
      {
        const int n = Ptr * Tm;
#pragma unroll 4
        for (int k = tid; k < n; k += nt) sgp[k] = ldcg(g_gp + k);
      }
      __syncthreads();
      // -------------------------------------------------------------- phase 3: back-projection onto own rows
      {
        const int* tptr = s_blob + bl.tr_tptr;
        const int* trow = s_blob + bl.tr_trow;
        const float* tval = reinterpret_cast<const float*>(s_blob + bl.tr_tval);
        const double g = scale * (pf ? 2.0 : 1.0) * __drcp_rn(pf ? (double)Ptr : (double)T * (double)Ptr);
#pragma unroll 1
        for (int it = tid; it < nr * Tm; it += nt) {
          const int t = fdiv(it, 1.0f / (float)nr), rl = it - t * nr;
          const int r = r0 + rl;
          double back = 0.0;
#pragma unroll 1
          for (int q = tptr[r]; q < tptr[r + 1]; ++q) back += (double)tval[q] * sgp[trow[q] * Tm + t];
          if (!pf) {
            const double x = (double)s_k[r] * (double)s_tp[t] * sxe[rl];
            back *= -x * (1.0 - sX[(size_t)t * R + r]);  // du/dlnPF = -x exp(-x)
          }
          sc_rt[(size_t)t * nrs + rl] = g * back;
        }
      }
      __syncthreads();
      if (tid < nr) {
#pragma unroll 1
        for (int t = 0; t < Tm; ++t) my_c += sc_rt[(size_t)t * nrs + tid];
      }
    }

    // ---- own rows: coefficients for the next pass; partial sums of this CTA -> L2
    {
      double v12 = 0.0, v13 = 0.0, v14 = 0.0;
      if (tid < nr) {
        a.cc[r0 + tid] = (float)(my_c * bc);
        a.ch[r0 + tid] = (float)(my_c * bh);
        v12 = my_c * my_lp;
        v13 = my_c * my_a;
        v14 = my_c * my_b;
      }
      if (warp * 32 < nr) {
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
    __syncthreads();
    if (tid < 15) {
      double t = 0.0;
#pragma unroll 1
      for (int w = 0; w < nw; ++w) t += s_red[w][tid];
      g_part[c * 16 + tid] = t;
    }
    cluster_sync_all();  // partial sums of all ranks visible
    if (c == 0) {
      if (tid < 15) {
        double t = 0.0;
#pragma unroll 1
        for (int k = 0; k < TAILC; ++k) t += ldcg(g_part + k * 16 + tid);  // fixed rank order
        s_fin[tid] = t;
      }
      __syncthreads();
      if (tid < 32) write_ctl_and_record(neu, a.records, s_sc, s_old, s_fin, gbc_reg, gbh_reg, pb.n_slots, a.klsum, lane, a.trk);
    }
  }

  // ================================================================== per-frame MaxEnt terms
  if (!frame_cta) {
    if (tid < KL_N) a.klpart[(size_t)blockIdx.x * KL_N + tid] = 0.0;
  } else {
    const long long nf = (gridDim.x == TAILC) ? TAILC : gridDim.x - TAILC;
    const long long cta = (gridDim.x == TAILC) ? blockIdx.x : blockIdx.x - TAILC;
    frame_terms(a, s_sc, s_old, cur_set, cta, nf, s_red);
  }
  fold_kl_partials(a, &s_last);
}
