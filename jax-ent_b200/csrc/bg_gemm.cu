// Tensor-core GEMMs of the batched (replicate / hyper-parameter sweep) mode, sm_100a tcgen05 + TMEM.
//
// With B replicas sharing one feature matrix the two sweeps of an optimise step become dense GEMMs
// (reference: jax.vmap over _optimise_pure, opt/batch.py:113-146, turns frame_average_features,
// utils/jax_fn.py:15-38, and its transpose into (R,F)x(F,B) products):
//
//   GEMM1 (backward)  H[f, b]   = sum_{a,r} N_a[r, f] * c_a[r, b]          M = frames, K = residues, N = replicas
//   GEMM2 (forward)   Acc[a*Rp + r, b] = sum_f N_a[r, f] * e[f, b]          M = residues, K = frames,  N = replicas
//
// Operands are bf16 with fp32 accumulation in tensor memory.  Hard-cutoff contact counts (integers <= 256) are
// exact in bf16; the fp32 replica-side operands (c, e) are split into NS bf16 pieces (hi, mid, lo) that are
// accumulated against the same feature tile, i.e. an NS-fold longer K, which keeps ~8*NS mantissa bits.
//
// No tensor maps: every operand lives in global memory as the byte image of the no-swizzle UMMA canonical layout
// (8 rows x 16 bytes core matrices), so a pipeline stage is filled by plain cp.async.bulk copies:
//
//   features  GF  [fb: 128 frames][a: 2][rg: Rp/8][fc: 16][8 rows][8 frames] bf16
//             the SAME tile is a K-major A operand for GEMM2 (rows = M, SBO 1024/2048, LBO 128) and an MN-major
//             A operand for GEMM1 (frames = M, SBO 128, LBO 2048)
//   c pieces  CC3 [s: NS][a: 2][kc: Rp/8][n: Bn][8 residues] bf16           K-major B operand (SBO 128, LBO 16*Bn)
//   e pieces  E3  [fb][s: NS][kc: 16][n: Bn][8 frames]       bf16           K-major B operand (SBO 128, LBO 16*Bn)
//
// Kernel shape (both GEMMs): warp 0 = bulk-copy producer, warp 1 = MMA issuer (one elected thread, M128 x N=Bn x K16
// instructions), warps 2..5 = epilogue (tcgen05.ld -> registers -> fp32 stores); 2 shared-memory stages of
// K = 64 (16 KB of A + NS x 128*Bn bytes of B) and 2 TMEM accumulator stages (2 x Bn columns) so that the epilogue of
// one tile overlaps the MMAs of the next.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "bg_mma.cuh"
#include "jx_internal.cuh"

namespace jx {
namespace bg {

constexpr int FB = 128;        // frames per frame block
constexpr int STAGE_K = 64;    // K elements per pipeline stage
constexpr int A_STAGE_BYTES = 128 * STAGE_K * 2;  // 16 KB
constexpr int N_STAGES = 2;
constexpr int GEMM_THREADS = 192;  // 6 warps

struct GemmArgs {
  const unsigned char* A;  // GF features
  const unsigned char* B;  // CC3 (mode 1) or E3 (mode 2)
  float* C;                // mode 1: H [n_fb*128][Bn]; mode 2: partials [ksplit][2*Rp][Bn]
  int n_fb;                // frame blocks
  int n_rg;                // Rp / 8
  int Bn;                  // replicas (padded to a multiple of 16, <= 256)
  int NS;                  // bf16 pieces of the fp32 operand (1..3)
  int ksplit;              // mode 2: k-splits
};

// MODE 1: tile = frame block (M = 128 frames); stages = (array, 64-row chunk).   MODE 2: tile = (M tile of 128 rows, k-split);
// stages = 64-frame half blocks of the split's frame range.
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1) bg_gemm_kernel(const GemmArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Bn = a.Bn, NS = a.NS;
  const uint32_t b_stage_bytes = 128u * (uint32_t)Bn;  // 8 k-chunks x Bn x 16 B
  const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)NS * b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)N_STAGES * stage_bytes);
  uint64_t* full = bars;                  // [N_STAGES]
  uint64_t* empty = bars + N_STAGES;      // [N_STAGES]
  uint64_t* tfull = bars + 2 * N_STAGES;  // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  // ---- tile schedule
  const int chunks_per_array = a.n_rg / 8;  // 64-row chunks
  const int n_mt = 2 * (a.n_rg / 16);       // MODE 2: M tiles (both arrays)
  const int n_tiles = (MODE == 1) ? a.n_fb : n_mt * a.ksplit;
  const int n_hb = 2 * a.n_fb;              // 64-frame half blocks

  if (tid == 0) {
    for (int s = 0; s < N_STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull[s]), 1);
      mbar_init(smem_u32(&tempty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 2 accumulator stages of Bn fp32 columns (power of two >= 32 columns in total)
    uint32_t cols = 32;
    while (cols < 2u * (uint32_t)Bn) cols <<= 1;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stages_of = [&](int tile, int& s_begin, int& s_end) {
    if (MODE == 1) {
      s_begin = 0;
      s_end = 2 * chunks_per_array;
    } else {
      const int ks = tile / n_mt;
      s_begin = (int)((long long)n_hb * ks / a.ksplit);
      s_end = (int)((long long)n_hb * (ks + 1) / a.ksplit);
    }
  };

  if (warp == 0) {
    // ================================================================ producer
    if (lane == 0) {
      int slot = 0;
      uint32_t par = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int s0, s1;
        stages_of(tile, s0, s1);
        for (int st = s0; st < s1; ++st) {
          mbar_wait(smem_u32(&empty[slot]), par ^ 1u);
          const uint32_t bar = smem_u32(&full[slot]);
          unsigned char* dst = smem + (size_t)slot * stage_bytes;
          mbar_expect_tx(bar, stage_bytes);
          if (MODE == 1) {
            const int arr = st / chunks_per_array, c8 = st - arr * chunks_per_array;
            const unsigned char* srcA = a.A + (((size_t)tile * 2 + arr) * a.n_rg + (size_t)c8 * 8) * 2048;
            bulk_g2s(smem_u32(dst), srcA, A_STAGE_BYTES, bar);
            for (int s = 0; s < NS; ++s) {
              const unsigned char* srcB = a.B + (((size_t)s * 2 + arr) * a.n_rg + (size_t)c8 * 8) * ((size_t)Bn * 16);
              bulk_g2s(smem_u32(dst + A_STAGE_BYTES + (size_t)s * b_stage_bytes), srcB, b_stage_bytes, bar);
            }
          } else {
            const int mt = tile % n_mt;
            const int arr = mt / (a.n_rg / 16), rg0 = (mt % (a.n_rg / 16)) * 16;
            const int fb = st >> 1, h = st & 1;
            const unsigned char* srcA = a.A + (((size_t)fb * 2 + arr) * a.n_rg + rg0) * 2048 + (size_t)h * 1024;
            for (int g = 0; g < 16; ++g) bulk_g2s(smem_u32(dst + g * 1024), srcA + (size_t)g * 2048, 1024, bar);
            for (int s = 0; s < NS; ++s) {
              const unsigned char* srcB = a.B + (((size_t)fb * NS + s) * 16 + (size_t)h * 8) * ((size_t)Bn * 16);
              bulk_g2s(smem_u32(dst + A_STAGE_BYTES + (size_t)s * b_stage_bytes), srcB, b_stage_bytes, bar);
            }
          }
          if (++slot == N_STAGES) {
            slot = 0;
            par ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, A MN-major in mode 1, N = Bn, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((MODE == 1 ? 1u : 0u) << 15) | ((uint32_t)(Bn >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_lbo = (MODE == 1) ? 2048u : 128u, a_sbo = (MODE == 1) ? 128u : 1024u;
      const uint32_t a_kstep = (MODE == 1) ? 2u * 2048u : 2u * 128u;
      const uint32_t b_lbo = 16u * (uint32_t)Bn, b_sbo = 128u, b_kstep = 32u * (uint32_t)Bn;
      int slot = 0, acc = 0;
      uint32_t par = 0, tpar = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int s0, s1;
        stages_of(tile, s0, s1);
        mbar_wait(smem_u32(&tempty[acc]), tpar ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * Bn);
        uint32_t accumulate = 0;
        for (int st = s0; st < s1; ++st) {
          mbar_wait(smem_u32(&full[slot]), par);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)slot * stage_bytes);
#pragma unroll
          for (int j = 0; j < STAGE_K / 16; ++j) {
            const uint64_t ad = make_desc(sa + j * a_kstep, a_lbo, a_sbo);
            for (int s = 0; s < NS; ++s) {
              const uint64_t bd = make_desc(sa + A_STAGE_BYTES + s * b_stage_bytes + j * b_kstep, b_lbo, b_sbo);
              umma_bf16(d, ad, bd, idesc, accumulate);
              accumulate = 1;
            }
          }
          tc_commit(smem_u32(&empty[slot]));  // frees the stage once these MMAs have read it
          if (++slot == N_STAGES) {
            slot = 0;
            par ^= 1u;
          }
        }
        tc_commit(smem_u32(&tfull[acc]));  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          tpar ^= 1u;
        }
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t tpar = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&tfull[acc]), tpar);
      tc_fence_after();
      const int row = quarter * 32 + lane;
      float* crow;
      if (MODE == 1) crow = a.C + ((size_t)tile * FB + row) * Bn;
      else {
        const int ks = tile / n_mt, mt = tile % n_mt;
        crow = a.C + (((size_t)ks * n_mt + mt) * 128 + row) * Bn;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * Bn);
      for (int c0 = 0; c0 < Bn; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        if (c0 + 32 <= Bn) {
#pragma unroll
          for (int k = 0; k < 32; k += 4)
            *reinterpret_cast<float4*>(crow + c0 + k) = make_float4(__uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c0 + k < Bn) crow[c0 + k] = __uint_as_float(v[k]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty[acc]));
      if (++acc == 2) {
        acc = 0;
        tpar ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    uint32_t cols = 32;
    while (cols < 2u * (uint32_t)Bn) cols <<= 1;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
  }
}

// (R,F) fp32 row-major -> GF bf16 layout; rows >= R and frames >= F are zero.  One thread per (row, 8-frame chunk).
__global__ void pack_gf_kernel(const float* __restrict__ heavy, const float* __restrict__ acceptor, const float* __restrict__ row_off, int Rp,
                               int R, long long F, long long ld, int n_rg, long long n_fb, unsigned char* __restrict__ out) {
  const long long n_items = n_fb * 2 * n_rg * 16 * 8;  // (fb, a, rg, fc, row-in-group)
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += (long long)gridDim.x * blockDim.x) {
    const int ri = (int)(it & 7);
    const int fc = (int)((it >> 3) & 15);
    long long rest = it >> 7;
    const int rg = (int)(rest % n_rg);
    rest /= n_rg;
    const int arr = (int)(rest & 1);
    const long long fb = rest >> 1;
    const int r = rg * 8 + ri;
    const long long f0 = fb * FB + fc * 8;
    const float* src = arr ? acceptor : heavy;
    const float off = (row_off && r < R) ? row_off[arr * Rp + r] : 0.f;
    __nv_bfloat16 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long f = f0 + k;
      v[k] = __float2bfloat16_rn((r < R && f < F) ? src[(long long)r * ld + f] - off : 0.f);  // padding frames stay 0: they carry e' = 0
    }
    *reinterpret_cast<uint4*>(out + it * 16) = *reinterpret_cast<const uint4*>(v);
  }
}

// flag <- 0 unless every value is exactly representable in bf16
__global__ void fit_bf16_kernel(const float* __restrict__ x, const float* __restrict__ off, long long n_rows, long long F, long long ld, int* flag) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  bool bad = false;
  for (long long i = i0; i < n_rows * F; i += stride) {
    const long long r = i / F, f = i - r * F;
    const float raw = x[r * ld + f];
    const float v = raw - (off ? off[r] : 0.f);
    // the stored value must be exact in bf16 AND the subtraction itself exact (raw == v + off in fp32)
    bad = bad || !(__bfloat162float(__float2bfloat16_rn(v)) == v) || (off && v + off[r] != raw);
  }
  if (bad) *flag = 0;
}

}  // namespace bg
}  // namespace jx

using namespace jx;

static size_t bg_gemm_smem(int Bn, int NS) { return (size_t)bg::N_STAGES * (bg::A_STAGE_BYTES + (size_t)NS * 128 * Bn) + 128; }

static int bg_check(const char* what, int n_fb, int n_rg, int Bn, int NS) {
  if (n_fb <= 0 || n_rg <= 0 || n_rg % 16 != 0 || Bn < 16 || Bn > 256 || Bn % 16 != 0 || NS < 1 || NS > 3) {
    set_error("%s: need n_fb > 0, padded residues a multiple of 128, 16 <= Bn <= 256 (multiple of 16), 1 <= NS <= 3", what);
    return JAXENT_E_INVALID;
  }
  return JAXENT_OK;
}

extern "C" size_t jaxent_bg_features_bytes(int32_t R, int64_t F) {
  if (R <= 0 || F <= 0) return 0;
  const size_t Rp = ((size_t)R + 127) / 128 * 128, Fp = ((size_t)F + 127) / 128 * 128;
  return Fp * Rp * 2 * 2;
}

extern "C" int jaxent_bg_pack_features(const float* heavy, const float* acceptor, int32_t R, int64_t F, int64_t ld, void* out,
                                       int32_t* exact_flag, jaxent_stream_t stream) {
  return jaxent_bg_pack_features_centred(heavy, acceptor, nullptr, R, F, ld, out, exact_flag, stream);
}

extern "C" int jaxent_bg_pack_features_centred(const float* heavy, const float* acceptor, const float* row_offsets, int32_t R, int64_t F,
                                               int64_t ld, void* out, int32_t* exact_flag, jaxent_stream_t stream) {
  if (!heavy || !acceptor || !out || R <= 0 || F <= 0 || ld < F) { set_error("bg_pack_features: invalid argument"); return JAXENT_E_INVALID; }
  const int n_rg = (R + 127) / 128 * 16;
  const int Rp = n_rg * 8;
  const long long n_fb = (F + 127) / 128;
  if (exact_flag) {
    bg::fit_bf16_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(heavy, row_offsets, R, F, ld, exact_flag);
    bg::fit_bf16_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(acceptor, row_offsets ? row_offsets + Rp : nullptr, R, F, ld, exact_flag);
  }
  bg::pack_gf_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(heavy, acceptor, row_offsets, Rp, R, F, ld, n_rg, n_fb, static_cast<unsigned char*>(out));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bg_pack_features launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

static int bg_launch(int mode, const bg::GemmArgs& a, int ctas, cudaStream_t s) {
  const size_t smem = bg_gemm_smem(a.Bn, a.NS);
  static size_t configured[3] = {0, 0, 0};
  if (smem > configured[mode]) {
    cudaError_t e = (mode == 1) ? cudaFuncSetAttribute(bg::bg_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(bg::bg_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("bg_gemm: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return JAXENT_E_UNSUPPORTED; }
    configured[mode] = smem;
  }
  if (mode == 1) bg::bg_gemm_kernel<1><<<ctas, bg::GEMM_THREADS, smem, s>>>(a);
  else bg::bg_gemm_kernel<2><<<ctas, bg::GEMM_THREADS, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bg_gemm_kernel<%d> launch: %s", mode, cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}

extern "C" int jaxent_bg_gemm1(const void* features, const void* cc3, float* H, int32_t n_fb, int32_t n_rg, int32_t Bn, int32_t NS,
                               int32_t sm_count, jaxent_stream_t stream) {
  int rc = bg_check("bg_gemm1", n_fb, n_rg, Bn, NS);
  if (rc) return rc;
  if (!features || !cc3 || !H) { set_error("bg_gemm1: NULL argument"); return JAXENT_E_INVALID; }
  bg::GemmArgs a;
  a.A = static_cast<const unsigned char*>(features);
  a.B = static_cast<const unsigned char*>(cc3);
  a.C = H;
  a.n_fb = n_fb;
  a.n_rg = n_rg;
  a.Bn = Bn;
  a.NS = NS;
  a.ksplit = 1;
  if (sm_count <= 0) sm_count = 148;
  return bg_launch(1, a, n_fb < sm_count ? n_fb : sm_count, (cudaStream_t)stream);
}

extern "C" int jaxent_bg_gemm2(const void* features, const void* e3, float* partials, int32_t n_fb, int32_t n_rg, int32_t Bn, int32_t NS,
                               int32_t ksplit, jaxent_stream_t stream) {
  int rc = bg_check("bg_gemm2", n_fb, n_rg, Bn, NS);
  if (rc) return rc;
  if (!features || !e3 || !partials || ksplit < 1 || ksplit > 2 * n_fb) { set_error("bg_gemm2: invalid argument"); return JAXENT_E_INVALID; }
  bg::GemmArgs a;
  a.A = static_cast<const unsigned char*>(features);
  a.B = static_cast<const unsigned char*>(e3);
  a.C = partials;
  a.n_fb = n_fb;
  a.n_rg = n_rg;
  a.Bn = Bn;
  a.NS = NS;
  a.ksplit = ksplit;
  return bg_launch(2, a, 2 * (n_rg / 16) * ksplit, (cudaStream_t)stream);
}
