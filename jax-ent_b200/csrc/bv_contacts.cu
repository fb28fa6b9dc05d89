// BV featuriser on the device: the contact counts that feed the optimise step.
//
// Reference: calc_BV_contacts_universe (jaxent/src/models/func/contacts.py:125-243), called by BV_model.featurise
// (models/HDX/BV/forwardmodel.py:214-300) with amide-N targets / heavy atoms / 6.5 A and amide-H targets / oxygens / 2.4 A.
// Per frame: distance matrix targets x contact atoms (minimum image when a box is given: orthorhombic lengths, or a general
// -- triclinic -- cell as the lower-triangular matrix MDAnalysis derives from `universe.dimensions`, contacts.py:210-215), pairs whose
// residue offset lies in [ignore_lo, ignore_hi] are skipped (:178-189, :218), then either the number of atoms within the
// radius (:236-240) or the sum of the rational switch (1 - x) / (1 - x^2), x = (r / r0)^6, over the atoms within the radius
// (:191-199, :220-234; the removable singularity at r = r0 has the limit 1/2 = 1 / (1 + x), the form used here).
//
// One CTA = 128 targets of one frame; the frame's contact atoms stream through shared memory in chunks of 1024
// (x, y, z, residue id).  Distances are fp32; a pair within 1e-5 (relative) of the cutoff is re-evaluated in fp64 so the
// hard-cutoff counts equal a double-precision evaluation of the same fp32 coordinates (MDAnalysis' distance_array is fp64).
#include "jx_internal.cuh"

namespace jx {

constexpr int CT_THREADS = 128;
constexpr int CT_CHUNK = 1024;

__global__ void __launch_bounds__(CT_THREADS) bv_contacts_kernel(const float* __restrict__ coords, long long F, int A,
                                                                 const int* __restrict__ target_idx, const int* __restrict__ target_resid,
                                                                 int Nt, const int* __restrict__ contact_idx,
                                                                 const int* __restrict__ contact_resid, int Nc, int ignore_lo, int ignore_hi,
                                                                 float radius, int use_switch, const float* __restrict__ box, int box_stride,
                                                                 float* __restrict__ out) {
  __shared__ float4 s_c[CT_CHUNK];  // x, y, z, residue id (bit pattern)
  const int n_tt = (Nt + CT_THREADS - 1) / CT_THREADS;
  const float r2max = radius * radius, inv_r0 = 1.0f / radius;
  for (long long job = blockIdx.x; job < F * n_tt; job += gridDim.x) {
    const long long f = job / n_tt;
    const int t = (int)(job - f * n_tt) * CT_THREADS + threadIdx.x;
    const float* xyz = coords + (size_t)f * A * 3;
    const bool has_box = box != nullptr;
    // cell rows a = (bx, 0, 0), b = (byx, by, 0), c = (czx, czy, bz); orthorhombic: the off-diagonal elements are 0.
    // With radius < min(bx, by, bz) / 2 (checked by the caller) reducing z, then y, then x leaves the ONLY image that can lie
    // within the radius: |d_z| <= bz/2 rules out any other multiple of c, then |d_y| <= by/2 any other multiple of b, ...
    float bx = 0.f, by = 0.f, bz = 0.f, byx = 0.f, czx = 0.f, czy = 0.f;
    if (has_box) {
      if (box_stride == 3) {
        bx = box[f * 3 + 0];
        by = box[f * 3 + 1];
        bz = box[f * 3 + 2];
      } else {  // [bx, byx, by, czx, czy, bz]
        bx = box[f * 6 + 0];
        byx = box[f * 6 + 1];
        by = box[f * 6 + 2];
        czx = box[f * 6 + 3];
        czy = box[f * 6 + 4];
        bz = box[f * 6 + 5];
      }
    }
    float tx = 0.f, ty = 0.f, tz = 0.f;
    int tres = 0;
    if (t < Nt) {
      const int ai = target_idx[t];
      tx = xyz[3 * ai + 0];
      ty = xyz[3 * ai + 1];
      tz = xyz[3 * ai + 2];
      tres = target_resid[t];
    }
    float acc = 0.f;
    for (int c0 = 0; c0 < Nc; c0 += CT_CHUNK) {
      __syncthreads();
      for (int k = threadIdx.x; k < CT_CHUNK && c0 + k < Nc; k += CT_THREADS) {
        const int ai = contact_idx[c0 + k];
        s_c[k] = make_float4(xyz[3 * ai + 0], xyz[3 * ai + 1], xyz[3 * ai + 2], __int_as_float(contact_resid[c0 + k]));
      }
      __syncthreads();
      const int n = min(CT_CHUNK, Nc - c0);
      if (t < Nt) {
        for (int k = 0; k < n; ++k) {
          const float4 c = s_c[k];
          const int d = __float_as_int(c.w) - tres;
          if (d >= ignore_lo && d <= ignore_hi) continue;  // ignored residue window (inclusive)
          float dx = c.x - tx, dy = c.y - ty, dz = c.z - tz;
          if (has_box) {  // minimum image (sequential reduction in the lower-triangular cell)
            const float nz = rintf(dz / bz);
            dz -= nz * bz;
            dy -= nz * czy;
            dx -= nz * czx;
            const float ny = rintf(dy / by);
            dy -= ny * by;
            dx -= ny * byx;
            dx -= bx * rintf(dx / bx);
          }
          float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
          bool in = r2 <= r2max;
          if (fabsf(r2 - r2max) <= 2e-5f * r2max) {  // borderline: decide in double on the same fp32 coordinates
            double ddx = (double)c.x - (double)tx, ddy = (double)c.y - (double)ty, ddz = (double)c.z - (double)tz;
            if (has_box) {
              const double nz = rint(ddz / (double)bz);
              ddz -= nz * (double)bz;
              ddy -= nz * (double)czy;
              ddx -= nz * (double)czx;
              const double ny = rint(ddy / (double)by);
              ddy -= ny * (double)by;
              ddx -= ny * (double)byx;
              ddx -= (double)bx * rint(ddx / (double)bx);
            }
            const double rr = ddx * ddx + ddy * ddy + ddz * ddz;
            in = sqrt(rr) <= (double)radius;
            r2 = (float)rr;
          }
          if (in) {
            if (use_switch) {
              const float q = r2 * inv_r0 * inv_r0;  // (r / r0)^2
              const float x = q * q * q;
              acc += 1.0f / (1.0f + x);
            } else {
              acc += 1.0f;
            }
          }
        }
      }
    }
    if (t < Nt) out[(size_t)t * F + f] = acc;
  }
}

}  // namespace jx

extern "C" int jaxent_bv_contacts(const float* coords, int64_t F, int32_t A, const int32_t* target_idx, const int32_t* target_resid,
                                  int32_t Nt, const int32_t* contact_idx, const int32_t* contact_resid, int32_t Nc, int32_t ignore_lo,
                                  int32_t ignore_hi, float radius, int32_t use_switch, const float* box, float* out,
                                  jaxent_stream_t stream) {
  return jaxent_bv_contacts_cell(coords, F, A, target_idx, target_resid, Nt, contact_idx, contact_resid, Nc, ignore_lo, ignore_hi, radius,
                                 use_switch, box, 3, out, stream);
}

extern "C" int jaxent_bv_contacts_cell(const float* coords, int64_t F, int32_t A, const int32_t* target_idx, const int32_t* target_resid,
                                       int32_t Nt, const int32_t* contact_idx, const int32_t* contact_resid, int32_t Nc, int32_t ignore_lo,
                                       int32_t ignore_hi, float radius, int32_t use_switch, const float* box, int32_t box_stride, float* out,
                                       jaxent_stream_t stream) {
  if (box_stride != 3 && box_stride != 6) { jx::set_error("bv_contacts: box_stride must be 3 (orthorhombic lengths) or 6 (lower-triangular cell)"); return JAXENT_E_INVALID; }
  if (!coords || !target_idx || !target_resid || !contact_idx || !contact_resid || !out || F <= 0 || A <= 0 || Nt <= 0 || Nc <= 0 ||
      !(radius > 0.f) || ignore_lo > ignore_hi) {
    jx::set_error("bv_contacts: invalid argument (F=%lld A=%d Nt=%d Nc=%d radius=%g)", (long long)F, A, Nt, Nc, (double)radius);
    return JAXENT_E_INVALID;
  }
  const long long jobs = (long long)F * ((Nt + jx::CT_THREADS - 1) / jx::CT_THREADS);
  const int grid = (int)(jobs < 148 * 16 ? jobs : 148 * 16);
  jx::bv_contacts_kernel<<<grid, jx::CT_THREADS, 0, (cudaStream_t)stream>>>(coords, F, A, target_idx, target_resid, Nt, contact_idx,
                                                                           contact_resid, Nc, ignore_lo, ignore_hi, radius, use_switch, box,
                                                                           box_stride, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { jx::set_error("bv_contacts launch: %s", cudaGetErrorString(e)); return JAXENT_E_CUDA; }
  return JAXENT_OK;
}
