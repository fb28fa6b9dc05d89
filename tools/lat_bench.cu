// Single-CTA latency / contention micro-benchmark for the R x T tail: cycles per dependent operation with W warps resident on ONE SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int n) {
  __shared__ double sd[4096];
  __shared__ int si[4096];
  const int tid = threadIdx.x;
  for (int i = tid; i < 4096; i += blockDim.x) { sd[i] = 1.0 + 1e-9 * i; si[i] = (i * 37 + 11) & 4095; }
  __syncthreads();
  long long t0, t1;
  // (0) dependent DFMA chain
  double a = tid * 1e-3;
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = a * 1.000001 + 1e-7;
  t1 = clock64();
  long long r0 = t1 - t0;
  __syncthreads();
  // (1) dependent LDS chain (pointer chase)
  int p = tid & 4095;
  t0 = clock64();
  for (int i = 0; i < n; ++i) p = si[p];
  t1 = clock64();
  long long r1 = t1 - t0;
  __syncthreads();
  // (2) dependent F2F.F64.F32 + DADD chain
  float f = tid;
  double d = 0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { d = (double)f + d; f = (float)d; }
  t1 = clock64();
  long long r2 = t1 - t0;
  __syncthreads();
  // (3) LDS(index) -> LDS.64 -> DFMA(acc), as the peptide-map loops
  double acc = 0; int q = tid & 4095;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { const int c = si[q]; acc += 0.5 * sd[c]; q = (q + 1) & 4095; }
  t1 = clock64();
  long long r3 = t1 - t0;
  __syncthreads();
  // (4) exp(double) chain
  double e = 1e-3 * tid;
  t0 = clock64();
  for (int i = 0; i < n; ++i) e = exp(-e);
  t1 = clock64();
  long long r4 = t1 - t0;
  if ((tid & 31) == 0) {
    long long* o = out + (tid >> 5) * 8;
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4; o[5] = (long long)(a + p + d + acc + e);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 32 * 8 * 8);
  long long h[32 * 8];
  const int n = 256;
  for (int threads : {32, 128, 512, 1024}) {
    k<<<1, threads>>>(d, n); cudaDeviceSynchronize();
    k<<<1, threads>>>(d, n); cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%4d threads: cycles per op (warp 0): DFMA %.1f  LDS-chase %.1f  F2F+DADD+F2F %.1f  LDS->LDS.64->DFMA %.1f  exp(double) %.1f\n", threads,
           h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n, h[4] / (double)n);
  }
  return 0;
}
