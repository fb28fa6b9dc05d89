#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void thr(T* out, int iters) {
  T a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const T b = 1.000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
    a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <typename T>
__global__ void lat(T* out, int iters) {
  T a0 = threadIdx.x * 1e-3;
  const T b = 1.000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) a0 = a0 * b + c;
  out[threadIdx.x] = a0;
}
template <typename T>
void run(const char* name) {
  T* d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(T));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  thr<T><<<148 * 8, 256>>>(d, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); thr<T><<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = 148.0 * 8 * 256 * 8 * iters;
  printf("%s throughput: %.2f TFMA/s (%.1f FMA/clk/SM at 1.965 GHz)\n", name, fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
  cudaEventRecord(e0); lat<T><<<1, 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%s dependent-chain latency: %.1f cycles per FMA\n", name, ms * 1e-3 * 1.965e9 / iters);
}
int main() { run<float>("fp32"); run<double>("fp64"); return 0; }
