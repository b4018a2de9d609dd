// Microbenchmark: throughput of scalar FFMA vs packed fma.rn.f32x2 (FFMA2) and add.f32x2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int iters) {
  float a[8], b = 1.0001f, c = 0.5f;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, int iters) {
  unsigned long long a[4], b, c;
  float2 bb = make_float2(1.0001f, 1.0001f), cc = make_float2(0.5f, 0.5f);
  b = *reinterpret_cast<unsigned long long*>(&bb); c = *reinterpret_cast<unsigned long long*>(&cc);
  for (int i = 0; i < 4; ++i) { float2 v = make_float2(threadIdx.x + i, threadIdx.x - i); a[i] = *reinterpret_cast<unsigned long long*>(&v); }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a[i]) : "l"(a[i]), "l"(b), "l"(c));
  float s = 0; for (int i = 0; i < 4; ++i) { float2 v = *reinterpret_cast<float2*>(&a[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_add(float* out, int iters) {
  unsigned long long a[4], c;
  float2 cc = make_float2(0.5f, 0.25f);
  c = *reinterpret_cast<unsigned long long*>(&cc);
  for (int i = 0; i < 4; ++i) { float2 v = make_float2(threadIdx.x + i, threadIdx.x - i); a[i] = *reinterpret_cast<unsigned long long*>(&v); }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(a[i]) : "l"(a[i]), "l"(c));
  float s = 0; for (int i = 0; i < 4; ++i) { float2 v = *reinterpret_cast<float2*>(&a[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 2, block = 1024;
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0); k_scalar<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8 * iters * (double)grid * block;
    printf("scalar FFMA : %.3f ms  %.1f TFLOP/s  (%.1f FMA lanes/clk/SM @1.965GHz)\n", ms, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
    cudaEventRecord(e0); k_packed<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("fma.f32x2   : %.3f ms  %.1f TFLOP/s  (%.1f FMA lanes/clk/SM)\n", ms, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
    cudaEventRecord(e0); k_packed_add<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("add.f32x2   : %.3f ms  %.1f Tadd/s  (%.1f add lanes/clk/SM)\n", ms, fl / 2 / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
