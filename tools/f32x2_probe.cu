// Does fma.rn.f32x2 (sm_100: FFMA2) halve the issue slots of fp32 code?  Clocks for the same number of fp32 FMAs issued as
// scalar FFMA and as packed FFMA2, 8 warps per SM (two per scheduler, like the chain kernel's epilogue warps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_probe tools/f32x2_probe.cu && /tmp/f32x2_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void k_scalar(float* out, int iters, long long* clk) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = threadIdx.x * 0.001f + i;
  const float m = 1.0001f, c = 0.5f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m), "f"(c));
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
__global__ void k_packed(float* out, int iters, long long* clk) {
  uint64_t a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float lo = threadIdx.x * 0.001f + 2 * i, hi = lo + 1.0f;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
  }
  uint64_t m, c;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(1.0001f));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(0.5f));
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 2000;
  for (int threads : {32, 256, 1024}) {
    long long h;
    k_scalar<<<148, threads>>>(out, iters, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double per_s = (double)h / (iters * 32.0);
    k_packed<<<148, threads>>>(out, iters, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double per_p = (double)h / (iters * 32.0);
    printf("%4d threads/SM: scalar FFMA %.3f clk per fp32 FMA per warp, packed FFMA2 %.3f (%s)\n", threads, per_s, per_p,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
