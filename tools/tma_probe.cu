// Microbenchmark: per-SM TMA load / store throughput as a function of the box row width.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu -lcuda && /tmp/tma_probe
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

static PFN_cuTensorMapEncodeTiled_v12000 enc() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return (PFN_cuTensorMapEncodeTiled_v12000)p;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// one thread per CTA: ring of `stages` buffers of `bytes`, loads boxes {inner cols, rows, 1 plane}
__global__ void k_load(const __grid_constant__ CUtensorMap tm, int inner, int rows, int bytes, int stages, int iters,
                       int rows_total, int cols_total) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x != 0) return;
  const uint32_t base = smem_u32(smem);
  const uint32_t bar0 = base + stages * bytes;
  for (int s = 0; s < stages; ++s)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // coordinates advance by plain adds (no divisions in the issue loop): row tile += gridDim, wrap -> next column tile
  const int row_tiles = rows_total / rows;
  int rt = blockIdx.x % row_tiles, c0 = 0;
  int s = 0;
  uint32_t ph = 0;
  for (int i = 0; i < iters + stages; ++i) {
    if (i >= stages) while (!mbar_try(bar0 + 8 * s, ph ^ 1)) {}
    if (i < iters) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(base + s * bytes), "l"(&tm), "r"(c0), "r"(rt * rows), "r"(bar0 + 8 * s) : "memory");
      rt += gridDim.x;
      if (rt >= row_tiles) { rt -= row_tiles; c0 += inner; if (c0 >= cols_total) c0 = 0; }
    }
    if (++s == stages) { s = 0; ph ^= 1; }
  }
}
__global__ void k_store(const __grid_constant__ CUtensorMap tm, int inner, int rows, int bytes, int iters,
                        int rows_total, int cols_total) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x != 0) return;
  const uint32_t base = smem_u32(smem);
  const int row_tiles = rows_total / rows;
  int rt = blockIdx.x % row_tiles, c0 = 0;
  for (int i = 0; i < iters; ++i) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(&tm), "r"(base + (i & 3) * bytes), "r"(c0), "r"(rt * rows) : "memory");
    rt += gridDim.x;
    if (rt >= row_tiles) { rt -= row_tiles; c0 += inner; if (c0 >= cols_total) c0 = 0; }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int rows_total = 38400, cols_total = 1024;     // bf16 [rows][cols], 78.6 MB
  void* d; cudaMalloc(&d, (size_t)rows_total * cols_total * 2); cudaMemset(d, 0, (size_t)rows_total * cols_total * 2);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  auto fn = enc();
  struct Cfg { int inner, rows; CUtensorMapSwizzle swz; const char* name; };
  Cfg cfgs[] = {{16, 32, CU_TENSOR_MAP_SWIZZLE_32B, "16 cols ( 32 B) x  32 rows"},
                {16, 128, CU_TENSOR_MAP_SWIZZLE_32B, "16 cols ( 32 B) x 128 rows"},
                {32, 32, CU_TENSOR_MAP_SWIZZLE_64B, "32 cols ( 64 B) x  32 rows"},
                {32, 128, CU_TENSOR_MAP_SWIZZLE_64B, "32 cols ( 64 B) x 128 rows"},
                {32, 256, CU_TENSOR_MAP_SWIZZLE_64B, "32 cols ( 64 B) x 256 rows"},
                {64, 32, CU_TENSOR_MAP_SWIZZLE_128B, "64 cols (128 B) x  32 rows"},
                {64, 128, CU_TENSOR_MAP_SWIZZLE_128B, "64 cols (128 B) x 128 rows"},
                {64, 256, CU_TENSOR_MAP_SWIZZLE_128B, "64 cols (128 B) x 256 rows"}};
  for (auto& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)cols_total, (cuuint64_t)rows_total};
    cuuint64_t strides[1] = {(cuuint64_t)cols_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)c.inner, (cuuint32_t)c.rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    c.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const int bytes = c.inner * c.rows * 2;
    for (int inflight_kb : {64, 128}) {
      int stages = inflight_kb * 1024 / bytes; if (stages < 2) stages = 2; if (stages > 64) stages = 64;
      const int iters = (int)((long long)24 * 1024 * 1024 / bytes);   // 24 MB per SM
      const int smem = stages * bytes + 8 * 64;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k_load<<<sms, 32, smem>>>(tm, c.inner, c.rows, bytes, stages, 64, rows_total, cols_total);
      cudaEventRecord(e0);
      k_load<<<sms, 32, smem>>>(tm, c.inner, c.rows, bytes, stages, iters, rows_total, cols_total);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double tb = (double)sms * iters * bytes / (ms * 1e-3) / 1e12;
      printf("load  %s  ring %2d x %6d B: %6.2f TB/s  %5.1f B/clk/SM @1.9GHz  (%s)\n", c.name, stages, bytes, tb,
             tb * 1e12 / sms / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    }
    {
      const int iters = (int)((long long)24 * 1024 * 1024 / bytes);
      const int smem = 4 * bytes + 64;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k_store<<<sms, 32, smem>>>(tm, c.inner, c.rows, bytes, 64, rows_total, cols_total);
      cudaEventRecord(e0);
      k_store<<<sms, 32, smem>>>(tm, c.inner, c.rows, bytes, iters, rows_total, cols_total);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double tb = (double)sms * iters * bytes / (ms * 1e-3) / 1e12;
      printf("store %s  4 in flight        : %6.2f TB/s  %5.1f B/clk/SM @1.9GHz  (%s)\n", c.name, tb,
             tb * 1e12 / sms / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
