// Microbenchmark: what paces tcgen05.mma.cta_group::2 kind::f16 (M = 256 per CTA pair) when both operands come from
// shared memory -- the tensor pipe itself, the shared-memory operand reads, or other shared-memory traffic on the SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate_probe tools/mma_rate_probe.cu && /tmp/mma_rate_probe
// One CTA pair per two SMs (whole chip), warp 1 of the leader issues MMAs over a ring of static operand stages laid
// out like the chain kernel's (K-major SWIZZLE_32B, two planes); the other warps optionally generate background
// shared-memory traffic: st.shared.v4 streams (what an epilogue's staging writes are) or cp.async.bulk global -> shared
// copies (what the operand loads are).  Prints clocks per MMA.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc32(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = 16u | (1u << 14) | (6u << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void mma_pair(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Params {
  int n;          // MMA N (128 or 256)
  int pattern;    // 0: three N-wide MMAs per k step (a0 w0, a0 w1, a1 w0)   1: same operand pair every time
                  // 2: one 2N-wide MMA (a0, [w0;w1]) + one N-wide (a1, w0)  [n must be 128]
  int ksteps;     // k steps (of 16) to issue in total
  int stages;     // operand ring depth (static contents)
  int bg;         // 0 none, 1 st.shared.v4 streams, 2 cp.async.bulk global -> shared, 3 ld.shared.v4 streams,
                  // 4 FMA + MUFU arithmetic (an activation epilogue's mix), 5 tcgen05.ld.x32 streams (TMEM reads)
  int commit;     // 1: tcgen05.commit to a scratch barrier after every six MMAs (what frees an operand stage)
  int bg_warps;   // warps per CTA generating it
  const uint8_t* src;      // global source for bg == 2 (L2 resident)
  unsigned long long* out; // [cluster][0] clocks of the MMA loop, [1] MMAs issued, [2] background bytes moved by CTA 0
};

constexpr int kThreads = 384;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_probe(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem);
  const uint32_t rank = blockIdx.x & 1u;
  const int aBytes = 2 * 128 * 32 * 2;                 // two planes x 128 rows x 32 k
  const int wRows = (p.pattern == 2 ? 2 * p.n : p.n) / 2;
  const int wBytes = 2 * wRows * 32 * 2;
  const int stageBytes = aBytes + wBytes;
  const uint32_t bgBase = base + p.stages * stageBytes;   // 8 KB per background warp
  const uint32_t barBase = bgBase + 6 * 8192;
  __shared__ uint32_t tmemSlot;
  __shared__ unsigned long long bgBytes;
  __shared__ volatile int stopFlag;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 40; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barBase + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    bgBytes = 0;
    stopFlag = 0;
  }
  // operands: small pseudo-random fp16 values (not zeros: real switching activity)
  for (int i = threadIdx.x; i < p.stages * stageBytes / 4; i += kThreads) {
    uint32_t x = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    x ^= x >> 13;
    const uint32_t h0 = 0x2800u | (x & 0x83ffu), h1 = 0x2800u | ((x >> 16) & 0x83ffu);   // |v| in [2^-5, 2^-4)
    reinterpret_cast<uint32_t*>(smem)[i] = h0 | (h1 << 16);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemSlot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmemSlot;

  if (warp == 1) {
    if (rank == 0) {
      const uint32_t idescN = (1u << 4) | ((uint32_t)(p.n >> 3) << 17) | (16u << 24);
      const uint32_t idesc2N = (1u << 4) | ((uint32_t)((2 * p.n) >> 3) << 17) | (16u << 24);
      const uint32_t wKs = (uint32_t)(wRows * 32) >> 4, wPlane = (uint32_t)(wRows * 64) >> 4;
      const long long t0 = clock64();
      int s = 0;
      for (int k = 0; k < p.ksteps; k += 2) {
        const uint32_t sa = base + s * stageBytes;
        const uint64_t da = desc32(sa), dw = desc32(sa + aBytes);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t a0 = da + ks * (4096u >> 4), a1 = a0 + (8192u >> 4);
          const uint64_t w0 = dw + ks * wKs, w1 = w0 + wPlane;
          const uint32_t acc = (k + ks) ? 1u : 0u;
          if (!elect_one()) continue;
          if (p.pattern == 0) {
            mma_pair(tmem, a0, w0, idescN, acc);
            mma_pair(tmem + 256u, a0, w1, idescN, acc);
            mma_pair(tmem + 256u, a1, w0, idescN, 1u);
          } else if (p.pattern == 1) {
            const uint64_t fa = desc32(base), fw = desc32(base + aBytes);
            mma_pair(tmem, fa, fw, idescN, acc);
            mma_pair(tmem + 256u, fa, fw, idescN, acc);
            mma_pair(tmem + 256u, fa, fw, idescN, 1u);
          } else {
            mma_pair(tmem, a0, w0, idesc2N, acc);           // [w0 ; w1] laid out as 2N contiguous rows
            mma_pair(tmem + 128u, a1, w0, idescN, 1u);
          }
        }
        if (p.commit && elect_one())
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(barBase + 8 * 30), "h"((uint16_t)3) : "memory");
        if (++s == p.stages) s = 0;
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(barBase), "h"((uint16_t)1) : "memory");
      __syncwarp();
      while (!mbar_try(barBase, 0)) {}
      const long long t1 = clock64();
      if (lane == 0) {
        p.out[(blockIdx.x >> 1) * 4 + 0] = (unsigned long long)(t1 - t0);
        p.out[(blockIdx.x >> 1) * 4 + 1] = (unsigned long long)p.ksteps * (p.pattern == 2 ? 2 : 3);
        stopFlag = 1;
        // tell the peer CTA's background warps to stop
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32((const void*)&stopFlag)), "r"(1));
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(1) : "memory");
      }
    }
  } else if (warp >= 4 && warp < 4 + p.bg_warps && p.bg >= 4) {
    unsigned long long moved = 0;
    if (p.bg == 4) {
      float v0 = lane * 0.001f, v1 = v0 + 1.f, v2 = v0 + 2.f, v3 = v0 + 3.f;
      while (!stopFlag) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // per element: ~9 FP32-pipe instructions and 1.5 MUFU, four independent chains
          float e0, e1, e2, e3, r0, r1;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(v0 * -5.77f, 60.f)));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(v1 * -5.77f, 60.f)));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fminf(v2 * -5.77f, 60.f)));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fminf(v3 * -5.77f, 60.f)));
          const float d0 = 1.f + e0, d1 = 1.f + e1, d2 = 1.f + e2, d3 = 1.f + e3;
          asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0 * d1));
          asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d2 * d3));
          v0 = fmaf(v0, r0 * d1, 0.25f); v1 = fmaf(v1, r0 * d0, 0.5f);
          v2 = fmaf(v2, r1 * d3, 0.75f); v3 = fmaf(v3, r1 * d2, 1.0f);
          v0 = fmaf(v0, 0.999f, v1 * 1e-3f); v2 = fmaf(v2, 0.999f, v3 * 1e-3f);
        }
        moved += 64;
      }
      if (v0 + v1 + v2 + v3 == 0.12345f) p.out[3] = 1;
    } else if (p.bg == 6 || p.bg == 7) {
      // an epilogue's store pattern: 4 KB of st.shared per warp, proxy fence, one bulk store shared -> global
      // (bg 7: the fence only, no bulk store)
      const uint32_t my = bgBase + (warp - 4) * 4096;
      uint8_t* dst = const_cast<uint8_t*>(p.src) + ((size_t)blockIdx.x * 8 + (warp - 4)) * 65536;
      uint32_t v = lane;
      while (!stopFlag) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(my + i * 512 + lane * 16), "r"(v) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && p.bg == 6) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (moved & 61440)), "r"(my), "r"(4096) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        moved += 4096;
        ++v;
      }
    } else {
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      uint32_t acc = 0;
      while (!stopFlag) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + (uint32_t)((moved >> 12) & 3) * 32u + 384u)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
        moved += 4096;
      }
      if (acc == 0x12345u) p.out[3] = acc;
    }
    if (lane == 0 && rank == 0) atomicAdd(&bgBytes, moved);
  } else if (warp >= 2 && warp < 2 + p.bg_warps && p.bg && p.bg < 4) {
    const uint32_t my = bgBase + (warp - 2) * 8192;
    unsigned long long moved = 0;
    if (p.bg == 1 || p.bg == 3) {
      uint32_t v = lane;
      while (!stopFlag) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t addr = my + ((i * 512 + lane * 16) & 8191);
          if (p.bg == 1) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(v) : "memory");
          } else {
            uint32_t a, b, c, d;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
            v += a + b + c + d;
          }
        }
        moved += 8 * 512;
      }
      if (v == 0x12345678u) p.out[3] = v;
    } else if (lane == 0) {
      // four 2 KB copies in flight per warp
      const uint32_t bar = barBase + 8 * (1 + (warp - 2) * 4);
      uint32_t ph = 0;
      const uint8_t* src = p.src + ((size_t)blockIdx.x * 8 + (warp - 2)) * 65536;
      int it = 0;
      for (int i = 0; i < 4; ++i) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + 8 * i), "r"(2048) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(my + i * 2048), "l"(src + i * 2048), "r"(2048), "r"(bar + 8 * i) : "memory");
      }
      while (!stopFlag) {
        const int i = it & 3;
        while (!mbar_try(bar + 8 * i, ph)) {}
        moved += 2048;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + 8 * i), "r"(2048) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(my + i * 2048), "l"(src + ((it * 2048) & 65535)), "r"(2048), "r"(bar + 8 * i) : "memory");
        if (i == 3) ph ^= 1;
        ++it;
      }
      for (int i = 0; i < 4; ++i) {                      // drain
        const int j = (it + i) & 3;
        const uint32_t phj = ((it + i) >> 2) & 1;
        while (!mbar_try(bar + 8 * j, phj)) {}
      }
    }
    if (lane == 0 && rank == 0) atomicAdd(&bgBytes, moved);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0 && threadIdx.x == 0) p.out[(blockIdx.x >> 1) * 4 + 2] = bgBytes;
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms & ~1;
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  uint8_t* src = nullptr;
  cudaMalloc(&src, (size_t)grid * 8 * 65536);
  cudaMemset(src, 1, (size_t)grid * 8 * 65536);
  unsigned long long* out = nullptr;
  cudaMalloc(&out, sizeof(unsigned long long) * 4 * grid);
  struct Cfg { const char* name; int n, pattern, bg, bg_warps, commit; };
  const Cfg cfgs[] = {
      {"N=128 3-term                 ", 128, 0, 0, 0, 0},
      {"N=128 3-term + commit/6      ", 128, 0, 0, 0, 1},
      {"N=128 3-term + fma/mufu x8   ", 128, 0, 4, 8, 0},
      {"N=128 + sts/fence/bulk-store x8", 128, 0, 6, 8, 0},
      {"N=128 + sts/fence/bulk-store x4", 128, 0, 6, 4, 0},
      {"N=128 + sts/fence x8         ", 128, 0, 7, 8, 0},
      {"N=128 + commit + store x8    ", 128, 0, 6, 8, 1},
  };
  for (const Cfg& c : cfgs) {
    Params p;
    p.n = c.n; p.pattern = c.pattern; p.ksteps = 4096; p.stages = 4; p.bg = c.bg; p.bg_warps = c.bg_warps; p.commit = c.commit;
    p.src = src; p.out = out;
    float best = 1e30f;
    unsigned long long h[4] = {0, 0, 0, 0};
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemset(out, 0, sizeof(unsigned long long) * 4 * grid);
      k_probe<<<grid, kThreads, smem>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      unsigned long long r[4];
      cudaMemcpy(r, out, sizeof r, cudaMemcpyDeviceToHost);
      const float per = (float)r[0] / (float)r[1];
      if (per < best) { best = per; memcpy(h, r, sizeof h); }
    }
    // normalise to "clocks per 256 x 128 x 16 of MMA work"
    const double work = (c.pattern == 2 ? 1.5 : 1.0) * (c.n / 128.0);
    printf("%s  %7.1f clk / MMA  (%6.1f per 256x128x16)   background %6.1f B/clk/SM\n", c.name, best, best / work,
           (double)h[2] / (double)h[0]);
  }
  return 0;
}
