// tcgen05 weight-gradient kernel (training mode):  part[z][n][k] = sum over the pixel slice z of
// G[m][n] * X[m][k]  --  the contraction of a 1x1 convolution's weight gradient runs over the PIXEL axis.
//
// Both operands are the pixel-major split-fp16 planes of the engine (common.cuh): [C/16 blocks][pixels][16 channels],
// 32 bytes per pixel row with the 16-byte halves swapped where bit 2 of the pixel index is set.  Read with the channel
// axis as the M / N dimension that is byte for byte the canonical MN-MAJOR SWIZZLE_32B operand image of tcgen05
// (mma_traits_sm100.hpp: ((2,n),(8,k)):((1,LBO),(2,SBO)) in 16-byte units): a row = one contraction index (pixel),
// 8 pixels per swizzle atom (SBO = 256 B), one 16-channel block per repeat along M / N (LBO = the block pitch in
// shared memory).  So a stage is a plain TMA copy of 32 pixels x (8 + 16) column blocks x planes, and no transposed
// copy of any activation is ever made.
//
// One CTA per (128 n) x (<= 256 k) output tile and pixel slice: warp 0 TMA producer (4 stages of 48 KB), warp 1 issues
// tcgen05.mma.cta_group::1.kind::f16 with a_major = b_major = MN -- per 16 pixels the three terms of the split product
// (Gh.Xh -> main accumulator; Gh.Xl', Gl'.Xh -> the 2^11-scaled one: 2 x 256 TMEM columns) --, warps 2..5 read the
// accumulators once at the end and store the fp32 partial tile.  k_reduce_partials (train.cu) adds the slices.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace dmc {

static char g_wu_err[512] = "";
const char* wgrad_umma_last_error() { return g_wu_err; }

constexpr int kWuChunk = 32;                       // pixels per stage
constexpr int kWuTN = 128, kWuTK = 256;            // output tile: G channels x X channels
constexpr int kWuGBytes = (kWuTN / 16) * 1024;     // one plane of a stage: [blocks][32 pixels][32 B]
constexpr int kWuXBytes = (kWuTK / 16) * 1024;
constexpr int kWuStages = 4;
constexpr int kWuThreads = 192;                    // TMA warp, MMA warp, 4 epilogue warps (one per TMEM lane quadrant)

static PFN_cuTensorMapEncodeTiled_v12000 wu_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}
// {256 elements = 16 pixels x 16 channels, pixels / 16, column blocks, planes}; box = 32 pixels x `blocks` column blocks of
// one plane (the data is pre-swizzled: no TMA swizzle; blocks past the tensor are zero-filled)
static int wu_tmap(CUtensorMap* out, View v, uint32_t blocks) {
  auto fn = wu_encode();
  if (!fn) {
    snprintf(g_wu_err, sizeof g_wu_err, "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t dims[4] = {256, (cuuint64_t)(v.bs / 256), (cuuint64_t)((v.C + 15) / 16), (cuuint64_t)kPlanes};
  cuuint64_t strides[3] = {512, (cuuint64_t)v.bs * 2, (cuuint64_t)v.ps * 2};
  cuuint32_t box[4] = {256, kWuChunk / 16, blocks, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, v.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_wu_err, sizeof g_wu_err, "wgrad tensor map: cuTensorMapEncodeTiled failed (%d), C=%d", (int)r, v.C);
    return -1;
  }
  return 0;
}

// ---- PTX
__device__ __forceinline__ uint32_t wu_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wu_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void wu_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool wu_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded: a broken pipeline traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void wu_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (wu_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!wu_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err) atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void wu_tma(uint32_t dst, const CUtensorMap* map, int row16, int block, int plane, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(0), "r"(row16), "r"(block), "r"(plane), "r"(bar) : "memory");
}
__device__ __forceinline__ void wu_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wu_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wu_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wu_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void wu_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// MN-major SWIZZLE_32B matrix descriptor: start >> 4 [0,14) | LBO >> 4 [16,30) = pitch of a 16-channel block |
// SBO >> 4 [32,46) = 8 pixels x 32 B | version 1 [46,48) | layout SWIZZLE_32B = 6 [61,64)
__device__ __forceinline__ uint64_t wu_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}

struct WuParams {
  int N, K;                 // G / X channels
  int chunks_total, chunks_per_split;
  int planes;               // 2: three-term split product, 1: hi planes only
  float comp;               // accumulate-truncation compensation (kernels.cu: acc_comp_scaled)
  float* part;
  int* err;
};

__global__ void __launch_bounds__(kWuThreads, 1)
k_wgrad_umma(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WuParams p) {
  extern __shared__ __align__(1024) uint8_t wu_smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kWuStages + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = wu_smem(wu_smem_raw);
  const uint32_t stageBytes = (uint32_t)p.planes * (kWuGBytes + kWuXBytes);
  auto bar_full = [&](int s) { return wu_smem(&bars[s]); };
  auto bar_empty = [&](int s) { return wu_smem(&bars[kWuStages + s]); };
  const uint32_t bar_done = wu_smem(&bars[2 * kWuStages]);

  const int n0 = blockIdx.x * kWuTN, k0 = blockIdx.y * kWuTK;
  const int c_begin = blockIdx.z * p.chunks_per_split;
  int c_end = c_begin + p.chunks_per_split;
  if (c_end > p.chunks_total) c_end = p.chunks_total;
  const int nch = c_end > c_begin ? c_end - c_begin : 0;
  int bn = p.K - k0;                                   // MMA N: the X channels of this tile, a multiple of 16
  if (bn > kWuTK) bn = kWuTK;
  bn = (bn + 15) & ~15;

  if (warp == 0 && lane == 0) {
    if (base & 1023u) {
      if (p.err) atomicExch(p.err, 9);
      __trap();
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int s = 0; s < kWuStages; ++s) {
      wu_bar_init(bar_full(s), 1);
      wu_bar_init(bar_empty(s), 1);
    }
    wu_bar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wu_smem(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  wu_fence_before();
  __syncthreads();
  wu_fence_after();
  pdl_prologue_done();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nch; ++i) {
        const int s = i % kWuStages;
        const uint32_t ph = (i / kWuStages) & 1;
        wu_wait(bar_empty(s), ph ^ 1, p.err, 1);
        const uint32_t sg = base + s * stageBytes;
        const uint32_t sx = sg + p.planes * kWuGBytes;
        wu_expect_tx(bar_full(s), stageBytes);
        const int row16 = (c_begin + i) * (kWuChunk / 16);
        for (int pl = 0; pl < p.planes; ++pl) {
          wu_tma(sg + pl * kWuGBytes, &tmG, row16, n0 >> 4, pl, bar_full(s));
          wu_tma(sx + pl * kWuXBytes, &tmX, row16, k0 >> 4, pl, bar_full(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nch > 0) {
      // instruction descriptor: D = f32 [4,6) = 1, A = B = f16, a_major [15] = b_major [16] = 1 (MN-major),
      // N >> 3 at [17,23), M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kWuTN >> 4) << 24);
      const int nterms = p.planes == 2 ? 3 : 1;
      const int tg[3] = {0, 0, 1};      // plane of G, plane of X per term: hh | hl', l'h
      const int tx[3] = {0, 1, 0};
      for (int i = 0; i < nch; ++i) {
        const int s = i % kWuStages;
        const uint32_t ph = (i / kWuStages) & 1;
        wu_wait(bar_full(s), ph, p.err, 2);
        wu_fence_after();
        const uint32_t sg = base + s * stageBytes;
        const uint32_t sx = sg + p.planes * kWuGBytes;
#pragma unroll
        for (int ks = 0; ks < kWuChunk / 16; ++ks) {
          for (int t = 0; t < nterms; ++t) {
            const uint64_t ad = wu_desc(sg + tg[t] * kWuGBytes + ks * 512);
            const uint64_t bd = wu_desc(sx + tx[t] * kWuXBytes + ks * 512);
            const uint32_t d = t == 0 ? tmem_base : tmem_base + 256u;
            const uint32_t acc = t == 0 ? ((i | ks) ? 1u : 0u) : ((i | ks | (t - 1)) ? 1u : 0u);
            wu_mma(d, ad, bd, idesc, acc);
          }
        }
        wu_commit(bar_empty(s));
      }
      wu_commit(bar_done);
    }
  } else {
    const int quad = warp & 3;
    const int n = n0 + quad * 32 + lane;
    float* dst = p.part + ((long long)blockIdx.z * p.N + n) * p.K + k0;
    const bool row_ok = n < p.N;
    if (nch > 0) {
      wu_wait(bar_done, 0, p.err, 3);
      wu_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c0 = 0; c0 < bn; c0 += 32) {
      uint32_t r[32];
      if (nch > 0) {
        wu_ld32(taddr + c0, r);
        if (p.planes == 2) {
          uint32_t s2[32];
          wu_ld32(taddr + 256u + c0, s2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i)
            r[i] = __float_as_uint(fmaf(fmaf(__uint_as_float(r[i]), p.comp, __uint_as_float(s2[i])), kLoInv, __uint_as_float(r[i])));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (k0 + c0 + i < p.K)      // K is a multiple of 4 (channels come in multiples of 16)
            *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                   __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        }
      }
    }
  }
  wu_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static void wu_geometry(long long M, int N, int K, int& chunks, int& per, int& S) {
  chunks = (int)((M + kWuChunk - 1) / kWuChunk);
  const int tiles = ((N + kWuTN - 1) / kWuTN) * ((K + kWuTK - 1) / kWuTK);
  S = num_sms() / tiles;                // one CTA per SM (192 KB of stages, all 512 TMEM columns)
  if (S < 1) S = 1;
  if (S > chunks) S = chunks;
  per = (chunks + S - 1) / S;
  S = (chunks + per - 1) / per;
}
int wgrad_umma_splits(long long M, int N, int K) {
  int chunks, per, S;
  wu_geometry(M, N, K, chunks, per, S);
  return S;
}
bool wgrad_umma_supported(View G, View X) {
  static int on = -1;
  if (on < 0) {
    const char* v = getenv("DMC_WGRAD_UMMA");      // DMC_WGRAD_UMMA=0: the mma.sync kernel of train.cu (A/B runs)
    on = (v && v[0] == '0') ? 0 : 1;
  }
  return on == 1 && G.C % 16 == 0 && X.C % 16 == 0 && (uintptr_t)G.p % 512 == 0 && (uintptr_t)X.p % 512 == 0 &&
         G.bs % 256 == 0 && X.bs % 256 == 0;
}
// returns the number of partial matrices written to `part` (N * K floats each), or -1
int wgrad_umma(View G, View X, long long M, int terms, float* part, cudaStream_t st) {
  static int* d_err_dev[64] = {};
  static bool attr_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  const int planes = terms == 1 ? 1 : 2;
  const int smem = kWuStages * kPlanes * (kWuGBytes + kWuXBytes);
  if (!attr_dev[dev]) {
    if (cudaFuncSetAttribute(k_wgrad_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      snprintf(g_wu_err, sizeof g_wu_err, "cudaFuncSetAttribute(k_wgrad_umma, %d) failed", smem);
      cudaGetLastError();
      return -1;
    }
    cudaMalloc(&d_err_dev[dev], sizeof(int));
    cudaMemset(d_err_dev[dev], 0, sizeof(int));
    attr_dev[dev] = true;
  }
  CUtensorMap tmG, tmX;
  if (wu_tmap(&tmG, G, kWuTN / 16) != 0 || wu_tmap(&tmX, X, kWuTK / 16) != 0) return -1;
  WuParams p;
  p.N = G.C; p.K = X.C;
  int S;
  wu_geometry(M, p.N, p.K, p.chunks_total, p.chunks_per_split, S);
  p.planes = planes;
  p.comp = planes == 2 ? acc_comp_scaled(p.chunks_per_split * kWuChunk) : 0.0f;
  p.part = part;
  p.err = d_err_dev[dev];
  dim3 grid((p.N + kWuTN - 1) / kWuTN, (p.K + kWuTK - 1) / kWuTK, S);
  note_launch();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kWuThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, k_wgrad_umma, tmG, tmX, p) != cudaSuccess) {
    snprintf(g_wu_err, sizeof g_wu_err, "k_wgrad_umma launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return S;
}

}  // namespace dmc
