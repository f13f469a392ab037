// Specialised tcgen05 contraction for the layers that carry the frame: S3 in, S3 out,
//   out = epilogue(A[M,K] . W[N,K]^T)   with the fp32-grade 6-term split product.
//
// What differs from the general kernel in gemm_umma.cu (which stays as the path for pixel-shuffle
// stores, fp32 outputs, two residuals and single-term products):
//   * always a cluster of two CTAs working on one 256 x BN tile with cta_group::2 MMAs: each CTA
//     loads its own 128 rows of A and HALF of the W tile, so the L2 -> SM operand traffic per MMA
//     cycle drops from 64 to 48 B/clk (the chip sustains ~42 B/clk/SM);
//   * K is staged in blocks of 32 (SWIZZLE_64B): a stage is 36 KB instead of 96 KB, four of them are
//     in flight, and TMA latency is covered with half the shared memory;
//   * the epilogue is compile-time specialised (activation / chunk-add pairing / residual) and moves
//     no global memory itself: residual tiles arrive by TMA (one 3-plane box per warp and 16-column
//     chunk, prefetched one chunk ahead) and results leave by TMA store from a swizzled staging
//     tile.  The epilogue of the general kernel was instruction-fetch bound (10 k SASS lines, local
//     memory spills); this one is ~1.5 k instructions per instantiation with no spills.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace dmc {

static char g_s3_err[512] = "";
const char* gemm_s3_last_error() { return g_s3_err; }

// ------------------------------------------------------------------ tensor maps (host)
static PFN_cuTensorMapEncodeTiled_v12000 s3_get_encode() {
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

// 3-D map {columns, rows, 3 planes} of an S3 tensor with a box of {box0 columns, box1 rows, 3 planes}
static int s3_encode(void* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_bytes,
                     uint64_t plane_bytes, uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz) {
  auto fn = s3_get_encode();
  if (!fn) {
    snprintf(g_s3_err, sizeof g_s3_err, "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t dims[3] = {cols, rows, 3};
  cuuint64_t strides[2] = {row_bytes, plane_bytes};
  cuuint32_t box[3] = {box0, box1, 3};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_s3_err, sizeof g_s3_err,
             "cuTensorMapEncodeTiled failed (%d): base=%p dims=%llu,%llu strides=%llu,%llu box=%u,%u", (int)r,
             base, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_bytes,
             (unsigned long long)plane_bytes, box0, box1);
    return -1;
  }
  return 0;
}

// A operand: box = 32 k x 128 rows x 3 planes, SWIZZLE_64B
int make_tmap_s3_act(void* tmap_out, View a, long long M) {
  return s3_encode(tmap_out, a.p, (uint64_t)a.C, (uint64_t)M, (uint64_t)a.ld * 2, (uint64_t)a.ps * 2, 32, 128,
                   CU_TENSOR_MAP_SWIZZLE_64B);
}
// W operand: box = 32 k x BN/2 rows x 3 planes (each CTA of the pair loads half of the tile)
int make_tmap_s3_weight(void* tmap_out, const GemmW& w) {
  return s3_encode(tmap_out, w.w, (uint64_t)w.Kld, (uint64_t)w.Npad, (uint64_t)w.Kld * 2,
                   (uint64_t)w.Npad * w.Kld * 2, 32, (uint32_t)(w.BN / 2), CU_TENSOR_MAP_SWIZZLE_64B);
}
// epilogue tiles (residual in / result out): box = 16 columns x 32 rows x 3 planes, SWIZZLE_32B;
// only the first `cols` columns of the view exist for the map, so partial chunks are clipped by TMA
int make_tmap_s3_rows(void* tmap_out, View v, int cols, long long M) {
  return s3_encode(tmap_out, v.p, (uint64_t)cols, (uint64_t)M, (uint64_t)v.ld * 2, (uint64_t)v.ps * 2, 16, 32,
                   CU_TENSOR_MAP_SWIZZLE_32B);
}

// fp32 result rows [M, ld]: 2-D map, box = 16 columns x 32 rows, SWIZZLE_64B
int make_tmap_f32_rows(void* tmap_out, float* base, int cols, int ld, long long M) {
  auto fn = s3_get_encode();
  if (!fn) {
    snprintf(g_s3_err, sizeof g_s3_err, "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn((CUtensorMap*)tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_s3_err, sizeof g_s3_err, "cuTensorMapEncodeTiled (fp32 rows) failed (%d): base=%p cols=%d ld=%d", (int)r,
             (void*)base, cols, ld);
    return -1;
  }
  return 0;
}

// ------------------------------------------------------------------ device helpers (PTX)
namespace s3 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default .release.cta semantics (what CUTLASS' ClusterBarrier::arrive uses): ordering of the TMEM reads is
  // carried by tcgen05.fence::before_thread_sync, a cluster-scope release would add a full MEMBAR per tile
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline traps (launch error) instead of hanging the GPU.
__device__ __noinline__ void mbar_timeout(int* err, int code) {
  if (err) atomicExch(err, code);
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 6000000000LL) mbar_timeout(err, code);
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// operand load of the pair kernel: bytes complete on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                              uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(0), "r"(leader_bar)
      : "memory");
}
// CTA-local load (residual tiles of the epilogue)
__device__ __forceinline__ void tma_load_local(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                               uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(0), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(0)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() {   // at most one store still reading shared memory
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in both CTAs
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// Issued from warp-uniform code by the lane whose `sel` is non-zero: no divergent branch around the
// instruction, so the (uniform) descriptors can stay in uniform registers.
__device__ __forceinline__ void tc_mma_pair_sel(uint32_t sel, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sel)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair_sel(uint32_t sel, uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar), "h"(mask), "r"(sel)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
               : "memory");
  return v;
}
// two floats -> packed bf16x2 (low half = a), round to nearest even
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// layers.py:8-10  silu(4x)/4 == x / (1 + exp(-4x)); MUFU.RCP instead of an IEEE division (<= 2 ulp
// from the reference's result; the GOP parity tests run through this path)
__device__ __forceinline__ float wsilu_fast(float x) {
  const float e = expf(mul_rn(-4.0f, x));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(add_rn(1.0f, e)));
  return mul_rn(x, r);
}

}  // namespace s3

struct S3Params {
  long long M;
  int m_tiles;     // 256-row tiles
  int n_tiles, k_blocks, BN, stages;
  int n_out;       // destination columns
  int dbg;         // probe switches: 1 no operand loads, 2 no MMA issue, 4 no epilogue TMA traffic, 8 no epilogue math
  const float* bias;
  const float* scale;
  int* err;
};

constexpr int kS3BK = 32;
constexpr int kS3APlane = 128 * kS3BK * 2;        // one plane of a 128 x 32 bf16 tile
constexpr int kS3EpiWarps = 8;
constexpr int kS3Threads = 64 + 32 * kS3EpiWarps; // TMA warp, MMA warp, 8 epilogue warps
constexpr int kS3ChunkBytes = 3 * 32 * 32;        // [3 planes][32 rows][16 bf16]
constexpr int kS3Ring = 3;                        // staging tiles per epilogue warp (residual in -> result out)
constexpr int kS3WarpSmem = kS3Ring * kS3ChunkBytes;
constexpr int kS3BarBytes = 512;

// Shared-memory descriptor of a K-major SWIZZLE_64B operand tile (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO (unused for one swizzle atom along K) = 1 [16,30) | SBO = 8 rows x 64 B = 512
// -> 32 [32,46) | version 1 [46,48) | layout SWIZZLE_64B = 4 [61,64)
__device__ __forceinline__ uint64_t s3_desc(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = 32u | (1u << 14) | (4u << 29);
  return ((uint64_t)hi << 32) | lo;
}

// kF32: the result leaves as fp32 rows [M, ld] (one 2-D TMA store of 16 columns x 32 rows per chunk,
// SWIZZLE_64B staging) instead of S3 planes -- used where the consumer is not a contraction
// (the depthwise 3x3 after dc.0, the pixel-shuffle tail after the reconstruction head).
template <int kAct, int kPack, int kRes, int kF32>
__global__ void __launch_bounds__(kS3Threads, 1)
k_gemm_s3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
          const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
          const S3Params p) {
  using namespace s3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t rank = blockIdx.x & 1u;             // cluster = (2,1,1): rank in the pair, provably warp-uniform
  const bool leader = rank == 0;
  const uint32_t wRows = (uint32_t)p.BN >> 1;
  const uint32_t wPlane = wRows * (kS3BK * 2);
  const uint32_t stageBytes = 3u * (kS3APlane + wPlane);
  const uint32_t epiBase = base + (uint32_t)p.stages * stageBytes;
  const uint32_t barBase = epiBase + kS3EpiWarps * kS3WarpSmem;
  // barriers: full[8] | empty[8] | tfull[2] | tempty[2] | res[8 warps][3 (+1 pad)] | tmem slot
  auto bar_full = [&](int s) { return barBase + 8u * s; };
  auto bar_empty = [&](int s) { return barBase + 64u + 8u * s; };
  auto bar_tfull = [&](int b) { return barBase + 128u + 8u * b; };
  auto bar_tempty = [&](int b) { return barBase + 144u + 8u * b; };
  auto bar_res = [&](int w, int b) { return barBase + 160u + 32u * w + 8u * b; };
  const uint32_t tmemSlot = barBase + 416u;

  if (warp == 0 && lane == 0) {
    if (base & 1023u) {
      if (p.err) atomicExch(p.err, 9);
      __trap();
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    if (kRes) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmRes) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull(b), 1);
      mbar_init(bar_tempty(b), kS3EpiWarps * 2);
    }
    for (int w = 0; w < kS3EpiWarps; ++w) {
      for (int b = 0; b < kS3Ring; ++b) mbar_init(bar_res(w, b), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    const uint32_t ncols = 512;
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmemSlot), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // peer barriers are initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmemSlot - base));

  const int unit = (int)(blockIdx.x >> 1);
  const int units = (int)(gridDim.x >> 1);
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------ operand producer (both CTAs)
    // (whole warp walks the loop, one elected lane issues: addresses stay in uniform registers)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = unit; tile < total_tiles; tile += units) {
      const int mt = tile / p.n_tiles;
      const int m_idx = mt * 256 + (int)rank * 128;
      const int n_idx = (tile - mt * p.n_tiles) * p.BN + (int)(rank * wRows);
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(bar_empty(s), ph ^ 1, p.err, 1);
        const uint32_t sa = base + s * stageBytes;
        const uint32_t sw = sa + 3 * kS3APlane;
        if (elect_one()) {
          if (p.dbg & 1) {
            if (leader) mbar_arrive(bar_full(s));
          } else {
            // both CTAs' bytes complete on the LEADER's barrier, which the leader arms for two stages' worth
            if (leader) mbar_expect_tx(bar_full(s), 2u * stageBytes);
            const uint32_t lbar = mapa(bar_full(s), 0);
            tma_load_pair(sa, &tmA, kb * kS3BK, m_idx, lbar);
            tma_load_pair(sw, &tmW, kb * kS3BK, n_idx, lbar);
          }
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp walks the loop so that every address below is warp-uniform (uniform registers,
    // no per-MMA register -> uniform-register shuffling); one elected lane issues.
    if (leader) {
      // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major both,
      // N>>3 at [17,23), M>>4 at [24,29) with M = 256 for the pair
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | (16u << 24);
      const uint32_t aStep = kS3APlane >> 4, wStep = wPlane >> 4;
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      for (int tile = unit; tile < total_tiles; tile += units, ++tcount) {
        const uint32_t buf = tcount & 1;
        mbar_wait(bar_tempty(buf), ((tcount >> 1) & 1) ^ 1, p.err, 2);
        tc_fence_after();
        const uint32_t d_main = tmem_base + buf * 256u;
        const uint32_t d_small = d_main + 128u;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(bar_full(s), ph, p.err, 3);
          tc_fence_after();
          const uint32_t sa = base + s * stageBytes;
          const uint64_t da = s3_desc(sa);
          const uint64_t dw = s3_desc(sa + 3 * kS3APlane);
          const uint32_t first = kb == 0 ? 0u : 1u;
          if (elect_one()) {
            if (!(p.dbg & 2)) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                // term 0 = hi*hi -> main accumulator; the five small terms (smallest first) -> second one
                // (plane index: 0 hi, 1 mid, 2 lo):  hl, lh, mm, hm, mh
                const uint64_t a0 = da + 2 * ks, a1 = a0 + aStep, a2 = a0 + 2 * aStep;
                const uint64_t w0 = dw + 2 * ks, w1 = w0 + wStep, w2 = w0 + 2 * wStep;
                tc_mma_pair(d_main, a0, w0, idesc, ks == 0 ? first : 1u);
                tc_mma_pair(d_small, a0, w2, idesc, ks == 0 ? first : 1u);
                tc_mma_pair(d_small, a2, w0, idesc, 1u);
                tc_mma_pair(d_small, a1, w1, idesc, 1u);
                tc_mma_pair(d_small, a0, w1, idesc, 1u);
                tc_mma_pair(d_small, a1, w0, idesc, 1u);
              }
            }
            tc_commit_pair(bar_empty(s));
            if (kb == p.k_blocks - 1) tc_commit_pair(bar_tfull(buf));
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = ew >> 2;                  // two warps per quadrant split the columns
    // chunks of 16 destination columns this warp produces per tile, and where they sit in the accumulator
    int nchunk, acc0;
    if (kPack == PACK_PAIR) {
      nchunk = (half * 64 < p.BN) ? 2 : 0;     // one 64-column group (32 values + 32 partners) per warp
      acc0 = half * 64;
    } else {
      nchunk = p.BN >> 5;
      acc0 = half * (p.BN >> 1);
    }
    // ring of staging tiles: chunk k of this warp lives in tile k % kS3Ring -- first as the residual
    // (TMA load, issued one chunk ahead), then overwritten in place by the result (TMA store)
    const uint32_t ringBuf = epiBase + (uint32_t)ew * kS3WarpSmem;
    const uint32_t swz = (uint32_t)((lane >> 2) & 1) << 4;        // SWIZZLE_32B: 16-byte unit ^= row bit 2
    const uint32_t rowOff = (uint32_t)lane * 32u;
    const bool epi_mem = !(p.dbg & 12);

    // destination column of chunk c of tile `tile`
    auto dest_col = [&](int tile, int c) {
      const int n_idx = (tile % p.n_tiles) * p.BN;
      if (kPack == PACK_PAIR) return ((n_idx >> 6) + half) * 32 + 16 * c;
      return n_idx + acc0 + 16 * c;
    };
    auto dest_row = [&](int tile) { return (tile / p.n_tiles) * 256 + (int)rank * 128 + quad * 32; };

    uint32_t ld_slot = 0, ld_par = 0;           // ring position / barrier parity of the next residual load
    uint32_t slot = 0, par = 0;                 // ... of the chunk being processed
    auto issue_res = [&](int tile, int c) {     // whole warp; lane 0 issues
      const int dcol = dest_col(tile, c);
      if (dcol >= p.n_out) return;
      if (lane == 0) {
        mbar_expect_tx(bar_res(ew, ld_slot), kS3ChunkBytes);
        tma_load_local(ringBuf + ld_slot * kS3ChunkBytes, &tmRes, dcol, dest_row(tile), bar_res(ew, ld_slot));
      }
      if (++ld_slot == kS3Ring) { ld_slot = 0; ld_par ^= 1; }
    };
    if (kRes && epi_mem && nchunk > 0 && unit < total_tiles) issue_res(unit, 0);

    uint32_t tcount = 0;
    for (int tile = unit; tile < total_tiles; tile += units, ++tcount) {
      const uint32_t buf = tcount & 1;
      const int n_idx = (tile % p.n_tiles) * p.BN;
      const int row0 = dest_row(tile);
      mbar_wait(bar_tfull(buf), (tcount >> 1) & 1, p.err, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256u;
      if (nchunk == 0 || (p.dbg & 8)) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(bar_tempty(buf));
          else mbar_arrive_cluster(mapa(bar_tempty(buf), 0));
        }
        continue;
      }
      for (int c = 0; c < nchunk; ++c) {
        const int acol = acc0 + 16 * c;                     // accumulator column of this chunk
        const int dcol = dest_col(tile, c);
        const bool valid = dcol < p.n_out;                  // warp-uniform
        // The staging tile two chunks back must have been read by its TMA store before it is reused
        // (by the residual load issued next, or by this chunk's own result when there is no residual).
        if (lane == 0) tma_store_wait_read1();
        __syncwarp();
        if (kRes && epi_mem) {                  // prefetch the residual tile of the next chunk
          int nt = tile, nc = c + 1;
          if (nc == nchunk) { nc = 0; nt += units; }
          if (nt < total_tiles) issue_res(nt, nc);
        }
        float v[16];
        if (valid) {
          uint32_t a[16], b[16];
          tc_ld16(taddr + acol, a);
          tc_ld16(taddr + 128u + acol, b);
          const float4* bp = reinterpret_cast<const float4*>(p.bias + n_idx + acol);
          float bias[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 q = __ldg(bp + i);
            bias[4 * i] = q.x; bias[4 * i + 1] = q.y; bias[4 * i + 2] = q.z; bias[4 * i + 3] = q.w;
          }
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float t = add_rn(add_rn(__uint_as_float(a[i]), __uint_as_float(b[i])), bias[i]);
            if (kAct == ACT_WSILU) t = wsilu_fast(t);
            v[i] = t;
          }
          if (kPack == PACK_PAIR) {
            tc_ld16(taddr + acol + 32, a);
            tc_ld16(taddr + 128u + acol + 32, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 q = __ldg(bp + 8 + i);
              bias[4 * i] = q.x; bias[4 * i + 1] = q.y; bias[4 * i + 2] = q.z; bias[4 * i + 3] = q.w;
            }
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float t = add_rn(add_rn(__uint_as_float(a[i]), __uint_as_float(b[i])), bias[i]);
              if (kAct == ACT_WSILU) t = wsilu_fast(t);
              v[i] = add_rn(v[i], t);
            }
          }
        }
        if (c == nchunk - 1) {                 // accumulator fully read: hand the TMEM buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(bar_tempty(buf));
            else mbar_arrive_cluster(mapa(bar_tempty(buf), 0));
          }
        }
        if (!valid) continue;
        if (kRes && epi_mem) {
          mbar_wait(bar_res(ew, slot), par, p.err, 5);
          const uint32_t src = ringBuf + slot * kS3ChunkBytes + rowOff;
          float t[16];
#pragma unroll
          for (int pl = 2; pl >= 0; --pl) {    // (lo + mid) + hi, exactly join3
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const uint4 q = ld_shared_v4(src + pl * 1024 + (((uint32_t)hf << 4) ^ swz));
              const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float lo = bf16lo(u[k]), hi = bf16hi(u[k]);
                const int i = 8 * hf + 2 * k;
                t[i] = pl == 2 ? lo : add_rn(t[i], lo);
                t[i + 1] = pl == 2 ? hi : add_rn(t[i + 1], hi);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = add_rn(v[i], t[i]);
        }
        if (p.scale) {
          const float4* sp = reinterpret_cast<const float4*>(p.scale + dcol);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 q = make_float4(1.f, 1.f, 1.f, 1.f);
            if (dcol + 4 * i < p.n_out) q = __ldg(sp + i);
            v[4 * i] = mul_rn(v[4 * i], q.x); v[4 * i + 1] = mul_rn(v[4 * i + 1], q.y);
            v[4 * i + 2] = mul_rn(v[4 * i + 2], q.z); v[4 * i + 3] = mul_rn(v[4 * i + 3], q.w);
          }
        }
        if (kF32) {
          // fp32 rows: 64 B per lane, 16-byte unit u of row r sits at u ^ ((r >> 1) & 3)  (SWIZZLE_64B)
          const uint32_t dst = ringBuf + slot * kS3ChunkBytes + (uint32_t)lane * 64u;
          const uint32_t sw64 = (uint32_t)((lane >> 1) & 3) << 4;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            st_shared_v4(dst + (((uint32_t)u << 4) ^ sw64), __float_as_uint(v[4 * u]), __float_as_uint(v[4 * u + 1]),
                         __float_as_uint(v[4 * u + 2]), __float_as_uint(v[4 * u + 3]));
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && epi_mem) tma_store_2d(&tmOut, ringBuf + slot * kS3ChunkBytes, dcol, row0);
          if (++slot == kS3Ring) { slot = 0; par ^= 1; }
          continue;
        }
        // exact 3-way split, two elements per conversion
        uint32_t ph_[8], pm_[8], pl_[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x0 = v[2 * i], x1 = v[2 * i + 1];
          const uint32_t h = cvt_bf16x2(x0, x1);
          const float r0 = sub_rn(x0, bf16lo(h)), r1 = sub_rn(x1, bf16hi(h));
          const uint32_t m = cvt_bf16x2(r0, r1);
          ph_[i] = h;
          pm_[i] = m;
          pl_[i] = cvt_bf16x2(sub_rn(r0, bf16lo(m)), sub_rn(r1, bf16hi(m)));
        }
        const uint32_t dst = ringBuf + slot * kS3ChunkBytes + rowOff;
        st_shared_v4(dst + swz, ph_[0], ph_[1], ph_[2], ph_[3]);
        st_shared_v4(dst + (16u ^ swz), ph_[4], ph_[5], ph_[6], ph_[7]);
        st_shared_v4(dst + 1024 + swz, pm_[0], pm_[1], pm_[2], pm_[3]);
        st_shared_v4(dst + 1024 + (16u ^ swz), pm_[4], pm_[5], pm_[6], pm_[7]);
        st_shared_v4(dst + 2048 + swz, pl_[0], pl_[1], pl_[2], pl_[3]);
        st_shared_v4(dst + 2048 + (16u ^ swz), pl_[4], pl_[5], pl_[6], pl_[7]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && epi_mem) tma_store(&tmOut, ringBuf + slot * kS3ChunkBytes, dcol, row0);
        if (++slot == kS3Ring) { slot = 0; par ^= 1; }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // nobody leaves while the peer can still touch its smem
  if (warp == 1) {
    const uint32_t ncols = 512;
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------ host launcher
static int g_s3_dbg = 0;
void gemm_s3_set_debug(int mask) { g_s3_dbg = mask; }

bool gemm_s3_supports(const GemmW& w, const Epi& e, int nsplit) {
  if (nsplit != 3 || !w.tmap_s3 || e.do_clamp || e.res2.p) return false;
  if (w.BN % 32 || w.BN > 128) return false;
  if (e.out_f32) {                       // fp32 rows: plain layout, no residual, and not both outputs at once
    return !e.out.p && e.pack == PACK_PLAIN && !e.res1.p && (e.act == ACT_NONE || e.act == ACT_WSILU) &&
           e.ld_f32 % 4 == 0 && (uintptr_t)e.out_f32 % 16 == 0;
  }
  if (!e.out.p) return false;
  if (e.pack == PACK_PLAIN) {
    if (e.act == ACT_NONE) return true;
    return e.act == ACT_WSILU && !e.res1.p;
  }
  if (e.pack == PACK_PAIR) return e.act == ACT_WSILU && !e.res1.p && (w.BN % 64 == 0);
  return false;
}

template <int kAct, int kPack, int kRes, int kF32>
static cudaError_t launch_s3(int grid, int smem, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tw,
                             const CUtensorMap& to, const CUtensorMap& tr, const S3Params& p) {
  static bool attr_set = false;
  if (!attr_set) {
    int dev = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaError_t e = cudaFuncSetAttribute(k_gemm_s3<kAct, kPack, kRes, kF32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_max);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kS3Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_gemm_s3<kAct, kPack, kRes, kF32>, ta, tw, to, tr, p);
}

int gemm_s3(const void* tmapA, const GemmW& w, const Epi& e, const void* tmapOut, const void* tmapRes,
            long long M, int K, cudaStream_t st) {
  static int smem_max = 0;
  static int* d_err = nullptr;
  if (!smem_max) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaMalloc(&d_err, sizeof(int));
    cudaMemset(d_err, 0, sizeof(int));
  }
  if (!gemm_s3_supports(w, e, 3)) {
    snprintf(g_s3_err, sizeof g_s3_err, "gemm_s3: unsupported configuration (BN=%d act=%d pack=%d)", w.BN, e.act,
             e.pack);
    return -1;
  }
  S3Params p;
  p.M = M;
  p.m_tiles = (int)((M + 255) / 256);
  p.n_tiles = (w.ncols + w.BN - 1) / w.BN;
  p.k_blocks = (K + kS3BK - 1) / kS3BK;
  p.BN = w.BN;
  p.n_out = e.n_out;
  p.bias = e.bias;
  p.scale = e.scale;
  p.dbg = g_s3_dbg;
  p.err = d_err;
  const int stage_bytes = 3 * (kS3APlane + (w.BN / 2) * kS3BK * 2);
  const int fixed = kS3EpiWarps * kS3WarpSmem + kS3BarBytes;
  int stages = (smem_max - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) {
    snprintf(g_s3_err, sizeof g_s3_err, "gemm_s3: stage of %d bytes does not fit", stage_bytes);
    return -1;
  }
  p.stages = stages;
  const int smem = stages * stage_bytes + fixed;
  int grid = 2 * p.m_tiles * p.n_tiles;
  const int cap = num_sms() & ~1;
  if (grid > cap) grid = cap;
  CUtensorMap ta, tw, to, tr;
  memcpy(&ta, tmapA, sizeof ta);
  memcpy(&tw, w.tmap_s3, sizeof tw);
  memcpy(&to, tmapOut, sizeof to);
  memcpy(&tr, tmapRes ? tmapRes : tmapOut, sizeof tr);
  note_launch();
  cudaError_t err;
  if (e.out_f32 && e.act == ACT_WSILU) err = launch_s3<ACT_WSILU, PACK_PLAIN, 0, 1>(grid, smem, st, ta, tw, to, tr, p);
  else if (e.out_f32) err = launch_s3<ACT_NONE, PACK_PLAIN, 0, 1>(grid, smem, st, ta, tw, to, tr, p);
  else if (e.pack == PACK_PAIR) err = launch_s3<ACT_WSILU, PACK_PAIR, 0, 0>(grid, smem, st, ta, tw, to, tr, p);
  else if (e.act == ACT_WSILU) err = launch_s3<ACT_WSILU, PACK_PLAIN, 0, 0>(grid, smem, st, ta, tw, to, tr, p);
  else if (e.res1.p) err = launch_s3<ACT_NONE, PACK_PLAIN, 1, 0>(grid, smem, st, ta, tw, to, tr, p);
  else err = launch_s3<ACT_NONE, PACK_PLAIN, 0, 0>(grid, smem, st, ta, tw, to, tr, p);
  if (err != cudaSuccess) {
    snprintf(g_s3_err, sizeof g_s3_err, "k_gemm_s3 launch: %s", cudaGetErrorString(err));
    return -1;
  }
  return 0;
}

}  // namespace dmc
