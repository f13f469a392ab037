// Persistent tcgen05 "chain" kernel for the layers that carry the frame: S3 in, S3 (or fp32 rows) out,
//   out_l = epilogue_l(A_l[M,K_l] . W_l[N_l,K_l]^T),  l = 0 .. L-1,
// where every layer is a 1x1 convolution over the same M rows and A_l is normally the output of
// layer l-1 (a DepthConvBlock is dc.3 -> ffn.0 -> ffn.2 -> the next block's dc.0).  ONE launch walks a
// host-built table of (layer, 256-row tile, N tile) entries; a tile of layer l starts as soon as all
// N tiles of layer l-1 for the SAME rows are stored (a global counter per (layer, row tile), release /
// acquire at gpu scope), so
//   * 148 SMs share  sum_l tiles_l  work items instead of rounding every layer up to whole waves
//     (M = 38400 gives 4.05 waves of 256 x 128 tiles per 256-channel layer: 19 % idle when launched alone),
//   * there are no launch gaps / pipeline refills / drained tails between the layers of a chain,
//   * the intermediate activations are consumed a few tiles after they are written: they come from L2.
// The table is layer-major, every dependency points backwards in it and all clusters are resident
// (grid <= #SM), so waits cannot deadlock; they are bounded anyway and trap instead of hanging.
//
// Inside a tile (what differs from the general kernel in gemm_umma.cu, which stays as the path for
// pixel-shuffle stores and odd shapes):
//   * a cluster of two CTAs works on one 256 x BN tile with cta_group::2 MMAs: each CTA loads its own 128 rows of A and
//     HALF of the W tile (L2 -> SM operand traffic per MMA cycle 64 -> 48 B/clk);
//   * K is staged in blocks of 32 (two 16-wide SWIZZLE_32B k blocks): a stage is 24 KB, six are in flight;
//   * warp roles: 0 issues the A loads, 3 the W loads (a TMA issue costs its warp a few hundred clocks: one warp doing
//     both bound the operand supply), 1 issues the MMAs -- two stages per trip through its loop, descriptors as 32-bit
//     words, because the tensor pipe queues hardly anything and idles through whatever else that warp does --, 2
//     publishes finished tiles (the gpu-scope release), 4..11 are the epilogue;
//   * every issuing loop runs warp-uniformly with one elected lane, so descriptors and addresses live in uniform
//     registers, and carries nothing that ptxas would spill: fence.proxy.async.global (every publication) and
//     ld.acquire.gpu (every dependency check) invalidate the whole L1, a spilled loop counter then costs an L2 round trip
//     per tile;
//   * the epilogue is one body with run-time layer kinds (its code streams through the instruction cache of eight
//     warps: a second unrolled copy cost 3 % of a DepthConvBlock) and moves no global memory itself: residual tiles
//     arrive by TMA (one 2-plane box per warp and 32-column chunk, prefetched one chunk ahead), results leave by TMA
//     store from a swizzled staging tile (ring of two per warp), the bias comes through shared memory.
// DESIGN.md 3.1 has the measurements behind each of these.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "kernels.h"

namespace dmc {

static char g_s3_err[512] = "";
const char* gemm_s3_last_error() { return g_s3_err; }
// 0, or which bounded wait of the chain kernel timed out before it trapped: 1 operand ring slot, 2 TMEM buffer,
// 3 operand bytes, 4 accumulator, 5 residual tile, 6 previous layer's row tile (global counter), 7 publisher waiting for the
// epilogue warps, 9 smem alignment
static volatile int* g_s3_trap = nullptr;
int gemm_s3_trap_code() { return g_s3_trap ? *g_s3_trap : 0; }

#ifndef DMC_S3_EPI_WARPS
#define DMC_S3_EPI_WARPS 8
#endif
#if DMC_S3_EPI_WARPS != 8 && DMC_S3_EPI_WARPS != 16
#error "eight or sixteen epilogue warps"
#endif
constexpr int kS3EpiWarps = DMC_S3_EPI_WARPS;
constexpr int kS3Split = kS3EpiWarps / 4;         // warps per TMEM lane quadrant: they split the tile's columns
// warp group 0: TMA producer, MMA issuer, two spare warps (56 registers); warp groups 1..: the epilogue warps
// (setmaxnreg: 224 registers with eight of them; sixteen would get 112, which the inlined epilogue variants do not
// fit -- ptxas spills ~2 KB per thread -- so the latency hiding comes from 32-column chunks instead).
// Sixteen epilogue warps (-DDMC_S3_EPI_WARPS=16) work in 16-column chunks -- one column block of the S3 layout, half the
// registers per chunk, which is what fits the 112 registers a 640-thread CTA leaves them.
constexpr int kS3ChunkCols = kS3EpiWarps == 16 ? 16 : 32;
constexpr int kS3ChunkBlocks = kS3ChunkCols / 16;           // 16-column blocks per chunk
constexpr int kS3PairUnits = 32 / kS3ChunkCols;             // chunks per 64-column chunk-add group (32 outputs)
constexpr int kS3ChunkBytes = kPlanes * 32 * kS3ChunkCols * 2;

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

// 4-D map of a tile-blocked S3 view (common.cuh): {256 elements = 16 rows x 16 columns (512 contiguous bytes),
// padded rows / 16, column blocks, 2 planes}; a box is {256, rows / 16, blocks, planes}.  No TMA swizzle: the data
// is stored pre-swizzled, global memory and shared memory images are identical.
static int s3_encode(void* out, View v, int cols, uint32_t box_rows, uint32_t box_blocks, uint32_t planes) {
  auto fn = s3_get_encode();
  if (!fn) {
    snprintf(g_s3_err, sizeof g_s3_err, "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t dims[4] = {256, (cuuint64_t)(v.bs / 256), (cuuint64_t)((cols + 15) / 16), (cuuint64_t)kPlanes};
  cuuint64_t strides[3] = {512, (cuuint64_t)v.bs * 2, (cuuint64_t)v.ps * 2};
  cuuint32_t box[4] = {256, box_rows / 16, box_blocks, planes};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, v.p, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_s3_err, sizeof g_s3_err,
             "cuTensorMapEncodeTiled failed (%d): base=%p cols=%d bs=%lld ps=%lld box=%u rows x %u blocks x %u planes",
             (int)r, (void*)v.p, cols, v.bs, v.ps, box_rows, box_blocks, planes);
    return -1;
  }
  return 0;
}

// A operand: box = 128 rows x 2 column blocks (32 k) x `planes` planes (2, or 1 = hi only for single-term products)
// 16-wide k blocks per operand stage of a single-term layer: 4 = 64 k of the hi plane, the same bytes in flight per
// stage as 32 k of both planes (DMC_S3_SINGLE_KBLK=2 restores half-filled stages for A/B runs)
static int s3_single_kblk() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("DMC_S3_SINGLE_KBLK");
    v = (e && e[0] == '2') ? 2 : 4;
  }
  return v;
}
int make_tmap_s3_act(void* tmap_out, View a, long long M, int planes) {
  (void)M;
  return s3_encode(tmap_out, a, a.C, 128, planes == 1 ? s3_single_kblk() : 2, (uint32_t)planes);
}
// A operand for 4-CTA clusters: box = 64 rows x 1 column block x 1 plane (one contiguous 2 KB piece)
int make_tmap_s3_act64(void* tmap_out, View a, long long M) {
  (void)M;
  return s3_encode(tmap_out, a, a.C, 64, 1, 1);
}
// W operand, from the tile-blocked copy [2][Kld/16][Npad][16]: 4-D map {256 elements, Npad/16, Kld/16, 2} with a box
// of {256, BN/32, 2 k blocks, planes}: the BN/2 x 32 half tile of one CTA arrives as BN/16 segments of 512 bytes per
// plane (row-major it was BN/2 segments of 64 bytes, and the TMA unit retires well under one segment per clock)
int make_tmap_s3_weight(void* tmap_out, const GemmW& w, int planes) {
  auto fn = s3_get_encode();
  if (!fn || !w.wb) {
    snprintf(g_s3_err, sizeof g_s3_err, "make_tmap_s3_weight: no blocked weight copy / encode entry point");
    return -1;
  }
  const uint64_t plane_bytes = (uint64_t)w.Npad * w.Kld * 2;
  cuuint64_t dims[4] = {256, (cuuint64_t)w.Npad / 16, (cuuint64_t)w.Kld / 16, (cuuint64_t)kPlanes};
  cuuint64_t strides[3] = {512, (cuuint64_t)w.Npad * 32, plane_bytes};
  cuuint32_t box[4] = {256, (cuuint32_t)(w.BN / 32), planes == 1 ? (cuuint32_t)s3_single_kblk() : 2u, (cuuint32_t)planes};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn((CUtensorMap*)tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, w.wb, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_s3_err, sizeof g_s3_err, "cuTensorMapEncodeTiled (blocked W) failed (%d): Npad=%d Kld=%d BN=%d", (int)r,
             w.Npad, w.Kld, w.BN);
    return -1;
  }
  return 0;
}
// epilogue tiles (residual in / result out): box = 32 rows x 2 column blocks (32 columns) x 2 planes = four
// contiguous 1 KB pieces; only the first `cols` columns of the view exist for the map (a box that reaches past
// them is clipped on store and zero-filled on load)
int make_tmap_s3_rows(void* tmap_out, View v, int cols, long long M) {
  (void)M;
  return s3_encode(tmap_out, v, cols, 32, kS3ChunkBlocks, kPlanes);
}
// fp32 result rows [M, ld]: 2-D map, box = 32 columns x 32 rows, SWIZZLE_128B
int make_tmap_f32_rows(void* tmap_out, float* base, int cols, int ld, long long M) {
  auto fn = s3_get_encode();
  if (!fn) {
    snprintf(g_s3_err, sizeof g_s3_err, "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kS3ChunkCols, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn((CUtensorMap*)tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, kS3ChunkCols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
  if (err) *reinterpret_cast<volatile int*>(err) = code;
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
// (tile-blocked maps: coordinates are {0, row / 16, column block, plane})
__device__ __forceinline__ void tma_load_pair(uint32_t dst, const CUtensorMap* map, int row16, int blk,
                                              uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(0), "r"(row16), "r"(blk), "r"(0), "r"(leader_bar)
      : "memory");
}
// W half tile from the blocked weight copy (4-D map: element, 16-row group, 16-wide k block, plane)
__device__ __forceinline__ void tma_load_pair_w(uint32_t dst, const CUtensorMap* map, int row16, int kblk16,
                                                uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(0), "r"(row16), "r"(kblk16), "r"(0), "r"(leader_bar)
      : "memory");
}
// same, delivered to every CTA of `mask` (cluster ranks) at the same smem offset; each copy completes on the
// barrier at this offset in the destination's pair leader (the address carries the issuer's leader: peer bit clear)
__device__ __forceinline__ void tma_load_pair_mcast(uint32_t dst, const CUtensorMap* map, int row16, int blk, int plane,
                                                    uint32_t leader_bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
      ::"r"(dst), "l"(map), "r"(0), "r"(row16), "r"(blk), "r"(plane), "r"(leader_bar), "h"(mask)
      : "memory");
}
// CTA-local load (residual tiles of the epilogue)
__device__ __forceinline__ void tma_load_local(uint32_t dst, const CUtensorMap* map, int col, int row,
                                               uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(0), "r"(row >> 4), "r"(col >> 4), "r"(0), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store(const CUtensorMap* map, uint32_t src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(0), "r"(row >> 4), "r"(col >> 4), "r"(0)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() {   // every bulk store has read its shared memory
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() {   // at most one store still reading shared memory
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read2() {   // at most two stores still reading shared memory
  asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_1() {       // all but the newest bulk store are complete
  asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
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
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t mask) {   // arrives on `bar` in the CTAs of `mask`
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
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
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
// layers.py:8-10  silu(4x)/4 == x / (1 + exp(-4x)); MUFU.RCP instead of an IEEE division (<= 2 ulp
// from the reference's result; the GOP parity tests run through this path)
__device__ __forceinline__ float wsilu_fast(float x) {
  const float e = expf(mul_rn(-4.0f, x));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(add_rn(1.0f, e)));
  return mul_rn(x, r);
}

// Two WSiLUs with ONE reciprocal: x0 / d0 and x1 / d1 with d = 1 + exp(-4x) are x0 * (r * d1) and x1 * (r * d0)
// for r = 1 / (d0 * d1).  The SFU (16 ops/clk/SM) is what bounds the activation epilogues: 1.5 MUFU per element
// instead of 2, and ex2.approx on a pre-scaled argument instead of expf (5 instructions fewer).  The exponent is
// clamped at 2^60 so the product of the two denominators stays finite: for x < -10.4 the result is then
// x * 2^-60 instead of something smaller still (|difference| < 1e-17).  Within a few ulp of wsilu().
__device__ __forceinline__ void wsilu2_fast(float& x0, float& x1) {
  const float c = -5.7707801635558535f;            // -4 * log2(e)
  float e0, e1, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(mul_rn(x0, c), 60.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(mul_rn(x1, c), 60.0f)));
  const float d0 = add_rn(1.0f, e0), d1 = add_rn(1.0f, e1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(mul_rn(d0, d1)));
  x0 = mul_rn(x0, mul_rn(r, d1));
  x1 = mul_rn(x1, mul_rn(r, d0));
}

}  // namespace s3


enum { S3_PLAIN = 0, S3_WSILU = 1, S3_RES = 2, S3_PAIR = 3, S3_F32 = 4, S3_F32_WSILU = 5, S3_RES2 = 6 };
constexpr int kS3MaxStages = 5;

struct alignas(64) S3StageDev {
  CUtensorMap tmA, tmW, tmOut, tmRes, tmRes2;
  CUtensorMap tmA64;     // A as 32 k x 64 rows x 1 plane boxes (4-CTA clusters: halves multicast between the two pairs)
  const float* bias;
  const float* scale;
  int k_blocks, BN, n_tiles, n_out;
  int kblk;          // 16-wide k blocks per operand stage: 2 (both planes) or 4 (single term: hi plane only)
  int kind;          // S3_*
  int nterms;        // 3: fp32-grade split product (two accumulators), 1: hi*hi only
  float comp;        // accumulate-truncation compensation, pre-scaled by 2^11 (see acc_comp_scaled in kernels.cu); 0 = off
  uint32_t need;     // increments of done[l-1][row tile] per launch that complete layer l-1 for a row tile
  int publish;       // a later layer of the chain waits for this one: completed tiles are counted in done[l]
};
struct S3ChainParams {
  S3StageDev st[kS3MaxStages];
  const uint32_t* table;   // entries: layer << 28 | n tile << 20 | row tile
  uint32_t* done;          // [layers][row tiles], monotonic across launches
  int n_entries, MT;
  uint32_t* epoch_ctr;     // device words {launches of this chain completed so far, CTAs of this launch that have left}:
                           // a launch reads its number (1-based) here -- layer l-1 is complete at done == number * need --
                           // and its last CTA to leave bumps it, so a launch captured in a CUDA graph can be replayed
  uint32_t stageBytes;     // operand ring stage (sized for the widest W tile of the chain)
  int stages;
  int eager;               // 1: few tiles per layer and cluster -- publish a tile as soon as it is stored (below)
  int cl4;                 // 1: clusters of 4 CTAs = two pairs on the two N tiles of an entry, sharing A by TMA multicast
  int dbg;                 // probe switches: 1 no operand loads, 2 no MMA issue, 4 no epilogue TMA traffic, 8 no epilogue math
  int* err;
};

constexpr int kS3BK = 32;
constexpr int kS3APlane = 128 * kS3BK * 2;        // one plane of a 128 x 32 fp16 tile
constexpr int kS3Threads = 128 + 32 * kS3EpiWarps;
// setmaxnreg only moves registers INSIDE the CTA's launch allocation (threads x the count ptxas reports: 384 x 168
// or 640 x 96): what warp group 0 gives up (down to 56) is all the epilogue warp groups can take -- 2 x 128 x 56 = 14 336
// -> 224 each with eight warps; 4 x 128 x 8 = 4 096 of the 5 120 freed -> 104 each with sixteen (asking for 112 blocks
// the last warp group forever).
constexpr int kS3EpiRegs = kS3EpiWarps == 16 ? 104 : 216;   // (2 x 128 x 216 + 128 x 72 = 64 512)
// The epilogue works in chunks of 32 rows x 32 columns per warp (16-column chunks left the warp waiting on one
// latency after the other: tcgen05.ld, bias loads, the shared-memory fence, the TMA issue -- 1 560 clocks per chunk
// measured, 6 200 per tile against 3 900 for the MMAs).  Staging tile of a chunk: split planes
// [2 planes][2 column blocks][32 rows][16 fp16], or fp32 rows [32 rows][32 fp32] in SWIZZLE_128B order.
#ifndef DMC_S3_RING
#define DMC_S3_RING 2
#endif
// Staging tiles per epilogue warp (residual in -> result out).  Two: the residual of chunk c+1 is prefetched into the
// tile chunk c-1 was stored from -- that store (issued a whole chunk earlier) has read its data long before, the wait
// for it is a formality -- and the 32 KB a third tile per warp would take go to the operand ring, whose bytes in
// flight are what bounds the operand supply (5 stages = 120 KB per CTA at ~4 000 clocks of loaded L2 latency).
constexpr int kS3Ring = DMC_S3_RING;
// 1: an epilogue warp pulls its whole share of the accumulators into registers and hands the TMEM buffer back before any
// arithmetic.  Measured (profiles/chain_early_release_experiment_r02.txt): the MMA warp no longer waits for buffers, but
// the 128 pinned registers cost the activation code its interleaving -- epilogue 4 800 -> 5 800 clocks per chunk-add
// tile, DepthConvBlock-256 148 -> 155 us.  Off: the buffer is released after the last chunk's loads.
#ifndef DMC_S3_EARLY_RELEASE
#define DMC_S3_EARLY_RELEASE 0
#endif
constexpr bool kS3EarlyRelease = DMC_S3_EARLY_RELEASE != 0;
// 1: the A and W loads of an operand stage are issued by two warps (see the W producer in the kernel)
#ifndef DMC_S3_SPLIT_PRODUCER
#define DMC_S3_SPLIT_PRODUCER 1
#endif
constexpr bool kS3SplitProducer = DMC_S3_SPLIT_PRODUCER != 0;
constexpr int kS3WarpSmem = kS3Ring * kS3ChunkBytes;
constexpr int kS3BarBytes = 1024;
// The bias of a tile, per epilogue warp: the <= 64 accumulator columns the warp works on, double buffered (the next
// tile's values are requested a tile ahead).  Read straight from global memory per chunk, an L1 miss under the
// operand traffic sat in front of the first add of the chunk (long-scoreboard stalls on the bias FADDs: 4 % of the
// kernel's samples, profiles/chain_dcb_r02_stall_summary.txt).
constexpr int kS3BiasCols = 64;
constexpr int kS3BiasBytes = kS3EpiWarps * 2 * kS3BiasCols * 4;

// Shared-memory descriptor (cute::UMMA::SmemDescriptor) of a K-major SWIZZLE_32B operand tile = rows of 16 fp16
// (32 B), which is what the tile-blocked activations and weights are in shared memory: start>>4 [0,14) | LBO (unused)
// = 1 [16,30) | SBO = 8 rows x 32 B = 256 -> 16 [32,46) | version 1 [46,48) | layout SWIZZLE_32B = 6 [61,64)
__device__ __forceinline__ uint64_t s3_desc32(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = 16u | (1u << 14) | (6u << 29);
  return ((uint64_t)hi << 32) | lo;
}

// ... and from its low word (start >> 4 | LBO), which is what changes from operand to operand
__device__ __forceinline__ uint64_t s3_desc_lo(uint32_t lo) {
  return ((uint64_t)(16u | (1u << 14) | (6u << 29)) << 32) | lo;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// orders this thread's generic-proxy global accesses (the flag acquire / release) with its async-proxy
// ones (TMA loads / stores of the tensors the flag guards)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

#ifdef DMC_EPI_TIMING
// phase clocks of epilogue warp 0 of CTA 0, summed over chunks: [0] ring-slot wait + residual prefetch + publish,
// [1] accumulator loads (tcgen05.ld + bias), [2] combine / bias / activation, [3] residual wait + join,
// [4] scale + split + st.shared + fence, [5] TMA store issue, [6] chunks, [7] wait for the accumulator (per tile)
__device__ unsigned long long g_epi_t[8];
// tile timeline of cluster 0 (CTA 0): per tile [0] MMA warp starts waiting for the TMEM buffer, [1] has it,
// [2] first operand stage arrived, [3] last MMA issued, [4] epilogue warp 0 starts waiting for the accumulator,
// [5] has it, [6] done with the tile, [7] table entry
__device__ unsigned long long g_tile_trace[256][14];   // [8] clocks the MMA warp waited for operand stages, [9] clocks in MMA issue
// [CTA 0/1][epilogue warp 0 / 7][tile][got accumulator, released TMEM, done]
__device__ unsigned long long g_warp_trace[2][2][256][3];
#define WARP_T(t, i) do { if (blockIdx.x < 2 && x.lane == 0 && (x.ew == 0 || x.ew == 7) && (t) < 256u) \
  g_warp_trace[blockIdx.x][x.ew == 7][(t)][(i)] = (unsigned long long)clock64(); } while (0)
#define TILE_T(t, i, v) do { if (blockIdx.x == 0 && lane == 0 && (t) < 256u) g_tile_trace[(t)][(i)] = (unsigned long long)(v); } while (0)
// (the phase clocks slow the timed warp by ~20 %: -DDMC_EPI_PHASES=0 keeps only the tile / warp traces)
#ifndef DMC_EPI_PHASES
#define DMC_EPI_PHASES 1
#endif
#define EPI_T(i) do { if (DMC_EPI_PHASES && x.timed) { const long long t_ = clock64(); x.tacc[i] += t_ - x.tlast; x.tlast = t_; } } while (0)
#else
#define EPI_T(i) do { } while (0)
#define TILE_T(t, i, v) do { } while (0)
#define WARP_T(t, i) do { } while (0)
#endif

// Per-warp epilogue state that lives across tiles and layers.
struct EpiCtx {
#ifdef DMC_EPI_TIMING
  bool timed;
  int ew;
  uint32_t tile;
  long long tlast;
  long long tacc[8];
#endif
  uint32_t ringBuf;          // this warp's staging ring
  uint32_t barRes;           // its kS3Ring residual barriers (8 B apart)
  uint32_t slot;             // ring position of the chunk being processed (advances with every stored chunk)
  uint32_t res_phase;        // bit s: parity the next residual wait on ring barrier s has to see
  int lane, quad, part;      // part: which share of the tile's columns (0 .. kS3Split-1)
  uint32_t rank;
  uint32_t depsOk;           // shared-memory word: tiles of this CTA whose dependencies the producer warp has seen met
  uint32_t biasBuf;          // shared-memory copy of the bias of this warp's accumulator columns of the current tile
  int biasCol0;              // first accumulator column (within the tile) it holds
  bool epi_mem;
};

// A tile has BN/kS3ChunkCols output chunks (PAIR: a 64-column group of accumulators = 32 values + their 32 chunk-add
// partners gives 32 outputs = 32/kS3ChunkCols chunks).  The warps of a quadrant take contiguous shares of `per` chunks.
__device__ __forceinline__ int s3_total(int kind, int BN) {
  return kind == S3_PAIR ? (BN >> 6) * kS3PairUnits : BN / kS3ChunkCols;
}
__device__ __forceinline__ int s3_per(int kind, int BN) { return (s3_total(kind, BN) + kS3Split - 1) / kS3Split; }
__device__ __forceinline__ int s3_nchunk(int kind, int BN, int part) {
  const int per = s3_per(kind, BN);
  return max(0, min(per, s3_total(kind, BN) - part * per));
}
// accumulator column (PAIR: of the value half; partners sit 32 columns further) of this warp's chunk c
__device__ __forceinline__ int s3_acc_col(int kind, int BN, int part, int c) {
  const int j = part * s3_per(kind, BN) + c;
  return kind == S3_PAIR ? (j / kS3PairUnits) * 64 + (j % kS3PairUnits) * kS3ChunkCols : j * kS3ChunkCols;
}
// destination column of chunk c of N tile nt for this warp
__device__ __forceinline__ int s3_dest_col(int kind, int BN, int part, int nt, int c) {
  const int n_idx = nt * BN;
  const int j = part * s3_per(kind, BN) + c;
  return (kind == S3_PAIR ? (n_idx >> 1) : n_idx) + j * kS3ChunkCols;
}

// One tile of one layer, one epilogue warp.  `next_res(c)` is called once per chunk (after the staging
// ring slot two chunks back is known to be free) so the caller can prefetch the residual of the item
// that follows chunk c.
// The fields of a layer the epilogue needs, read ONCE per tile into registers (the layer record sits in the
// kernel parameters under a run-time index: every access is an indexed constant load).
struct StageRegs {
  const float* bias;
  const float* scale;
  const CUtensorMap* tmOut;
  const CUtensorMap* tmRes;
  const CUtensorMap* tmRes2;
  int BN, n_out, kind;
  uint32_t need;
  bool two_acc;
  float comp;
};
__device__ __forceinline__ StageRegs s3_load_stage(const S3StageDev& S) {
  StageRegs r;
  r.bias = S.bias; r.scale = S.scale; r.tmOut = &S.tmOut; r.tmRes = &S.tmRes; r.tmRes2 = &S.tmRes2;
  r.BN = S.BN; r.n_out = S.n_out; r.kind = S.kind; r.need = S.need; r.two_acc = S.nterms != 1;
  r.comp = S.comp;
  return r;
}

// One body for all layer kinds (warp-uniform run-time branches, taken once per 32-column chunk): six inlined
// template instances of it made ptxas spill ~5 KB per thread, each instance alone compiles without a spill.
template <class NextRes, class Release>
__device__ __forceinline__ void s3_epilogue_tile(const StageRegs& S, EpiCtx& x, int mt, int nt, uint32_t taddr,
                                                 int* err, NextRes next_res, Release release_tmem) {
  using namespace s3;
  const int lane = x.lane;
  const bool kF32 = S.kind == S3_F32 || S.kind == S3_F32_WSILU;
  const bool kRes = S.kind == S3_RES;
  const bool kRes2 = S.kind == S3_RES2;     // shortcut blocks: (acc + res1) + res2, layers.py:75-76
  const bool kWsilu = S.kind == S3_WSILU || S.kind == S3_PAIR || S.kind == S3_F32_WSILU;
  const bool kPair = S.kind == S3_PAIR;
  const int kKind = kPair ? S3_PAIR : S3_PLAIN;
  const int nchunk = s3_nchunk(kKind, S.BN, x.part);
  const int row0 = mt * 256 + (int)x.rank * 128 + x.quad * 32;
  const uint32_t swz = (uint32_t)((lane >> 2) & 1) << 4;        // SWIZZLE_32B: 16-byte unit ^= row bit 2
  const uint32_t rowOff = (uint32_t)lane * 32u;
  const bool two_acc = S.two_acc;
  const float comp = S.comp;
  // ---- phase 1: this warp's whole share of the accumulators leaves TMEM at once (at most two units of kS3ChunkCols
  // columns x two accumulators: the two chunks of a plain tile, or the value and partner halves of a chunk-add chunk)
  // and the buffer goes back to the MMA warp BEFORE any arithmetic.  With the release after the last chunk's loads
  // (i.e. after the whole first chunk: ~3 400 clocks into the tile) the two TMEM buffers tied the MMA side and the
  // epilogue side together: period = (time to release + time to refill) / 2 instead of max(MMA, epilogue).
  uint32_t ua[2][kS3ChunkCols], ub[2][kS3ChunkCols];
#pragma unroll
  for (int u = 0; u < (kS3EarlyRelease ? 2 : 0); ++u) {
    const int cu = kPair ? 0 : u;
    const bool uvalid = cu < nchunk && s3_dest_col(kKind, S.BN, x.part, nt, cu) < S.n_out;   // warp-uniform
    const int col = s3_acc_col(kKind, S.BN, x.part, cu) + (kPair ? 32 * u : 0);
    if (uvalid) {
      if (kS3ChunkCols == 32) {
        tc_ld32(taddr + col, ua[u]);
        if (two_acc) tc_ld32(taddr + 128u + col, ub[u]);
      } else {
        tc_ld16(taddr + col, ua[u]);
        if (two_acc) tc_ld16(taddr + 128u + col, ub[u]);
      }
    }
  }
  if (kS3EarlyRelease) {
    tc_wait_ld();
    release_tmem();
#ifdef DMC_EPI_TIMING
    WARP_T(x.tile, 1);
#endif
    EPI_T(1);
  }
  // ---- phase 2: chunk by chunk (from registers, or loading as it goes).  Unrolled only when the accumulators sit in
  // registers (constant indices): a second copy of this body raised the kernel's instruction-fetch stalls from 4 % to
  // 10 % of its samples (stall_no_inst, profiles/chain_dcb_r02b_stall_summary.txt).
#pragma unroll(kS3EarlyRelease ? 2 : 1)
  for (int c = 0; c < 2; ++c) {
    if (c >= nchunk) break;
    const int acol = s3_acc_col(kKind, S.BN, x.part, c);         // accumulator column of this chunk
    const int dcol = s3_dest_col(kKind, S.BN, x.part, nt, c);
    const bool valid = dcol < S.n_out;                           // warp-uniform
    EPI_T(7);
    // Layers with a residual: the tile of chunk c-1 (its store was issued a moment ago) receives the residual of chunk
    // c+1 now -- it has to arrive a whole chunk ahead (a loaded L2 answers after ~3 000 clocks), so that store's read of
    // shared memory is waited for here.  Layers without one only need the tile of chunk c-2 back, and not before their
    // own results are written: that wait (further down) finds the store long done.
    if (kRes || kRes2) {
      if (lane == 0) {
        if (kS3Ring >= 3) tma_store_wait_read1(); else tma_store_wait_read0();
      }
      __syncwarp();
    }
    next_res(c);
    if (kRes2 && valid && x.epi_mem) {
      // both staging tiles are this chunk's: the first residual lands where the result will be written, the second
      // in the other tile (every earlier store has read its data: wait_group.read 0 above)
      const uint32_t other = x.slot + 1 == kS3Ring ? 0u : x.slot + 1;
      if (lane == 0) {
        if (kS3Ring >= 3) tma_store_wait_read0();
        // (the first residual may come from an earlier layer of this chain: the producer warp acquired its
        // completion before this tile's MMAs could start -- pick that up, then order the TMA reads after it)
        (void)ld_acquire_cta_shared(x.depsOk);
        fence_proxy_async_global();
        mbar_expect_tx(x.barRes + 8u * x.slot, kS3ChunkBytes);
        tma_load_local(x.ringBuf + x.slot * kS3ChunkBytes, S.tmRes, dcol, row0, x.barRes + 8u * x.slot);
        mbar_expect_tx(x.barRes + 8u * other, kS3ChunkBytes);
        tma_load_local(x.ringBuf + other * kS3ChunkBytes, S.tmRes2, dcol, row0, x.barRes + 8u * other);
      }
      __syncwarp();
    }
    EPI_T(0);
    float v[kS3ChunkCols];
    if (valid) {
      // v (+)= act(main + small * 2^-11 + bias) for the accumulator columns [col, col + kS3ChunkCols) held in a / b
      auto act = [&](uint32_t* a, uint32_t* b, int col, bool accumulate) {
        float w[kS3ChunkCols];
        float4 bq[kS3ChunkCols / 4];
#pragma unroll
        for (int i = 0; i < kS3ChunkCols / 4; ++i) {
          const uint4 q = ld_shared_v4(x.biasBuf + (uint32_t)(col - x.biasCol0) * 4u + 16u * i);
          bq[i] = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
        }
        if (!kS3EarlyRelease) {
          if (kS3ChunkCols == 32) {
            tc_ld32(taddr + col, a);
            if (two_acc) tc_ld32(taddr + 128u + col, b);
          } else {
            tc_ld16(taddr + col, a);
            if (two_acc) tc_ld16(taddr + 128u + col, b);
          }
          tc_wait_ld();
          EPI_T(1);
        }
#pragma unroll
        for (int i = 0; i < kS3ChunkCols / 4; ++i) {
          const float4 q = bq[i];
          const float bias4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float t = __uint_as_float(a[4 * i + k]);
            // main + small' * 2^-11 with one rounding; small' = small + main * comp puts back what the tensor core's
            // truncating accumulation of the main term lost on average (comp is 2^11-scaled like `small`)
            if (two_acc) t = fmaf(fmaf(t, comp, __uint_as_float(b[4 * i + k])), kLoInv, t);
            w[4 * i + k] = add_rn(t, bias4[k]);
          }
        }
        if (kWsilu) {
#pragma unroll
          for (int i = 0; i < kS3ChunkCols / 2; ++i) wsilu2_fast(w[2 * i], w[2 * i + 1]);
        }
        if (accumulate) {
#pragma unroll
          for (int i = 0; i < kS3ChunkCols; ++i) v[i] = add_rn(v[i], w[i]);
        } else {
#pragma unroll
          for (int i = 0; i < kS3ChunkCols; ++i) v[i] = w[i];
        }
      };
      // (three inlined copies of the activation code.  ONE copy in a two-trip loop for the value / partner halves was
      // tried for the instruction cache's sake: chunk-add epilogue 53.7 -> 59.2 us, the loop-carried v[] costs more.)
      if (kPair) {
        act(ua[0], ub[0], acol, false);
        act(ua[kS3EarlyRelease ? 1 : 0], ub[kS3EarlyRelease ? 1 : 0], acol + 32, true);
      } else {
        act(ua[kS3EarlyRelease ? c : 0], ub[kS3EarlyRelease ? c : 0], acol, false);
      }
    }
    if (!kS3EarlyRelease && c == nchunk - 1) {
      release_tmem();      // accumulator fully read: hand the TMEM buffer back
#ifdef DMC_EPI_TIMING
      WARP_T(x.tile, 1);
#endif
    }
    EPI_T(2);
    if (!valid) continue;
    const uint32_t tileBuf = x.ringBuf + x.slot * kS3ChunkBytes;
    // v += hi + lo * 2^-11 of the residual tile in staging tile `slot` (exactly join2)
    auto add_residual = [&](uint32_t slot) {
      mbar_wait(x.barRes + 8u * slot, (x.res_phase >> slot) & 1u, err, 5);
      x.res_phase ^= 1u << slot;
      const uint32_t src = x.ringBuf + slot * kS3ChunkBytes + rowOff;
#pragma unroll
      for (int blk = 0; blk < kS3ChunkBlocks; ++blk) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint4 qh = ld_shared_v4(src + blk * 1024 + (((uint32_t)hf << 4) ^ swz));
          const uint4 ql = ld_shared_v4(src + kS3ChunkBlocks * 1024 + blk * 1024 + (((uint32_t)hf << 4) ^ swz));
          const uint32_t uh[4] = {qh.x, qh.y, qh.z, qh.w};
          const uint32_t ul[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 16 * blk + 8 * hf + 2 * k;
            v[i] = add_rn(v[i], join2(h2lo(uh[k]), h2lo(ul[k])));
            v[i + 1] = add_rn(v[i + 1], join2(h2hi(uh[k]), h2hi(ul[k])));
          }
        }
      }
    };
    if ((kRes || kRes2) && x.epi_mem) add_residual(x.slot);
    if (kRes2 && x.epi_mem) add_residual(x.slot + 1 == kS3Ring ? 0u : x.slot + 1);
    EPI_T(3);
    if (!(kRes || kRes2)) {
      // this chunk's staging tile was last read by the store of chunk c-2: all but the newest store have read their data
      if (lane == 0) {
        if (kS3Ring >= 3) tma_store_wait_read2(); else tma_store_wait_read1();
      }
      __syncwarp();
    }
    if (S.scale) {
      const float4* sp = reinterpret_cast<const float4*>(S.scale + dcol);
#pragma unroll
      for (int i = 0; i < kS3ChunkCols / 4; ++i) {
        float4 q = make_float4(1.f, 1.f, 1.f, 1.f);
        if (dcol + 4 * i < S.n_out) q = __ldg(sp + i);
        v[4 * i] = mul_rn(v[4 * i], q.x); v[4 * i + 1] = mul_rn(v[4 * i + 1], q.y);
        v[4 * i + 2] = mul_rn(v[4 * i + 2], q.z); v[4 * i + 3] = mul_rn(v[4 * i + 3], q.w);
      }
    }
    if (kF32) {
      // fp32 rows: 4 * kS3ChunkCols bytes per lane; 16-byte unit u of row r sits at u ^ (r & 7) (128-byte rows,
      // SWIZZLE_128B) or at u ^ ((r >> 1) & 3) (64-byte rows, SWIZZLE_64B)
      const uint32_t dst = tileBuf + (uint32_t)lane * (4u * kS3ChunkCols);
      const uint32_t swf = kS3ChunkCols == 32 ? (uint32_t)(lane & 7) << 4 : (uint32_t)((lane >> 1) & 3) << 4;
#pragma unroll
      for (int u = 0; u < kS3ChunkCols / 4; ++u)
        st_shared_v4(dst + (((uint32_t)u << 4) ^ swf), __float_as_uint(v[4 * u]), __float_as_uint(v[4 * u + 1]),
                     __float_as_uint(v[4 * u + 2]), __float_as_uint(v[4 * u + 3]));
      fence_async_smem();
      __syncwarp();
      EPI_T(4);
      if (lane == 0 && x.epi_mem) tma_store_2d(S.tmOut, tileBuf, dcol, row0);
    } else {
      // hi / 2^11-scaled lo split, two elements per conversion
      const uint32_t dst = tileBuf + rowOff;
#pragma unroll
      for (int blk = 0; blk < kS3ChunkBlocks; ++blk) {
        uint32_t ph_[8], pl_[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split2x2(v[16 * blk + 2 * i], v[16 * blk + 2 * i + 1], ph_[i], pl_[i]);
        st_shared_v4(dst + blk * 1024 + swz, ph_[0], ph_[1], ph_[2], ph_[3]);
        st_shared_v4(dst + blk * 1024 + (16u ^ swz), ph_[4], ph_[5], ph_[6], ph_[7]);
        st_shared_v4(dst + kS3ChunkBlocks * 1024 + blk * 1024 + swz, pl_[0], pl_[1], pl_[2], pl_[3]);
        st_shared_v4(dst + kS3ChunkBlocks * 1024 + blk * 1024 + (16u ^ swz), pl_[4], pl_[5], pl_[6], pl_[7]);
      }
      fence_async_smem();
      __syncwarp();
      EPI_T(4);
      if (lane == 0 && x.epi_mem) tma_store(S.tmOut, tileBuf, dcol, row0);
    }
    if (++x.slot == kS3Ring) x.slot = 0;
    EPI_T(5);
#ifdef DMC_EPI_TIMING
    x.tacc[6] += 1;
#endif
  }
}

__global__ void __launch_bounds__(kS3Threads, 1)
k_gemm_s3_chain(const __grid_constant__ S3ChainParams p) {
  using namespace s3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem_raw);
  // cluster = (2,1,1) or (4,1,1): rank in the CTA pair / which pair of the cluster (provably warp-uniform)
  const uint32_t rank = blockIdx.x & 1u;
  const uint32_t pairIdx = p.cl4 ? ((blockIdx.x >> 1) & 1u) : 0u;
  const uint32_t leaderRank = pairIdx << 1;          // cluster rank of this pair's leader CTA
  const uint16_t pairMask = (uint16_t)(3u << leaderRank);
  const uint16_t clusterMask = p.cl4 ? (uint16_t)0xF : (uint16_t)0x3;
  const bool leader = rank == 0;
  const uint32_t stageBytes = p.stageBytes;
  const uint32_t epiBase = base + (uint32_t)p.stages * stageBytes;
  const uint32_t barBase = epiBase + kS3EpiWarps * kS3WarpSmem;
  // barriers: full[8] | empty[8] | tfull[2] | tempty[2] | res[8 warps][3 (+1 pad)] | tmem slot | deps_ok
  auto bar_full = [&](int s) { return barBase + 8u * s; };
  auto bar_empty = [&](int s) { return barBase + 64u + 8u * s; };
  auto bar_tfull = [&](int b) { return barBase + 128u + 8u * b; };
  auto bar_tempty = [&](int b) { return barBase + 144u + 8u * b; };
  auto bar_res = [&](int w, int b) { return barBase + 160u + 32u * w + 8u * b; };
  const uint32_t tmemSlot = barBase + 160u + 32u * kS3EpiWarps;
  const uint32_t depsOk = tmemSlot + 4u;             // number of this CTA's tiles whose dependencies are met
  const uint32_t warpsDone = tmemSlot + 8u;          // [8] epilogue warps of this CTA that finished tile (t & 7)
  const uint32_t biasBase = barBase + kS3BarBytes;   // [epilogue warp][2][kS3BiasCols] fp32

  if (warp == 0 && lane == 0) {
    if (base & 1023u) {
      mbar_timeout(p.err, 9);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full(s), (p.dbg & 1) ? 2 : 1);   // probe mode without loads: one plain arrival per CTA
      mbar_init(bar_empty(s), p.cl4 ? 2 : 1);        // every pair that writes into this stage has consumed it
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull(b), 1);
      mbar_init(bar_tempty(b), kS3EpiWarps * 2);
    }
    for (int w = 0; w < kS3EpiWarps; ++w)
      for (int b = 0; b < kS3Ring; ++b) mbar_init(bar_res(w, b), 1);
    *reinterpret_cast<volatile uint32_t*>(smem_raw + (depsOk - base)) = 0u;
    for (int i = 0; i < 8; ++i) *reinterpret_cast<volatile uint32_t*>(smem_raw + (warpsDone - base) + 4 * i) = 0u;
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
  pdl_prologue_done();                     // everything above overlapped the previous kernel's tail
  // (read where it is used: kept live across the role branches it ended up in local memory, see my_tiles below)
  auto tmem_base_ld = [&]() { return *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmemSlot - base)); };
  const int unit = (int)(p.cl4 ? blockIdx.x >> 2 : blockIdx.x >> 1);
  const int units = (int)(p.cl4 ? gridDim.x >> 2 : gridDim.x >> 1);
  // this cluster's entries are table[unit + t * units], t = 0 .. my_tiles-1.  The role loops below count t only: with
  // an entry index carried as a loop variable ptxas kept it in LOCAL memory, and every publication's
  // fence.proxy.async.global (CCTL.IVALL: the whole L1 is invalidated) turned the reload at the top of the next tile into
  // an L2 round trip -- ~770 clocks between two tiles of an epilogue warp, ~300 in the MMA warp.
  const uint32_t my_tiles = unit < p.n_entries ? (uint32_t)((p.n_entries - unit + units - 1) / units) : 0u;
  const uint32_t* const my_tab = p.table + unit;
  auto tab_at = [&](uint32_t t) { return __ldg(my_tab + (size_t)t * (uint32_t)units); };

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kS3EpiWarps == 16 ? 64 : 72) : "memory");
  if (warp == 0) {
    // ------------------------------------------------------------ operand producer (both CTAs)
    // (whole warp walks the loop, one elected lane issues: addresses stay in uniform registers)
    int s = 0;
    uint32_t ph = 0, tcount = 0;
    const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(p.epoch_ctr) + 1u;
    // what does not change from stage to stage is worked out once: the issuing lane, the leader's barrier of stage 0 as
    // a cluster address (the others follow at 8-byte steps), whether this launch runs the plain path at all
    const bool sel = elect_one();
    const uint32_t lbar0 = mapa(bar_full(0), leaderRank);
    const bool plain_path = !p.cl4 && !(p.dbg & 1);
    // (the table entry of the next tile is fetched a tile ahead: its ~700 clocks of global latency sat on the
    // critical path of every tile in this warp and in the MMA warp)
    uint32_t e_nxt = my_tiles ? tab_at(0) : 0u;
    for (; tcount < my_tiles; ++tcount) {
      const uint32_t e = e_nxt;
      if (tcount + 1 < my_tiles) e_nxt = tab_at(tcount + 1);
      const int l = (int)(e >> 28), mt = (int)(e & 0xfffffu);
      const int nt = p.cl4 ? 2 * (int)((e >> 20) & 0xffu) + (int)pairIdx : (int)((e >> 20) & 0xffu);
      const S3StageDev& S = p.st[l];
      if (l > 0 && S.need != 0u) {
        // all N tiles of the previous layer for these rows must be stored
        const uint32_t* flag = p.done + (size_t)(l - 1) * p.MT + mt;
        const uint32_t target = epoch * S.need;
        if ((int)(ld_acquire_gpu(flag) - target) < 0) {
          const long long t0 = clock64();
          while ((int)(ld_acquire_gpu(flag) - target) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > 6000000000LL) mbar_timeout(p.err, 6);
          }
        }
        fence_proxy_async_global();          // the TMA reads below are ordered after the acquire
      }
      if (lane == 0) st_release_cta_shared(depsOk, tcount + 1);
      const uint32_t wRows = (uint32_t)S.BN >> 1;
      // both planes x 32 k, or (single term) the hi plane x 64 k: the same bytes
      const int kblk = S.kblk;
      // (probe switch 32: no W loads -- what a W tile kept resident across row tiles would leave of the load time)
      const uint32_t tx = (S.nterms == 1 ? (uint32_t)kblk : 4u) * (kS3APlane + ((p.dbg & 32) ? 0u : wRows * (kS3BK * 2)));
      const int m_idx = mt * 256 + (int)rank * 128;
      const int n_idx = nt * S.BN + (int)(rank * wRows);
      if (plain_path) {
        // The lean loop: per stage one barrier test, (leader) one expect_tx, one TMA issue -- nothing read from the
        // kernel parameters, no branch on the launch's modes.  The general loop below spent ~350 clocks per stage on
        // exactly that (indexed constant loads of the layer record, mode tests, mapa, elect and reconvergence per
        // stage): with NO loads and NO MMAs the kernel still needed 3 900 clocks per tile (tools/epi_probe.py,
        // "hand-over only"), the floor under everything else.
        const CUtensorMap* tmA = &S.tmA;
        const int k_blocks = S.k_blocks, m16 = m_idx >> 4;
        const CUtensorMap* tmWl = (!kS3SplitProducer && !(p.dbg & 32)) ? &S.tmW : nullptr;
        const int n16 = n_idx >> 4;
#ifdef DMC_EPI_TIMING
        long long tp_wait = 0, tp_issue = 0;
#endif
        for (int kb = 0; kb < k_blocks; ++kb) {
#ifdef DMC_EPI_TIMING
          const long long tp0 = clock64();
#endif
          mbar_wait(bar_empty(s), ph ^ 1, p.err, 1);
#ifdef DMC_EPI_TIMING
          const long long tp1 = clock64();
          tp_wait += tp1 - tp0;
#endif
          if (sel) {
            const uint32_t sa = base + s * stageBytes;
            if (leader) mbar_expect_tx(bar_full(s), tx);
            tma_load_pair(sa, tmA, m16, kb * kblk, lbar0 + 8u * s);
            if (tmWl) tma_load_pair_w(sa + kPlanes * kS3APlane, tmWl, n16, kb * kblk, lbar0 + 8u * s);
          }
#ifdef DMC_EPI_TIMING
          tp_issue += clock64() - tp1;
#endif
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        TILE_T(tcount, 11, tp_wait);
        TILE_T(tcount, 12, tp_issue);
        continue;
      }
#ifdef DMC_EPI_TIMING
      long long tp_wait = 0, tp_issue = 0;
#endif
      for (int kb = 0; kb < S.k_blocks; ++kb) {
#ifdef DMC_EPI_TIMING
        const long long tp0 = clock64();
#endif
        mbar_wait(bar_empty(s), ph ^ 1, p.err, 1);
#ifdef DMC_EPI_TIMING
        const long long tp1 = clock64();
        tp_wait += tp1 - tp0;
#endif
        const uint32_t sa = base + s * stageBytes;
        const uint32_t sw = sa + kPlanes * kS3APlane;
        if (elect_one()) {
          const uint32_t lbar = mapa(bar_full(s), leaderRank);
          if (p.dbg & 1) {
            if (leader) mbar_arrive(bar_full(s));
            else mbar_arrive_cluster(lbar);
          } else {
            // both CTAs' bytes complete on the LEADER's barrier, which the leader arms for both
            if (leader) mbar_expect_tx(bar_full(s), tx);
            if (p.cl4) {
              // the two pairs work on the same rows: this CTA fetches 64 of its 128 rows and multicasts them to
              // the CTA of the same rank in the other pair (which sends the other 64)
              const uint16_t mc = (uint16_t)((1u << rank) | (4u << rank));
              const int planes = S.nterms == 1 ? 1 : kPlanes;
              for (int pl = 0; pl < planes; ++pl)
                for (int blk = 0; blk < kblk; ++blk)
                  tma_load_pair_mcast(sa + pl * kS3APlane + blk * 4096u + pairIdx * 2048u, &S.tmA64,
                                      (m_idx + (int)pairIdx * 64) >> 4, kb * kblk + blk, pl, lbar, mc);
            } else {
              tma_load_pair(sa, &S.tmA, m_idx >> 4, kb * kblk, lbar);
            }
            if (!kS3SplitProducer && !(p.dbg & 32)) tma_load_pair_w(sw, &S.tmW, n_idx >> 4, kb * kblk, lbar);
          }
        }
        __syncwarp();
#ifdef DMC_EPI_TIMING
        tp_issue += clock64() - tp1;
#endif
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      TILE_T(tcount, 11, tp_wait);
      TILE_T(tcount, 12, tp_issue);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp walks the loop so that every address below is warp-uniform (uniform registers,
    // no per-MMA register -> uniform-register shuffling); one elected lane issues.
    if (leader) {
      const uint32_t aStep = kS3APlane >> 4;
      const bool sel = elect_one();                               // the lane that issues
      const uint32_t alo0 = ((base >> 4) & 0x3FFFu) | (1u << 16);  // low descriptor word of stage 0 (s3_desc32)
      const uint32_t stageLo = stageBytes >> 4;
      const bool do_mma = !(p.dbg & 2), twice = (p.dbg & 64) != 0;
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      uint32_t e_nxt = my_tiles ? tab_at(0) : 0u;
      // the fields of a layer, refreshed when the layer changes (indexed constant loads: ~300 clocks between the last
      // MMA of a tile and the first of the next when they were read per tile)
      uint32_t l_c = 0xffffffffu, idesc = 0, wStep = 0, wKs = 0;
      int k_blocks = 0, kblk = 0;
      bool split = false;
      for (; tcount < my_tiles; ++tcount) {
        const uint32_t e = e_nxt;
        if (tcount + 1 < my_tiles) e_nxt = tab_at(tcount + 1);
        if ((e >> 28) != l_c) {
          l_c = e >> 28;
          const S3StageDev& S = p.st[l_c];
          // instruction descriptor: D=f32 [4,6)=1, A=f16 [7,10)=0, B=f16 [10,13)=0, K-major both,
          // N>>3 at [17,23), M>>4 at [24,29) with M = 256 for the pair
          idesc = (1u << 4) | ((uint32_t)(S.BN >> 3) << 17) | (16u << 24);
          wStep = ((uint32_t)(S.BN >> 1) * (kS3BK * 2)) >> 4;   // plane stride of the W stage
          wKs = ((uint32_t)(S.BN >> 1) * 32u) >> 4;             // second 16-wide k block of a plane
          k_blocks = S.k_blocks;
          split = S.nterms != 1;
          kblk = S.kblk;
        }
        const uint32_t buf = tcount & 1;
        TILE_T(tcount, 0, clock64());
        TILE_T(tcount, 7, e);
        mbar_wait(bar_tempty(buf), ((tcount >> 1) & 1) ^ 1, p.err, 2);
        TILE_T(tcount, 1, clock64());
        tc_fence_after();
        const uint32_t d_main = tmem_base_ld() + buf * 256u;
        const uint32_t d_small = d_main + 128u;
#ifdef DMC_EPI_TIMING
        long long tw_full = 0, tw_issue = 0, tw_pre = 0, tw_mma = 0, tw_commit = 0;
        (void)tw_pre; (void)tw_mma; (void)tw_commit;
#endif
        // The tensor pipe takes one MMA per 64 clocks and queues hardly any: whatever this warp does between the last
        // MMA of a stage and the first of the next is idle time of the pipe (tile trace of the one-stage-per-iteration
        // loop: ~390 clocks of barrier wait / descriptor arithmetic / commit / reconvergence / loop around 360 clocks of
        // MMA issue per stage: tensor pipe active 46 %).  So TWO operand stages per trip through the loop (one elected
        // section, one reconvergence), descriptors as 32-bit words stepped by adds, the second stage's barrier
        // already tested when the first one's MMAs go out.
        for (int kb = 0; kb < k_blocks; kb += 2) {
          const bool two = kb + 1 < k_blocks;                       // (an odd count leaves a single stage at the end)
          int s1 = s + 1;
          uint32_t ph1 = ph;
          if (s1 == p.stages) { s1 = 0; ph1 ^= 1; }
#ifdef DMC_EPI_TIMING
          const long long tq0 = clock64();
#endif
          mbar_wait(bar_full(s), ph, p.err, 3);
          if (two) mbar_wait(bar_full(s1), ph1, p.err, 3);
#ifdef DMC_EPI_TIMING
          const long long tq1 = clock64();
          tw_full += tq1 - tq0;
#endif
          if (kb == 0) TILE_T(tcount, 2, clock64());
          tc_fence_after();
          const uint32_t alo = alo0 + (uint32_t)s * stageLo;       // low descriptor word: hi plane, first k block
          const uint32_t wlo = alo + ((kPlanes * kS3APlane) >> 4);
          const uint32_t alo1 = alo0 + (uint32_t)s1 * stageLo;
          const uint32_t wlo1 = alo1 + ((kPlanes * kS3APlane) >> 4);
          const uint32_t acc0 = kb == 0 ? 0u : 1u;
          const bool last = kb + 2 >= k_blocks;
          if (sel) {
            // hi*hi -> main accumulator; the two 2^11-scaled cross terms hi*lo', lo'*hi -> second one
            // (a plane of the A stage is 8 KB = 512 descriptor units, a 16-wide k block 4 KB = 256; a single-term
            // stage holds kblk = 2 or 4 k blocks of the hi plane)
            auto stage_mmas = [&](uint32_t a, uint32_t w, uint32_t acc) {
              if (!do_mma) return;
              if (split) {
                tc_mma_pair(d_main, s3_desc_lo(a), s3_desc_lo(w), idesc, acc);
                tc_mma_pair(d_small, s3_desc_lo(a), s3_desc_lo(w + wStep), idesc, acc);
                tc_mma_pair(d_small, s3_desc_lo(a + aStep), s3_desc_lo(w), idesc, 1u);
                tc_mma_pair(d_main, s3_desc_lo(a + 256u), s3_desc_lo(w + wKs), idesc, 1u);
                tc_mma_pair(d_small, s3_desc_lo(a + 256u), s3_desc_lo(w + wKs + wStep), idesc, 1u);
                tc_mma_pair(d_small, s3_desc_lo(a + 256u + aStep), s3_desc_lo(w + wKs), idesc, 1u);
                if (twice) {           // probe: every MMA twice (is a stage bound by the tensor pipe or by its barriers?)
                  tc_mma_pair(d_main, s3_desc_lo(a), s3_desc_lo(w), idesc, 1u);
                  tc_mma_pair(d_small, s3_desc_lo(a), s3_desc_lo(w + wStep), idesc, 1u);
                  tc_mma_pair(d_small, s3_desc_lo(a + aStep), s3_desc_lo(w), idesc, 1u);
                  tc_mma_pair(d_main, s3_desc_lo(a + 256u), s3_desc_lo(w + wKs), idesc, 1u);
                  tc_mma_pair(d_small, s3_desc_lo(a + 256u), s3_desc_lo(w + wKs + wStep), idesc, 1u);
                  tc_mma_pair(d_small, s3_desc_lo(a + 256u + aStep), s3_desc_lo(w + wKs), idesc, 1u);
                }
              } else {
                tc_mma_pair(d_main, s3_desc_lo(a), s3_desc_lo(w), idesc, acc);
                tc_mma_pair(d_main, s3_desc_lo(a + 256u), s3_desc_lo(w + wKs), idesc, 1u);
                if (kblk == 4) {
                  tc_mma_pair(d_main, s3_desc_lo(a + 512u), s3_desc_lo(w + 2u * wKs), idesc, 1u);
                  tc_mma_pair(d_main, s3_desc_lo(a + 768u), s3_desc_lo(w + 3u * wKs), idesc, 1u);
                }
              }
            };
            stage_mmas(alo, wlo, acc0);
            tc_commit_pair(bar_empty(s), clusterMask);
            if (two) {
              stage_mmas(alo1, wlo1, 1u);
              tc_commit_pair(bar_empty(s1), clusterMask);
            }
            if (last) tc_commit_pair(bar_tfull(buf), pairMask);
          }
          if (last) TILE_T(tcount, 3, clock64());
#ifdef DMC_EPI_TIMING
          tw_issue += clock64() - tq1;
#endif
          if (two) { s = s1; ph = ph1; }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        TILE_T(tcount, 8, tw_full);
        TILE_T(tcount, 9, tw_issue);
      }
    }
  } else if (warp == 3 && kS3SplitProducer) {
    // ------------------------------------------------------------ W producer (both CTAs)
    // Issuing a TMA load costs the issuing warp ~200 clocks whatever its size (producer trace: 440 clocks in the issue
    // section per operand stage = two loads, against 80 waiting for a free slot): with A and W issued by ONE warp the
    // operand supply was bound by that warp's instruction stream, 3 500 of a tile's 5 800 clocks, while the ring was never
    // full.  The otherwise idle fourth warp takes the W half tiles; the leader's A producer still arms the stage's
    // barrier for all the bytes (complete_tx of W may land before that expect_tx: the phase cannot complete before the
    // leader's arrival).
    if (!(p.dbg & (1 | 32))) {
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      const bool sel = elect_one();
      const uint32_t lbar0 = mapa(bar_full(0), leaderRank);
      uint32_t e_nxt = my_tiles ? tab_at(0) : 0u;
      for (; tcount < my_tiles; ++tcount) {
        const uint32_t e = e_nxt;
        if (tcount + 1 < my_tiles) e_nxt = tab_at(tcount + 1);
        const int nt = p.cl4 ? 2 * (int)((e >> 20) & 0xffu) + (int)pairIdx : (int)((e >> 20) & 0xffu);
        const S3StageDev& S = p.st[e >> 28];
        const int kblk = S.kblk, k_blocks = S.k_blocks;
        const int n16 = (nt * S.BN + (int)(rank * ((uint32_t)S.BN >> 1))) >> 4;
        const CUtensorMap* tmW = &S.tmW;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(bar_empty(s), ph ^ 1, p.err, 1);
          if (sel) tma_load_pair_w(base + s * stageBytes + kPlanes * kS3APlane, tmW, n16, kb * kblk, lbar0 + 8u * s);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ publisher (both CTAs)
    // Announces a finished tile to the layers that depend on it.  The epilogue warps only count themselves in
    // (cta scope) once their stores of the tile are complete; the gpu-scope release -- a MEMBAR.ALL.GPU of ~1 000
    // clocks -- is this otherwise idle warp's job.  (When the last epilogue warp did it, it paid those clocks on the
    // critical path of the tile, stayed the last one, and set the pace of every chunk-add layer: 6 600 clocks per
    // tile with the MMAs done after 4 300.)
    uint32_t tcount = 0;
    uint32_t e_nxt = my_tiles ? tab_at(0) : 0u;
    for (; tcount < my_tiles; ++tcount) {
      const uint32_t e = e_nxt;
      if (tcount + 1 < my_tiles) e_nxt = tab_at(tcount + 1);
      const int l = (int)(e >> 28), mt = (int)(e & 0xfffffu);
      if (!p.st[l].publish) continue;
      const uint32_t cnt = warpsDone + 4u * (tcount & 7u);
      if (ld_acquire_cta_shared(cnt) < (uint32_t)kS3EpiWarps) {
        const long long t0 = clock64();
        while (ld_acquire_cta_shared(cnt) < (uint32_t)kS3EpiWarps) {
          __nanosleep(100);
          if (clock64() - t0 > 6000000000LL) mbar_timeout(p.err, 7);
        }
      }
      if (lane == 0) {
        // (tiles t and t + 8 of a CTA cannot be in flight together: two TMEM buffers)
        asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(cnt), "r"(0u) : "memory");
        fence_proxy_async_global();
        red_release_gpu_add(p.done + (size_t)l * p.MT + mt, 1u);
      }
      __syncwarp();
    }
  }
  } else {
    // ------------------------------------------------------------ epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kS3EpiRegs) : "memory");
    const int ew = warp - 4;
    EpiCtx x;
    x.ringBuf = epiBase + (uint32_t)ew * kS3WarpSmem;
    x.barRes = bar_res(ew, 0);
    x.slot = 0; x.res_phase = 0;
    x.lane = lane;
    x.quad = warp & 3;                         // TMEM lane quadrant this warp may read
    x.part = ew >> 2;                          // the warps of a quadrant split the columns
    x.rank = rank;
    x.depsOk = depsOk;
    x.epi_mem = !(p.dbg & 12);
#ifdef DMC_EPI_TIMING
    x.timed = blockIdx.x == 0 && ew == 0;
    x.ew = ew;
    x.tlast = clock64();
    for (int i = 0; i < 8; ++i) x.tacc[i] = 0;
#endif

    // Residual prefetch of chunk c2 of this CTA's tile number t2 (table entry e2); whole warp, lane 0 issues.
    // The first chunk of a NEW tile is only prefetched if the producer warp has already seen that tile's
    // dependencies satisfied (its residual may be written by an earlier layer of this chain); this check
    // never blocks -- blocking here could close a cycle through a tile this warp has not published yet --
    // and a prefetch that was skipped is issued when the tile starts (res_tile != tile number).
    uint32_t res_tile = 0xffffffffu;
    // `slot2` is the ring tile that chunk will be processed in; R2 are the (cached) fields of its layer.
    auto issue_res = [&](const StageRegs& R2, uint32_t e2, int c2, uint32_t t2, bool force, uint32_t slot2) {
      if (R2.kind != S3_RES || !x.epi_mem) return;
      const int mt2 = (int)(e2 & 0xfffffu);
      const int nt2 = p.cl4 ? 2 * (int)((e2 >> 20) & 0xffu) + (int)pairIdx : (int)((e2 >> 20) & 0xffu);
      if (c2 == 0) {
        if (!force && (int)(ld_acquire_cta_shared(depsOk) - (t2 + 1)) < 0) return;
        if (R2.need != 0u) fence_proxy_async_global();
        res_tile = t2;
      }
      const int dcol = s3_dest_col(S3_PLAIN, R2.BN, x.part, nt2, c2);
      if (dcol >= R2.n_out) return;
      if (lane == 0) {
        const uint32_t bar = x.barRes + 8u * slot2;
        mbar_expect_tx(bar, kS3ChunkBytes);
        tma_load_local(x.ringBuf + slot2 * kS3ChunkBytes, R2.tmRes, dcol,
                       mt2 * 256 + (int)rank * 128 + x.quad * 32, bar);
      }
    };

    uint32_t tcount = 0;
    uint32_t pend_t = 0;                           // tile whose completion is still to be published
    bool pending = false;
    // lane 0, after this warp's stores of the pending tile are complete: count this warp in (cta scope); the
    // publisher warp does the gpu-scope release once all epilogue warps of the CTA are in
    auto publish = [&]() {
      const uint32_t cnt = warpsDone + 4u * (pend_t & 7u);
#ifndef DMC_S3_NO_EPI_PUBLISH_FENCE
      fence_proxy_async_global();
#endif
      asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(cnt), "r"(1u) : "memory");
    };
    // table entries are fetched TWO tiles ahead: the entry of the next tile is needed at the top of this one (is it
    // in the same layer?), and a load issued there put its ~700-900 clocks of global latency between every two tiles
    // of this warp (tile trace: done(t) -> wait_acc(t+1) = 900 clocks with the accumulator long ready)
    uint32_t e_next = my_tiles ? tab_at(0) : 0u;
    uint32_t e_next2 = my_tiles > 1 ? tab_at(1) : 0u;
    int l_cached = -1;
    bool publish_l = false;
    uint32_t bias_tile = 0xffffffffu;            // tile number whose bias is already in its shared-memory buffer
    StageRegs S;
    for (; tcount < my_tiles; ++tcount) {
      const uint32_t e = e_next;
      const int l = (int)(e >> 28), mt = (int)(e & 0xfffffu);
      const int nt = p.cl4 ? 2 * (int)((e >> 20) & 0xffu) + (int)pairIdx : (int)((e >> 20) & 0xffu);
      if (l != l_cached) { S = s3_load_stage(p.st[l]); publish_l = p.st[l].publish != 0; l_cached = l; }
      const uint32_t buf = tcount & 1;
      const int nchunk = s3_nchunk(S.kind, S.BN, x.part);
      // bias of this warp's accumulator columns: lane i holds columns 2i, 2i+1 of the warp's window
      const int bias_col0 = s3_acc_col(S.kind == S3_PAIR ? S3_PAIR : S3_PLAIN, S.BN, x.part, 0);   // (+ up to 64 columns)
      auto bias_fetch = [&](int nt_) {
        float2 v = make_float2(0.f, 0.f);
        if (bias_col0 + 2 * lane < S.BN) v = __ldg(reinterpret_cast<const float2*>(S.bias + nt_ * S.BN + bias_col0) + lane);
        return v;
      };
      auto bias_put = [&](uint32_t b, float2 v) {
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(biasBase + ((uint32_t)ew * 2u + b) * (kS3BiasCols * 4u) + 8u * lane),
                     "f"(v.x), "f"(v.y) : "memory");
      };
      if (bias_tile != tcount) {                 // first tile of a layer: not requested a tile ahead
        bias_put(buf, bias_fetch(nt));
        __syncwarp();
      }
      x.biasBuf = biasBase + ((uint32_t)ew * 2u + buf) * (kS3BiasCols * 4u);
      x.biasCol0 = bias_col0;
      // the first residual of the next tile is prefetched early only within a layer (same cached fields);
      // across a layer boundary it is issued when that tile starts
      bool has_next = tcount + 1 < my_tiles;
      e_next = e_next2;
      if (tcount + 2 < my_tiles) e_next2 = tab_at(tcount + 2);
      if (has_next) has_next = (e_next >> 28) == (uint32_t)l;
      // (requested now, stored after this tile's work: its latency is never waited for)
      float2 bias_nxt = make_float2(0.f, 0.f);
      if (has_next)
        bias_nxt = bias_fetch(p.cl4 ? 2 * (int)((e_next >> 20) & 0xffu) + (int)pairIdx : (int)((e_next >> 20) & 0xffu));
      // Publishing the previous tile: if this tile's accumulator is not ready yet there is idle time (and this
      // tile may even depend on the previous one): wait for the stores and publish now.  If it is ready, its
      // dependencies were met long ago, so the publication can ride along with chunk 1 below at no cost.
      if (pending && !mbar_try(bar_tfull(buf), (tcount >> 1) & 1)) {
        if (lane == 0) { tma_store_wait_all(); publish(); }
        pending = false;
      }
      if (ew == 0) TILE_T(tcount, 4, clock64());
      mbar_wait(bar_tfull(buf), (tcount >> 1) & 1, p.err, 4);
      if (ew == 0) TILE_T(tcount, 5, clock64());
#ifdef DMC_EPI_TIMING
      x.tile = tcount;
      WARP_T(tcount, 0);
#endif
      tc_fence_after();
      if (res_tile != tcount && S.kind == S3_RES) {    // the early prefetch of chunk 0 was not possible
        // (a tile of a layer without residuals may have left the store two back still reading this staging tile)
        if (lane == 0) tma_store_wait_read1();
        __syncwarp();
        issue_res(S, e, 0, tcount, true, x.slot);
      }
      const uint32_t taddr = tmem_base_ld() + ((uint32_t)(x.quad * 32) << 16) + buf * 256u;
      auto release_tmem = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(bar_tempty(buf));
          else mbar_arrive_cluster(mapa(bar_tempty(buf), leaderRank));
#ifdef DMC_EPI_TIMING
          if (blockIdx.x == 0 && tcount < 256u) atomicMax(&g_tile_trace[tcount][13], (unsigned long long)clock64());
          if (blockIdx.x == 0 && tcount < 256u && ew == 1) g_tile_trace[tcount][10] = (unsigned long long)clock64();
#endif
        }
      };
      const bool stored0 = x.epi_mem && s3_dest_col(S.kind, S.BN, x.part, nt, 0) < S.n_out;
      auto next_res = [&](int c) {
        if (c == 1 && pending) {               // only chunk 0's store (if any) is newer than the previous tile's
          if (lane == 0) {
            if (stored0) tma_store_wait_1(); else tma_store_wait_all();
            publish();
          }
          pending = false;
        }
        // the item after chunk c lands one ring tile further if chunk c itself stores something
        const bool valid_c = s3_dest_col(S.kind, S.BN, x.part, nt, c) < S.n_out;
        uint32_t slot2 = x.slot + (valid_c ? 1u : 0u);
        if (slot2 == kS3Ring) slot2 = 0;
        if (c + 1 < nchunk) issue_res(S, e, c + 1, tcount, false, slot2);
        else if (has_next) issue_res(S, e_next, 0, tcount + 1, false, slot2);
      };
      if (nchunk == 0 || (p.dbg & 8)) {
        release_tmem();
      } else {
        s3_epilogue_tile(S, x, mt, nt, taddr, p.err, next_res, release_tmem);
      }
      if (pending) {
        // (tiles in which this warp has fewer than two chunks -- every tile of a chunk-add layer.)  Only the stores of
        // the PREVIOUS tile have to be complete: waiting for the store just issued put its whole round trip on the
        // critical path of the warp that publishes, which is the last one of the CTA and stays the last one.
        if (lane == 0) {
          if (nchunk == 1 && stored0) tma_store_wait_1(); else tma_store_wait_all();
          publish();
        }
        pending = false;
      }
      if (has_next) {
        bias_put(buf ^ 1u, bias_nxt);            // (that buffer was last read two tiles ago by this same warp)
        __syncwarp();
        bias_tile = tcount + 1;
      }
      if (ew == 0) TILE_T(tcount, 6, clock64());
#ifdef DMC_EPI_TIMING
      WARP_T(tcount, 2);
#endif
      pend_t = tcount;
      pending = publish_l;
      // Chains with fewer than ~3 waves of tiles per layer (the H/16 ... H/64 stages at batch 1) run into their
      // dependencies: a cluster's next tile needs rows another cluster has only just finished, and the deferred
      // publication above adds most of a tile time (2 500-5 000 clocks) to that wait.  There the ~1 500 clocks an
      // epilogue warp spends waiting for its stores right away are cheaper.
      if (pending && p.eager) {
        if (lane == 0) { tma_store_wait_all(); publish(); }
        pending = false;
      }
    }
    if (pending && lane == 0) { tma_store_wait_all(); publish(); }
    if (lane == 0) tma_store_wait_all();
#ifdef DMC_EPI_TIMING
    if (x.timed && lane == 0) {
      for (int i = 0; i < 8; ++i) atomicAdd(&g_epi_t[i], (unsigned long long)x.tacc[i]);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // nobody leaves while the peer can still touch its smem
  if (warp == 1) {
    const uint32_t ncols = 512;
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_ld()), "r"(ncols) : "memory");
  }
  if (threadIdx.x == 0) {
    // every CTA read the launch number when it started; the last one to leave closes the launch
    __threadfence();
    if (atomicAdd(p.epoch_ctr + 1, 1u) == gridDim.x - 1u) {
      p.epoch_ctr[1] = 0u;
      __threadfence();
      atomicAdd(p.epoch_ctr, 1u);
    }
  }
}

// ------------------------------------------------------------------ host side
static int g_s3_dbg = 0;
void gemm_s3_set_debug(int mask) { g_s3_dbg = mask; }
// Measurement pass of bench.py (dmc_profile_enable): an event pair around EVERY contraction launch.  The events break the
// back-to-back pipelining of the launches, and a cooperative launch then shows its full launch latency (~4 us per
// launch in a loop of single launches, none inside a frame: DepthConvBlock-256 141.8 vs 141.6 us, bench 209.2 vs 210.1
// P-frames/s with DMC_S3_COOP=1 / 0) -- the per-launch times would overstate the kernel's share of the frame (92 %
// against 81.5 % in the ncu launch list).  The pass therefore launches plainly; nothing else runs next to it.
static int g_s3_plain_launch = 0;
void gemm_s3_set_plain_launch(int on) { g_s3_plain_launch = on; }

bool gemm_s3_supports(const GemmW& w, const Epi& e, int nsplit) {
  if (!w.tmap_s3 || !w.tmap_s3_hi || e.do_clamp) return false;
  if (e.res2.p && !(e.res1.p && e.out.p && !e.out_f32 && e.pack == PACK_PLAIN && e.act == ACT_NONE)) return false;
  if (w.BN % 32 || w.BN > 128 || e.n_out % 16 || e.n_out < 32) return false;
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

struct S3Chain {
  S3ChainParams p;
  int n_stages = 0;
  int grid = 0, smem = 0;
  uint32_t* d_table = nullptr;
  uint32_t* d_done = nullptr;
  const float* scale_table[kS3MaxStages] = {};   // per-QP tables: row picked at launch
  int scale_C[kS3MaxStages] = {};
};

static int s3_kind(const Epi& e) {
  if (e.out_f32) return e.act == ACT_WSILU ? S3_F32_WSILU : S3_F32;
  if (e.pack == PACK_PAIR) return S3_PAIR;
  if (e.act == ACT_WSILU) return S3_WSILU;
  if (e.res2.p) return S3_RES2;
  return e.res1.p ? S3_RES : S3_PLAIN;
}

int s3_chain_max_stages() { return kS3MaxStages; }

// Per-device launch state: opt-in shared memory, how many CTA pairs can be co-resident (the tile dependencies of
// the persistent kernel need every cluster of the grid on an SM at the same time), the trap-code word.
struct S3Device {
  int smem_max = 0;
  int pair_clusters = 0;       // cudaOccupancyMaxActiveClusters for (2,1,1) clusters at smem_max
  int cl4_clusters = 0;        // same for (4,1,1) clusters (DMC_GEMM_CL4=1), 0 = unavailable
  int* d_err = nullptr;
};
static S3Device* s3_device() {
  static S3Device devs[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  S3Device& d = devs[dev];
  if (d.smem_max) return &d;
  int smem = 0;
  cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (cudaFuncSetAttribute(k_gemm_s3_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    snprintf(g_s3_err, sizeof g_s3_err, "cudaFuncSetAttribute(max dynamic smem=%d) failed", smem);
    return nullptr;
  }
  // the code of a bounded wait that gave up lives in mapped host memory: it survives the trap
  int* h_trap = nullptr;
  if (cudaHostAlloc((void**)&h_trap, sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    *h_trap = 0;
    g_s3_trap = h_trap;
    cudaHostGetDevicePointer((void**)&d.d_err, h_trap, 0);
  }
  for (int csz = 2; csz <= 4; csz += 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((num_sms() / csz) * csz);
    cfg.blockDim = dim3(kS3Threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, k_gemm_s3_chain, &cfg) != cudaSuccess) nc = 0;
    (csz == 2 ? d.pair_clusters : d.cl4_clusters) = nc;
  }
  cudaGetLastError();
  if (d.pair_clusters < 1) {
    snprintf(g_s3_err, sizeof g_s3_err, "k_gemm_s3_chain: no CTA pair can be resident on this device (smem %d)", smem);
    return nullptr;
  }
  d.smem_max = smem;
  return &d;
}

S3Chain* s3_chain_create(const S3StageDesc* stages, int n, long long M) {
  S3Device* dv = s3_device();
  if (!dv) return nullptr;
  const int smem_max = dv->smem_max;
  int* d_err = dv->d_err;
  if (n < 1 || n > kS3MaxStages) {
    snprintf(g_s3_err, sizeof g_s3_err, "s3_chain_create: %d stages (max %d)", n, kS3MaxStages);
    return nullptr;
  }
  // 4-CTA clusters (A shared between two CTA pairs by TMA multicast) need every stage's 64-row A map and are
  // limited to the clusters that can be co-resident (33 = 132 SMs on B200); DMC_GEMM_CL4=1 selects them
  static int cl4_mode = -1;
  if (cl4_mode < 0) {
    const char* v = getenv("DMC_GEMM_CL4");
    cl4_mode = (v && v[0] == '1') ? 1 : 0;
  }
  const int cl4_clusters = dv->cl4_clusters;
  if (cl4_mode == 1 && cl4_clusters < 1) cl4_mode = 0;
  bool cl4 = cl4_mode == 1;
  for (int l = 0; l < n; ++l) cl4 = cl4 && stages[l].tmA64 != nullptr;
  S3Chain* c = new S3Chain();
  memset(&c->p, 0, sizeof c->p);
  c->n_stages = n;
  c->p.cl4 = cl4 ? 1 : 0;
  const int MT = (int)((M + 255) / 256);
  int maxBN = 0, tiles_per_step = 0;
  for (int l = 0; l < n; ++l) {
    const S3StageDesc& d = stages[l];
    if (!gemm_s3_supports(*d.w, d.e, d.nsplit)) {
      snprintf(g_s3_err, sizeof g_s3_err, "s3_chain_create: stage %d unsupported (BN=%d act=%d pack=%d)", l, d.w->BN,
               d.e.act, d.e.pack);
      delete c;
      return nullptr;
    }
    S3StageDev& S = c->p.st[l];
    memcpy(&S.tmA, d.tmA, sizeof(CUtensorMap));
    memcpy(&S.tmA64, d.tmA64 ? d.tmA64 : d.tmA, sizeof(CUtensorMap));
    memcpy(&S.tmW, d.nsplit != 1 ? d.w->tmap_s3 : d.w->tmap_s3_hi, sizeof(CUtensorMap));
    memcpy(&S.tmOut, d.tmOut, sizeof(CUtensorMap));
    memcpy(&S.tmRes, d.tmRes ? d.tmRes : d.tmOut, sizeof(CUtensorMap));
    memcpy(&S.tmRes2, d.tmRes2 ? d.tmRes2 : d.tmOut, sizeof(CUtensorMap));
    S.bias = d.e.bias;
    S.scale = d.e.scale;
    S.kblk = d.nsplit != 1 ? 2 : s3_single_kblk();
    S.k_blocks = (d.K + 16 * S.kblk - 1) / (16 * S.kblk);
    S.BN = d.w->BN;
    S.n_tiles = (d.w->ncols + d.w->BN - 1) / d.w->BN;
    S.n_out = d.e.n_out;
    S.kind = s3_kind(d.e);
    S.nterms = d.nsplit != 1 ? 3 : 1;
    S.comp = S.nterms == 3 ? acc_comp_scaled(d.K) : 0.0f;
    // one count per CTA and table entry of the previous layer for these rows
    S.need = l == 0 ? 0u : (cl4 ? (uint32_t)((c->p.st[l - 1].n_tiles + 1) / 2) * 4u : (uint32_t)c->p.st[l - 1].n_tiles * 2u);
    S.publish = l + 1 < n ? 1 : 0;
    c->scale_table[l] = d.scale_table;
    c->scale_C[l] = d.scale_C;
    if (S.n_tiles > 255 || MT >= (1 << 20)) {
      snprintf(g_s3_err, sizeof g_s3_err, "s3_chain_create: tile index out of range");
      delete c;
      return nullptr;
    }
    if (S.BN > maxBN) maxBN = S.BN;
    tiles_per_step += S.n_tiles;
  }
  // smem plan
  const int stage_bytes = kPlanes * (kS3APlane + (maxBN / 2) * kS3BK * 2);
  const int fixed = kS3EpiWarps * kS3WarpSmem + kS3BarBytes + kS3BiasBytes;
  int nst = (smem_max - fixed) / stage_bytes;
  if (nst > 8) nst = 8;
  if (const char* v = getenv("DMC_S3_STAGES")) {       // experiments: cap the operand ring depth
    const int cap_st = atoi(v);
    if (cap_st >= 2 && cap_st < nst) nst = cap_st;
  }
  if (nst < 2) {
    snprintf(g_s3_err, sizeof g_s3_err, "s3_chain_create: stage of %d bytes does not fit", stage_bytes);
    delete c;
    return nullptr;
  }
  c->p.stageBytes = (uint32_t)stage_bytes;
  c->p.stages = nst;
  c->smem = nst * stage_bytes + fixed;
  // Tile table, layer-major: all tiles of layer 0 (row tile major, N tile minor), then layer 1, ...  Cluster u
  // takes entries u, u + U, ...: within a layer every tile costs the same, so the round-robin stays balanced, and
  // a tile's producers sit a whole layer earlier in the table -- they finished long before it is reached (a
  // simulated DepthConvBlock chain at 1920x1280 runs at 95 % of the no-dependency bound; interleaving the
  // layers with a lag of d row tiles reached 86-90 %, because the ramps starve and tile costs differ per layer).
  // the grid never exceeds what can be co-resident (fewer SMs under MPS / green contexts): every dependency of the
  // table points backwards and each cluster walks its entries in order, so a fully resident grid cannot deadlock
  const int cap = cl4 ? 4 * cl4_clusters : std::min(num_sms() & ~1, 2 * dv->pair_clusters);
  std::vector<uint32_t> table;
  // DMC_S3_GROUPS=g (experiment): the row tiles are cut into g groups and the table is layer-major INSIDE a group, group
  // after group -- the intermediates of a group (1/g of o1, u, ...) are consumed while they are still in L2 instead of
  // after a whole layer has been written (358 MB of DRAM traffic per DepthConvBlock-256 chain against 157 MB at its
  // boundary with g = 1).  Dependencies still point backwards; groups do not depend on each other.
  static int groups_env = -1;
  if (groups_env < 0) {
    const char* v = getenv("DMC_S3_GROUPS");
    groups_env = v ? std::max(1, atoi(v)) : 1;
  }
  const int groups = (n > 1) ? std::min(groups_env, MT) : 1;
  for (int g = 0; g < groups; ++g) {
    const int mt0 = (int)((long long)MT * g / groups), mt1 = (int)((long long)MT * (g + 1) / groups);
    for (int l = 0; l < n; ++l)
      for (int mt = mt0; mt < mt1; ++mt) {
        // an entry is one N tile for a CTA pair, or two adjacent N tiles for the two pairs of a 4-CTA cluster
        const int per_mt = cl4 ? (c->p.st[l].n_tiles + 1) / 2 : c->p.st[l].n_tiles;
        for (int nt = 0; nt < per_mt; ++nt)
          table.push_back(((uint32_t)l << 28) | ((uint32_t)nt << 20) | (uint32_t)mt);
      }
  }
  (void)tiles_per_step;
  c->p.n_entries = (int)table.size();
  {
    const char* v = getenv("DMC_S3_EAGER");          // 0 / 1 forces the publication mode (A/B runs)
    const int units = cap / (cl4 ? 4 : 2);
    int min_layer = 1 << 30;                          // fewest entries of any layer of the chain
    for (int l = 0; l < n; ++l) {
      const int per_mt = cl4 ? (c->p.st[l].n_tiles + 1) / 2 : c->p.st[l].n_tiles;
      if (per_mt * MT < min_layer) min_layer = per_mt * MT;
    }
    c->p.eager = v ? (v[0] == '1') : (n > 1 && min_layer / groups < 3 * units);
  }
  c->p.MT = MT;
  c->p.err = d_err;
  int grid = (cl4 ? 4 : 2) * c->p.n_entries;
  if (grid > cap) grid = cap;
  c->grid = grid;
  if (cudaMalloc(&c->d_table, table.size() * sizeof(uint32_t)) != cudaSuccess ||
      cudaMalloc(&c->d_done, ((size_t)n * MT + 2) * sizeof(uint32_t)) != cudaSuccess) {
    snprintf(g_s3_err, sizeof g_s3_err, "s3_chain_create: cudaMalloc failed");
    s3_chain_destroy(c);
    return nullptr;
  }
  cudaMemcpy(c->d_table, table.data(), table.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
  cudaMemset(c->d_done, 0, ((size_t)n * MT + 2) * sizeof(uint32_t));
  c->p.table = c->d_table;
  c->p.done = c->d_done;
  c->p.epoch_ctr = c->d_done + (size_t)n * MT;
  return c;
}

void s3_chain_destroy(S3Chain* c) {
  if (!c) return;
  if (c->d_table) cudaFree(c->d_table);
  if (c->d_done) cudaFree(c->d_done);
  delete c;
}

int s3_chain_stages(const S3Chain* c) { return c ? c->n_stages : 0; }

int s3_chain_launch(S3Chain* c, int qp, cudaStream_t st) {
  c->p.dbg = g_s3_dbg;
  for (int l = 0; l < c->n_stages; ++l)
    if (c->scale_table[l]) c->p.st[l].scale = c->scale_table[l] + (size_t)qp * c->scale_C[l];
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(c->grid);
  cfg.blockDim = dim3(kS3Threads);
  cfg.dynamicSmemBytes = c->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = c->p.cl4 ? 4 : 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  // Cooperative launch: the grid only starts once ALL its clusters can be resident.  The tile dependencies make
  // resident clusters spin on tiles owned by clusters that are not scheduled yet; alone on the device the grid
  // (capped by the occupancy query) is always fully resident, but next to another persistent kernel on a second stream
  // two partially resident grids could wait for each other until the bounded waits trap.  DMC_S3_COOP=0 switches it
  // off (A/B runs); a driver that refuses the combination with clusters falls back once and for all.
  static int coop = -1;
  if (coop < 0) {
    const char* v = getenv("DMC_S3_COOP");
    if (v) {
      coop = v[0] == '0' ? 0 : 1;
    } else {
      // Nsight Compute (and the other injection-based tools) fail a cooperative launch of a cluster kernel with
      // "LaunchFailed" -- a sticky error that takes the context with it -- and they serialise kernels anyway, so no second
      // persistent kernel can be co-resident with this one: under a tool the plain launch is both necessary and safe.
      static const char* const kTools[] = {"NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NV_NSIGHT_INJECTION_PORT_BASE",
                                           "NV_SANITIZER_INJECTION_PORT_BASE", "CUDA_INJECTION64_PATH"};
      coop = 1;
      for (const char* k : kTools)
        if (getenv(k)) coop = 0;
    }
  }
  const bool use_coop = coop && !g_s3_plain_launch;
  if (use_coop) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  } else if (pdl_enabled()) {                  // (programmatic launch and cooperative launch exclude each other)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  note_launch();
  cudaError_t err = cudaLaunchKernelEx(&cfg, k_gemm_s3_chain, c->p);
  if (err != cudaSuccess && use_coop) {        // cooperative + cluster launch refused: plain launch from now on
    cudaGetLastError();
    coop = 0;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, k_gemm_s3_chain, c->p);
  }
  if (err != cudaSuccess) {
    snprintf(g_s3_err, sizeof g_s3_err, "k_gemm_s3_chain launch: %s", cudaGetErrorString(err));
    return -1;
  }
  return 0;
}

}  // namespace dmc

#ifdef DMC_EPI_TIMING
extern "C" __attribute__((visibility("default"))) int dmc_debug_warp_trace(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, dmc::g_warp_trace, sizeof(unsigned long long) * 2 * 2 * 256 * 3) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int dmc_debug_tile_trace(unsigned long long* out, int ntiles) {
  return cudaMemcpyFromSymbol(out, dmc::g_tile_trace, sizeof(unsigned long long) * 14 * (size_t)ntiles) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int dmc_debug_epi_timing(unsigned long long* out8, int reset) {
  if (out8) cudaMemcpyFromSymbol(out8, dmc::g_epi_t, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(dmc::g_epi_t, z, sizeof z);
  }
  return 0;
}
#endif
