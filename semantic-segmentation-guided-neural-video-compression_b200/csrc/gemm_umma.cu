// tcgen05 / TMA / TMEM contraction kernel for sm_100a:  out = epilogue(A[M,K] . W[N,K]^T).
//
// Operands are split-fp16 tensors (two planes hi / 2^11-scaled lo, see common.cuh).  One persistent CTA
// per SM walks 128 x BN output tiles.  Warp roles:
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D maps {K, rows, plane}, SWIZZLE_128B)
//               into a ring of smem stages, completion on mbarriers
//   warp 1      MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::f16 (fp16 x fp16 ->
//               fp32 in TMEM).  nsplit != 1 issues the 3 terms hh | hl, lh per 16-wide k step
//               (fp32-grade product), nsplit == 1 issues hh only.
//               hh has its own accumulator; the two cross terms carry the 2^11 scale of the lo planes
//               and share a second one; the epilogue computes main + small * 2^-11 in one fp32 FMA.
//               (The tensor core adds into its fp32 accumulator with truncation, a bias that grows
//               with the number of adds -- measured ~0.2 ulp per MMA: keeping the small terms out of
//               the main accumulator also keeps its truncating adds to one per k step.)
//   warps 2..9  epilogue (two warps per TMEM lane quadrant, alternating 32-column chunks):
//               tcgen05.ld the accumulator (lane == output row), bias / WSiLU / chunk-add pairing /
//               residuals / per-channel scale, split back into S3 planes; residual loads and
//               output stores are staged through shared memory so they are coalesced.  TMEM is double buffered (column bases 0 and 256, each
//               holding the main accumulator at +0 and the small-terms one at +128) so the
//               epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace dmc {

static char g_umma_err[512] = "";
const char* umma_last_error() { return g_umma_err; }

// ------------------------------------------------------------------ tensor maps (host)
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

// 4-D maps over tile-blocked planes (common.cuh): {256 elements = 16 rows x 16 columns, rows / 16, column blocks,
// planes}; a box is one plane of `box_rows` rows x 4 column blocks (64 k), data pre-swizzled -> no TMA swizzle.
static int encode4d(void* out, const void* base, uint64_t rows16, uint64_t blocks, uint64_t block_bytes,
                    uint64_t plane_bytes, uint32_t box_rows) {
  auto fn = get_encode();
  if (!fn) {
    snprintf(g_umma_err, sizeof g_umma_err, "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t dims[4] = {256, rows16, blocks, (cuuint64_t)kPlanes};
  cuuint64_t strides[3] = {512, block_bytes, plane_bytes};
  cuuint32_t box[4] = {256, box_rows / 16, 4, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_umma_err, sizeof g_umma_err,
             "cuTensorMapEncodeTiled failed (%d): base=%p rows16=%llu blocks=%llu box_rows=%u", (int)r, base,
             (unsigned long long)rows16, (unsigned long long)blocks, box_rows);
    return -1;
  }
  return 0;
}

int make_tmap_act(void* tmap_out, View a, long long M) {
  (void)M;
  return encode4d(tmap_out, a.p, (uint64_t)(a.bs / 256), (uint64_t)((a.C + 15) / 16), (uint64_t)a.bs * 2,
                  (uint64_t)a.ps * 2, 128);
}
int make_tmap_weight(void* tmap_out, const GemmW& w, int box_rows) {
  if (!w.wb) {
    snprintf(g_umma_err, sizeof g_umma_err, "make_tmap_weight: no blocked weight copy");
    return -1;
  }
  return encode4d(tmap_out, w.wb, (uint64_t)w.Npad / 16, (uint64_t)w.Kld / 16, (uint64_t)w.Npad * 32,
                  (uint64_t)w.Npad * w.Kld * 2, (uint32_t)box_rows);
}

// ------------------------------------------------------------------ device helpers (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 6000000000LL) {   // ~3 s at 2 GHz: no legitimate wait is this long
      if (err) atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
}
// (tile-blocked maps: k0 and row are element / row coordinates, converted to {0, row / 16, k block, plane})
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int k0, int row,
                                            int plane, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(0), "r"(row >> 4), "r"(k0 >> 4), "r"(plane), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_32B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout) -- a stage plane is
// four 16-wide k blocks of [rows][32 B], the image of the tile-blocked tensors:
// start>>4 [0,14) | LBO>>4 = 1 [16,30) | SBO>>4 = 16 (8 rows x 32 B) [32,46) | version 1 [46,48)
// | layout SWIZZLE_32B = 6 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)16 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}

struct UmmaParams {
  long long M;
  int m_tiles, n_tiles, k_blocks;
  int BN, nsplit, stages;
  int dbg;     // probe switches: 1 = no TMA loads, 2 = no MMA issue, 4 = no epilogue global traffic
  float comp;  // accumulate-truncation compensation of the main accumulator (acc_comp_scaled in kernels.cu)
  int* err;
};

constexpr int kATileBytes = 128 * 64 * 2;   // one plane of a 128 x 64 fp16 tile
constexpr int kThreads = 64 + 32 * 8;   // TMA warp, MMA warp, 8 epilogue warps

// ---- epilogue of one 32-column chunk, executed by a whole warp (lane == accumulator row) ----
// The epilogue is issue-bound (the first versions spent 30-60 warp instructions per output element,
// more SM issue slots per tile than the 6144-cycle main loop), so everything here is written for
// instruction count: packed bf16x2 conversions, a reciprocal-based WSiLU, row addresses computed once
// per tile, no shuffles.  Global traffic is staged through a per-warp shared-memory tile (16 rows x
// 64 B payload, rows padded to 80 B): a warp-level access covers 8 rows x 64 contiguous bytes
// instead of 32 rows x 16 B, which cuts the L1 wavefronts per tile 4x.
constexpr int kStageRowBytes = 80;
constexpr int kStageBytes = 16 * kStageRowBytes;   // rows pass through 16 at a time
constexpr int kEpiWarps = 8;

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
               : "memory");
  return v;
}
// layers.py:8-10 silu(4x)/4 == x / (1 + exp(-4x)); the reciprocal is MUFU.RCP (<= 1 ulp) instead of
// an IEEE division: <= 2 ulp from the reference's result, 10 instructions instead of ~30.
__device__ __forceinline__ float wsilu_fast(float x) {
  const float e = expf(mul_rn(-4.0f, x));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(add_rn(1.0f, e)));
  return mul_rn(x, r);
}
__device__ __forceinline__ float act_fast(float v, int act) {
  if (act == ACT_WSILU) return wsilu_fast(v);
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// Rows of the staging passes this lane touches: row group i covers tile rows i*8 + lane/4.
struct RowMap {
  long long base[4];   // destination row (before the pixel-shuffle group offset)
  bool ok[4];
};

// one plane of 32 rows x 32 columns through the staging tile: t = g * 2^-11 (lo plane, first) or t += g (hi)
template <bool kFirst>
__device__ __forceinline__ void stage_plane_in(const uint4 (&g)[4], uint32_t wr, uint32_t rd, int lane, float* t) {
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    st_shared_v4(wr, g[2 * pass]);
    st_shared_v4(wr + 8 * kStageRowBytes, g[2 * pass + 1]);
    __syncwarp();
    if ((lane >> 4) == pass) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 q = ld_shared_v4(rd + j * 16);
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = h2lo(u[k]), hi = h2hi(u[k]);
          t[8 * j + 2 * k] = kFirst ? mul_rn(lo, kLoInv) : add_rn(t[8 * j + 2 * k], lo);
          t[8 * j + 2 * k + 1] = kFirst ? mul_rn(hi, kLoInv) : add_rn(t[8 * j + 2 * k + 1], hi);
        }
      }
    }
    __syncwarp();
  }
}

// t[32] = hi + lo * 2^-11 of the 32 columns [dcol, dcol+32) of this lane's row of `src` (exactly
// join2).  All 8 global loads (2 planes x 4 row groups) are issued before the first use, so a
// residual costs one memory round trip; the rows then pass through the staging tile 16 at a time.
__device__ __forceinline__ void staged_load_s3(const View& src, const RowMap& rm, long long goff, int dcol,
                                               int ncols, uint32_t stage, int lane, float* t) {
  const int slot = lane & 3;
  const bool col_ok = slot * 8 < ncols;
  uint4 gh[4], gl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const h16* q = src.p + s3_unit_offset(src, rm.base[i] + goff, dcol + slot * 8);
    gh[i] = gl[i] = make_uint4(0, 0, 0, 0);
    if (rm.ok[i] && col_ok) {
      gh[i] = *reinterpret_cast<const uint4*>(q);
      gl[i] = *reinterpret_cast<const uint4*>(q + src.ps);
    }
  }
  const uint32_t wr = stage + (lane >> 2) * kStageRowBytes + slot * 16;
  const uint32_t rd = stage + (lane & 15) * kStageRowBytes;
  stage_plane_in<true>(gl, wr, rd, lane, t);
  stage_plane_in<false>(gh, wr, rd, lane, t);
}

// v: accumulator (+ partner for PACK_PAIR in r1) of columns [n0, n0+32) of this lane's row.
__device__ __forceinline__ void epilogue_chunk(const Epi& e, const RowMap& rm, long long m_own, bool own_ok,
                                               int n0, uint32_t* r0, const uint32_t* r1, uint32_t stage,
                                               int lane) {
  // destination of packed columns [n0, n0+32) (PACK_PAIR: [n0, n0+32) + partners [n0+32, n0+64))
  int dcol = n0, limit = e.n_out;
  long long goff = 0;                              // pixel-shuffle group offset of the destination row
  if (e.pack == PACK_PAIR) {
    dcol = (n0 >> 6) * 32;
  } else if (e.pack == PACK_SHUF2) {
    const int g = n0 / e.Cg_pad;
    dcol = n0 - g * e.Cg_pad;
    limit = e.Cg;
    goff = (long long)(g >> 1) * (2LL * e.W) + (g & 1);
  }
  if (dcol >= limit) return;                       // warp-uniform
  const int ncols = min(32, limit - dcol);         // multiple of 8, warp-uniform
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]);
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + i));
      v[i] = add_rn(v[i], b.x); v[i + 1] = add_rn(v[i + 1], b.y);
      v[i + 2] = add_rn(v[i + 2], b.z); v[i + 3] = add_rn(v[i + 3], b.w);
    }
  }
  if (e.act != ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = act_fast(v[i], e.act);
  }
  if (e.pack == PACK_PAIR) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + 32 + i));
      v[i] = add_rn(v[i], act_fast(add_rn(__uint_as_float(r1[i]), b.x), e.act));
      v[i + 1] = add_rn(v[i + 1], act_fast(add_rn(__uint_as_float(r1[i + 1]), b.y), e.act));
      v[i + 2] = add_rn(v[i + 2], act_fast(add_rn(__uint_as_float(r1[i + 2]), b.z), e.act));
      v[i + 3] = add_rn(v[i + 3], act_fast(add_rn(__uint_as_float(r1[i + 3]), b.w), e.act));
    }
  }
  if (e.res1.p) {
    float t[32];
    staged_load_s3(e.res1, rm, goff, dcol, ncols, stage, lane, t);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = add_rn(v[i], t[i]);
  }
  if (e.res2.p) {
    float t[32];
    staged_load_s3(e.res2, rm, goff, dcol, ncols, stage, lane, t);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = add_rn(v[i], t[i]);
  }
  if (e.scale) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (i < ncols) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(e.scale + dcol + i));
        v[i] = mul_rn(v[i], s4.x); v[i + 1] = mul_rn(v[i + 1], s4.y);
        v[i + 2] = mul_rn(v[i + 2], s4.z); v[i + 3] = mul_rn(v[i + 3], s4.w);
      }
    }
  }
  if (e.do_clamp) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i], e.clamp_lo), e.clamp_hi);
  }
  if (e.out_f32 && own_ok) {                       // fp32 rows (2 launches per frame): direct stores
    long long drow = m_own;
    if (e.pack == PACK_SHUF2) {
      const int w = (int)(m_own % e.W);
      const long long tt = m_own / e.W;
      drow = ((tt / e.H) * (2 * e.H) + 2 * (int)(tt % e.H)) * (2LL * e.W) + 2 * w + goff;
    }
    float* d = e.out_f32 + drow * e.ld_f32 + dcol;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      if (i < ncols) *reinterpret_cast<float4*>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  if (e.out.p) {
    // 2-way split, two elements per conversion: hi = f16x2(v), lo = f16x2((v - hi) * 2^11)
    uint32_t* ph = r0;                              // the accumulator registers are dead by now
    uint32_t pl[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) split2x2(v[2 * i], v[2 * i + 1], ph[i], pl[i]);
    const int slot = lane & 3;
    const bool col_ok = slot * 8 < ncols;
    const uint32_t wr = stage + (lane & 15) * kStageRowBytes;
    const uint32_t rd = stage + (lane >> 2) * kStageRowBytes + slot * 16;
    h16* dst[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = e.out.p + s3_unit_offset(e.out, rm.base[i] + goff, dcol + slot * 8);
    auto store_plane = [&](const uint32_t* q, long long poff) {
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        if ((lane >> 4) == pass) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(wr + j * 16, make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint4 val = ld_shared_v4(rd + i * 8 * kStageRowBytes);
          if (rm.ok[2 * pass + i] && col_ok) *reinterpret_cast<uint4*>(dst[2 * pass + i] + poff) = val;
        }
        __syncwarp();
      }
    };
    store_plane(ph, 0);
    store_plane(pl, e.out.ps);
  }
}

// ---- cluster / pair (cta_group::2) helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {   // same offset in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, int k0, int row, int plane,
                                                 uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(0), "r"(row >> 4), "r"(k0 >> 4), "r"(plane), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in both CTAs of the pair
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

// kPair == false: one CTA per 128 x BN tile (cta_group::1).
// kPair == true : a cluster of two CTAs owns a 256 x BN tile (cta_group::2).  Each CTA loads its own
//                 128 rows of A and HALF of the W tile (BN/2 rows); the leader CTA issues one
//                 M=256 MMA that reads W from both CTAs' shared memory and writes each CTA's 128
//                 accumulator rows into that CTA's TMEM.  L2->SM operand traffic per MMA drops to
//                 3/4 and a stage shrinks from 64 KB to 48 KB, so the ring holds 4 stages.
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const Epi e, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const uint32_t wRows = kPair ? (uint32_t)p.BN / 2u : (uint32_t)p.BN;
  const uint32_t wTileBytes = wRows * 128u;
  const uint32_t stageBytes = (uint32_t)p.nsplit * (kATileBytes + wTileBytes);
  const uint32_t barBase = base + (uint32_t)p.stages * stageBytes;
  // barriers: full[8] | empty[8] | tfull[2] | tempty[2] | tmem ptr
  auto bar_full = [&](int s) { return barBase + 8u * s; };
  auto bar_empty = [&](int s) { return barBase + 64u + 8u * s; };
  auto bar_tfull = [&](int b) { return barBase + 128u + 8u * b; };
  auto bar_tempty = [&](int b) { return barBase + 144u + 8u * b; };
  const uint32_t tmemSlot = barBase + 160u;
  const uint32_t stageBase = barBase + 256u;          // kEpiWarps x kStageBytes of epilogue staging

  if (warp == 0 && lane == 0) {
    if (base & 1023u) {                               // SWIZZLE_128B tiles need 1024-byte alignment
      if (p.err) atomicExch(p.err, 9);
      __trap();
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull(b), 1);
      mbar_init(bar_tempty(b), kEpiWarps * (kPair ? 2 : 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    const uint32_t ncols = 512;
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmemSlot), "r"(ncols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmemSlot), "r"(ncols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();                      // peer barriers are initialised before any remote use
  tc_fence_after();
  pdl_prologue_done();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmemSlot) : "memory");

  // tile walk: (cluster of) CTA(s) `unit` takes tiles unit, unit + units, ...
  const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int units = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int rowsPerTile = kPair ? 256 : 128;
  const int total_tiles = p.m_tiles * p.n_tiles;      // m_tiles counts 256-row tiles when kPair

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = unit; tile < total_tiles; tile += units) {
        const int m_idx = (tile / p.n_tiles) * rowsPerTile + (int)rank * 128;
        const int n_idx = (tile % p.n_tiles) * p.BN + (int)rank * (int)(kPair ? wRows : 0u);
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(bar_empty(s), ph ^ 1, p.err, 1);
          const uint32_t sa = base + s * stageBytes;
          const uint32_t sw = sa + p.nsplit * kATileBytes;
          if (p.dbg & 1) {
            if (leader) mbar_arrive(bar_full(s));
          } else if (kPair) {
            // both CTAs' bytes complete on the LEADER's barrier, which the leader arms for 2 stages' worth
            if (leader) mbar_expect_tx(bar_full(s), 2u * stageBytes);
            const uint32_t lbar = mapa(bar_full(s), 0);
            for (int pl = 0; pl < p.nsplit; ++pl) {
              tma_load_3d_pair(sa + pl * kATileBytes, &tmA, kb * 64, m_idx, pl, lbar);
              tma_load_3d_pair(sw + pl * wTileBytes, &tmW, kb * 64, n_idx, pl, lbar);
            }
          } else {
            mbar_expect_tx(bar_full(s), stageBytes);
            for (int pl = 0; pl < p.nsplit; ++pl) {
              tma_load_3d(sa + pl * kATileBytes, &tmA, kb * 64, m_idx, pl, bar_full(s));
              tma_load_3d(sw + pl * wTileBytes, &tmW, kb * 64, n_idx, pl, bar_full(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // instruction descriptor: D=f32 [4,6)=1, A=f16 [7,10)=0, B=f16 [10,13)=0, K-major both,
      // N>>3 at [17,23), M>>4 at [24,29)  (M = 256 for the pair)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)((kPair ? 256 : 128) >> 4) << 24);
      const int nterms = (p.nsplit != 1) ? 3 : 1;
      // term 0 = hi*hi -> main accumulator; hi*lo', lo'*hi (2^11-scaled) -> second accumulator
      const int ta[3] = {0, 0, 1};
      const int tw[3] = {0, 1, 0};
      uint32_t it = 0, tcount = 0;
      for (int tile = unit; tile < total_tiles; tile += units, ++tcount) {
        const int buf = tcount & 1;
        const uint32_t tph = (tcount >> 1) & 1;
        mbar_wait(bar_tempty(buf), tph ^ 1, p.err, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u;
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(bar_full(s), ph, p.err, 3);
          tc_fence_after();
          const uint32_t sa = base + s * stageBytes;
          const uint32_t sw = sa + p.nsplit * kATileBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (p.dbg & 2) break;
            for (int t = 0; t < nterms; ++t) {
              const uint64_t ad = make_desc(sa + ta[t] * kATileBytes + ks * 4096);
              const uint64_t bd = make_desc(sw + tw[t] * wTileBytes + ks * (wRows * 32u));
              const uint32_t d = t == 0 ? d_tmem : d_tmem + 128u;
              const uint32_t acc = t == 0 ? ((kb | ks) ? 1u : 0u) : ((kb | ks | (t - 1)) ? 1u : 0u);
              if (kPair) tc_mma_pair(d, ad, bd, idesc, acc);
              else tc_mma(d, ad, bd, idesc, acc);
            }
          }
          if (kPair) tc_commit_pair(bar_empty(s));
          else tc_commit(bar_empty(s));
        }
        if (kPair) tc_commit_pair(bar_tfull(buf));
        else tc_commit(bar_tfull(buf));
      }
    }
  } else {
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;           // two warps per quadrant split the column chunks
    const uint32_t stage = stageBase + (uint32_t)(warp - 2) * kStageBytes;
    uint32_t tcount = 0;
    for (int tile = unit; tile < total_tiles; tile += units, ++tcount) {
      const int buf = tcount & 1;
      const uint32_t tph = (tcount >> 1) & 1;
      const long long m_warp0 = (long long)(tile / p.n_tiles) * rowsPerTile + (long long)rank * 128 + quad * 32;
      const long long m = m_warp0 + lane;
      const int n_idx = (tile % p.n_tiles) * p.BN;
      const bool row_ok = (m < p.M) && !(p.dbg & 4);
      RowMap rm;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long mr = m_warp0 + i * 8 + (lane >> 2);
        rm.ok[i] = (mr < p.M) && !(p.dbg & 4);
        rm.base[i] = mr;
        if (e.pack == PACK_SHUF2) {                  // nn.PixelShuffle(2): (b, h, w) -> (b, 2h, 2w) + group offset
          const int w = (int)(mr % e.W);
          const long long tt = mr / e.W;
          rm.base[i] = ((tt / e.H) * (2 * e.H) + 2 * (int)(tt % e.H)) * (2LL * e.W) + 2 * w;
        }
      }
      mbar_wait(bar_tfull(buf), tph, p.err, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * 256u;
      const bool two_acc = p.nsplit != 1;
      auto load_chunk = [&](int c, uint32_t* r) {     // accumulator columns [c, c+32) of this row
        tc_ld32(taddr + c, r);
        if (two_acc) {
          uint32_t s2[32];
          tc_ld32(taddr + 128u + c, s2);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            r[i] = __float_as_uint(fmaf(fmaf(__uint_as_float(r[i]), p.comp, __uint_as_float(s2[i])), kLoInv,
                                        __uint_as_float(r[i])));
        } else {
          tc_wait_ld();
        }
      };
      if (e.pack == PACK_PAIR) {
        for (int c0 = half * 64; c0 < p.BN; c0 += 128) {
          uint32_t r0[32], r1[32];
          load_chunk(c0, r0);
          load_chunk(c0 + 32, r1);
          epilogue_chunk(e, rm, m, row_ok, n_idx + c0, r0, r1, stage, lane);
        }
      } else {
        for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
          uint32_t r0[32];
          load_chunk(c0, r0);
          epilogue_chunk(e, rm, m, row_ok, n_idx + c0, r0, r0, stage, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!kPair || leader) mbar_arrive(bar_tempty(buf));
        else mbar_arrive_cluster(mapa(bar_tempty(buf), 0));   // the leader's MMA thread waits for both CTAs
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();                      // nobody leaves while the peer can still touch its smem
  if (warp == 1) {
    const uint32_t ncols = 512;
    if (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

static int g_umma_dbg = 0;
void umma_set_debug(int mask) { g_umma_dbg = mask; }
static bool g_pair_enabled = true;
void umma_set_pair(bool on) { g_pair_enabled = on; }
static void read_env_once() {
  static bool done = false;
  if (done) return;
  done = true;
  const char* v = getenv("DMC_UMMA_PAIR");     // DMC_UMMA_PAIR=0 forces the one-CTA kernel (A/B runs)
  if (v && v[0] == '0') g_pair_enabled = false;
}

int gemm_umma(const void* tmapA, const GemmW& w, const Epi& e, long long M, int K, int nsplit,
              cudaStream_t st) {
  static int smem_max_dev[64] = {};
  static int* d_err_dev[64] = {};
  read_env_once();
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!smem_max_dev[dev]) {
    int smem = 0;
    cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (cudaFuncSetAttribute(k_gemm_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_gemm_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      snprintf(g_umma_err, sizeof g_umma_err, "cudaFuncSetAttribute(max dynamic smem=%d) failed", smem);
      return -1;
    }
    cudaMalloc(&d_err_dev[dev], sizeof(int));
    cudaMemset(d_err_dev[dev], 0, sizeof(int));
    smem_max_dev[dev] = smem;
  }
  const int smem_max = smem_max_dev[dev];
  int* d_err = d_err_dev[dev];
  if (nsplit != 1) nsplit = kPlanes;             // planes staged per operand: both, or hi only
  if (!w.tmap || (w.BN % 32) || w.BN > (nsplit != 1 ? 128 : 256) || (e.pack == PACK_PAIR && (w.BN % 64))) {
    snprintf(g_umma_err, sizeof g_umma_err, "gemm_umma: unsupported weight tiling BN=%d", w.BN);
    return -1;
  }
  // the pair kernel needs BN/2 rows of W per CTA to be a whole number of 8-row swizzle groups
  const bool pair = g_pair_enabled && w.tmap_half && (w.BN % 32 == 0) && M > 128;
  UmmaParams p;
  p.M = M;
  p.m_tiles = pair ? (int)((M + 255) / 256) : (int)((M + 127) / 128);
  p.n_tiles = (w.ncols + w.BN - 1) / w.BN;
  p.k_blocks = (K + 63) / 64;
  p.BN = w.BN;
  p.nsplit = nsplit;
  const int stage_bytes = nsplit * (kATileBytes + (pair ? w.BN / 2 : w.BN) * 128);
  const int fixed = 256 + kEpiWarps * kStageBytes;   // barriers + epilogue staging (base is 1024-aligned)
  int stages = (smem_max - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 1) {
    snprintf(g_umma_err, sizeof g_umma_err, "gemm_umma: stage of %d bytes does not fit", stage_bytes);
    return -1;
  }
  p.stages = stages;
  p.dbg = g_umma_dbg;
  p.comp = nsplit != 1 ? acc_comp_scaled(K) : 0.0f;
  p.err = d_err;
  const int smem = stages * stage_bytes + fixed;
  CUtensorMap ta, tw;
  memcpy(&ta, tmapA, sizeof ta);
  memcpy(&tw, pair ? w.tmap_half : w.tmap, sizeof tw);
  cudaError_t err;
  note_launch();
  if (pair) {
    int grid = 2 * p.m_tiles * p.n_tiles;
    const int cap = num_sms() & ~1;
    if (grid > cap) grid = cap;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    err = cudaLaunchKernelEx(&cfg, k_gemm_umma<true>, ta, tw, e, p);
  } else {
    int grid = p.m_tiles * p.n_tiles;
    if (grid > num_sms()) grid = num_sms();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    err = cudaLaunchKernelEx(&cfg, k_gemm_umma<false>, ta, tw, e, p);
  }
  if (err != cudaSuccess) {
    snprintf(g_umma_err, sizeof g_umma_err, "k_gemm_umma launch: %s", cudaGetErrorString(err));
    return -1;
  }
  return 0;
}

}  // namespace dmc
