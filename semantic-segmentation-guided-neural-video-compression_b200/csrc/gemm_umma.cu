// tcgen05 / TMA / TMEM contraction kernel for sm_100a:  out = epilogue(A[M,K] . W[N,K]^T).
//
// Operands are S3 tensors (three bf16 planes hi/mid/lo, see common.cuh).  One persistent CTA
// per SM walks 128 x BN output tiles.  Warp roles:
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D maps {K, rows, plane}, SWIZZLE_128B)
//               into a ring of smem stages, completion on mbarriers
//   warp 1      MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 ->
//               fp32 in TMEM).  nsplit == 3 issues the 6 cross terms hh,hm,mh,hl,lh,mm per
//               16-wide k step (fp32-grade product), nsplit == 1 issues hh only.
//               The tensor core adds into its fp32 accumulator with truncation, a bias that grows
//               with the number of adds (measured: ~0.2 ulp per MMA).  The dominant hh term
//               therefore has its own accumulator and the five small terms (<= 2^-8 of it) share
//               a second one; the epilogue adds the two in round-to-nearest fp32.  That cuts the
//               truncating adds into the large accumulator 6x.
//   warps 2..9  epilogue (two warps per TMEM lane quadrant, alternating 32-column chunks):
//               tcgen05.ld the accumulator (lane == output row), bias / WSiLU / chunk-add pairing /
//               residuals / per-channel scale, split back into S3 planes; residual loads and
//               output stores are staged through shared memory so they are coalesced.  TMEM is double buffered (column bases 0 and 256, each
//               holding the main accumulator at +0 and the small-terms one at +128) so the
//               epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
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

static int encode3d(void* out, const void* base, uint64_t d0, uint64_t d1, uint64_t s1_bytes,
                    uint64_t s2_bytes, uint32_t box1) {
  auto fn = get_encode();
  if (!fn) {
    snprintf(g_umma_err, sizeof g_umma_err, "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t dims[3] = {d0, d1, 3};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {64, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_umma_err, sizeof g_umma_err,
             "cuTensorMapEncodeTiled failed (%d): base=%p dims=%llu,%llu strides=%llu,%llu box1=%u",
             (int)r, base, (unsigned long long)d0, (unsigned long long)d1,
             (unsigned long long)s1_bytes, (unsigned long long)s2_bytes, box1);
    return -1;
  }
  return 0;
}

int make_tmap_act(void* tmap_out, View a, long long M) {
  return encode3d(tmap_out, a.p, (uint64_t)a.C, (uint64_t)M, (uint64_t)a.ld * 2, (uint64_t)a.ps * 2, 128);
}
int make_tmap_weight(void* tmap_out, const GemmW& w) {
  return encode3d(tmap_out, w.w, (uint64_t)w.Kld, (uint64_t)w.Npad, (uint64_t)w.Kld * 2,
                  (uint64_t)w.Npad * w.Kld * 2, (uint32_t)w.BN);
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 = 1 [16,30) | SBO>>4 = 64 (8 rows x 128 B) [32,46) | version 1 [46,48)
// | layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct UmmaParams {
  long long M;
  int m_tiles, n_tiles, k_blocks;
  int BN, nsplit, stages;
  int* err;
};

constexpr int kATileBytes = 128 * 64 * 2;   // one plane of a 128 x 64 bf16 tile
constexpr int kThreads = 64 + 32 * 8;   // TMA warp, MMA warp, 8 epilogue warps

// ---- epilogue of one 32-column chunk, executed by a whole warp (lane == accumulator row) ----
// Global traffic is staged through a per-warp shared-memory tile (32 rows x 64 B payload, rows
// padded to 80 B so both the row-wise and the 8-rows-x-4-slots access patterns are conflict-free or
// 2-way): a warp-level access then covers 8 rows x 64 contiguous bytes instead of 32 rows x 16 B,
// which cuts the L1 wavefronts per tile 4x (the first version of this kernel was LSU-bound there).
constexpr int kStageRowBytes = 80;
constexpr int kStageBytes = 32 * kStageRowBytes;
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

// t[i] (+)= plane values of the 32 columns [dcol, dcol+32) of this lane's row, loaded coalesced.
__device__ __forceinline__ void staged_load_plane(const View& src, int plane, long long drow, bool row_ok,
                                                  int dcol, int ncols, uint32_t stage, int lane, float* t,
                                                  bool first) {
  const int slot = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    const long long rr = __shfl_sync(0xffffffffu, drow, r);
    const bool ok = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, r) && (slot * 8 < ncols);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ok) val = *reinterpret_cast<const uint4*>(src.p + plane * src.ps + rr * src.ld + dcol + slot * 8);
    st_shared_v4(stage + r * kStageRowBytes + slot * 16, val);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 q = ld_shared_v4(stage + lane * kStageRowBytes + j * 16);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = bf16lo(u[k]), hi = bf16hi(u[k]);
      t[8 * j + 2 * k] = first ? lo : add_rn(t[8 * j + 2 * k], lo);
      t[8 * j + 2 * k + 1] = first ? hi : add_rn(t[8 * j + 2 * k + 1], hi);
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void epilogue_chunk(const Epi& e, long long m, bool row_ok, int n0,
                                               const uint32_t* r0, const uint32_t* r1, uint32_t stage,
                                               int lane) {
  // destination of packed columns [n0, n0+32) (PACK_PAIR: [n0, n0+32) + partners [n0+32, n0+64))
  long long drow = m;
  int dcol = n0, limit = e.n_out;
  if (e.pack == PACK_PAIR) {
    dcol = (n0 >> 6) * 32;
  } else if (e.pack == PACK_SHUF2) {
    int g = n0 / e.Cg_pad;
    dcol = n0 - g * e.Cg_pad;
    limit = e.Cg;
    int w = (int)(m % e.W);
    long long t = m / e.W;
    int h = (int)(t % e.H);
    long long b = t / e.H;
    drow = (b * (2 * e.H) + (2 * h + (g >> 1))) * (2LL * e.W) + (2 * w + (g & 1));
  }
  if (dcol >= limit) return;                       // warp-uniform
  const int ncols = min(32, limit - dcol);         // multiple of 8, warp-uniform
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]);
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(e.bias + n0 + i);
      v[i] = add_rn(v[i], b.x); v[i + 1] = add_rn(v[i + 1], b.y);
      v[i + 2] = add_rn(v[i + 2], b.z); v[i + 3] = add_rn(v[i + 3], b.w);
    }
  }
  if (e.act != ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = apply_act(v[i], e.act);
  }
  if (e.pack == PACK_PAIR) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.bias) b = *reinterpret_cast<const float4*>(e.bias + n0 + 32 + i);
      v[i] = add_rn(v[i], apply_act(add_rn(__uint_as_float(r1[i]), b.x), e.act));
      v[i + 1] = add_rn(v[i + 1], apply_act(add_rn(__uint_as_float(r1[i + 1]), b.y), e.act));
      v[i + 2] = add_rn(v[i + 2], apply_act(add_rn(__uint_as_float(r1[i + 2]), b.z), e.act));
      v[i + 3] = add_rn(v[i + 3], apply_act(add_rn(__uint_as_float(r1[i + 3]), b.w), e.act));
    }
  }
  // residuals: x = (lo + mid) + hi exactly as join3, then v += x
  if (e.res1.p) {
    float t[32];
    staged_load_plane(e.res1, 2, drow, row_ok, dcol, ncols, stage, lane, t, true);
    staged_load_plane(e.res1, 1, drow, row_ok, dcol, ncols, stage, lane, t, false);
    staged_load_plane(e.res1, 0, drow, row_ok, dcol, ncols, stage, lane, t, false);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = add_rn(v[i], t[i]);
  }
  if (e.res2.p) {
    float t[32];
    staged_load_plane(e.res2, 2, drow, row_ok, dcol, ncols, stage, lane, t, true);
    staged_load_plane(e.res2, 1, drow, row_ok, dcol, ncols, stage, lane, t, false);
    staged_load_plane(e.res2, 0, drow, row_ok, dcol, ncols, stage, lane, t, false);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = add_rn(v[i], t[i]);
  }
  if (e.scale) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (i < ncols) {
        float4 s4 = *reinterpret_cast<const float4*>(e.scale + dcol + i);
        v[i] = mul_rn(v[i], s4.x); v[i + 1] = mul_rn(v[i + 1], s4.y);
        v[i + 2] = mul_rn(v[i + 2], s4.z); v[i + 3] = mul_rn(v[i + 3], s4.w);
      }
    }
  }
  if (e.do_clamp) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i], e.clamp_lo), e.clamp_hi);
  }
  if (e.out_f32 && row_ok) {                       // fp32 rows (2 launches per frame): direct stores
    float* d = e.out_f32 + drow * e.ld_f32 + dcol;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      if (i < ncols) *reinterpret_cast<float4*>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  if (e.out.p) {
    uint32_t ph[16], pm[16], pl[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      bf16 h0, m0, l0, h1, m1, l1;
      split3(v[2 * i], h0, m0, l0);
      split3(v[2 * i + 1], h1, m1, l1);
      ph[i] = pack_bf16(h0, h1);
      pm[i] = pack_bf16(m0, m1);
      pl[i] = pack_bf16(l0, l1);
    }
    const int slot = lane & 3;
#pragma unroll
    for (int plane = 0; plane < 3; ++plane) {
      const uint32_t* q = plane == 0 ? ph : (plane == 1 ? pm : pl);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(stage + lane * kStageRowBytes + j * 16, make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2);
        const long long rr = __shfl_sync(0xffffffffu, drow, r);
        const bool ok = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, r) && (slot * 8 < ncols);
        const uint4 val = ld_shared_v4(stage + r * kStageRowBytes + slot * 16);
        if (ok) *reinterpret_cast<uint4*>(e.out.p + plane * e.out.ps + rr * e.out.ld + dcol + slot * 8) = val;
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
k_gemm_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const Epi e, const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t wTileBytes = (uint32_t)p.BN * 128u;
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
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull(b), 1);
      mbar_init(bar_tempty(b), kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    const uint32_t ncols = 512;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmemSlot),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmemSlot) : "memory");

  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_idx = (tile / p.n_tiles) * 128, n_idx = (tile % p.n_tiles) * p.BN;
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(bar_empty(s), ph ^ 1, p.err, 1);
          mbar_expect_tx(bar_full(s), stageBytes);
          const uint32_t sa = base + s * stageBytes;
          const uint32_t sw = sa + p.nsplit * kATileBytes;
          for (int pl = 0; pl < p.nsplit; ++pl) {
            tma_load_3d(sa + pl * kATileBytes, &tmA, kb * 64, m_idx, pl, bar_full(s));
            tma_load_3d(sw + pl * wTileBytes, &tmW, kb * 64, n_idx, pl, bar_full(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major both,
      // N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const int nterms = (p.nsplit == 3) ? 6 : 1;
      // term 0 = hi*hi -> main accumulator; terms 1..5 (smallest first) -> second accumulator
      const int ta[6] = {0, 0, 2, 1, 0, 1};
      const int tw[6] = {0, 2, 0, 1, 1, 0};
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
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
            for (int t = 0; t < nterms; ++t) {
              const uint64_t ad = make_desc(sa + ta[t] * kATileBytes + ks * 32);
              const uint64_t bd = make_desc(sw + tw[t] * wTileBytes + ks * 32);
              if (t == 0) tc_mma(d_tmem, ad, bd, idesc, (kb | ks) ? 1u : 0u);
              else tc_mma(d_tmem + 128u, ad, bd, idesc, (kb | ks | (t - 1)) ? 1u : 0u);
            }
          }
          tc_commit(bar_empty(s));
        }
        tc_commit(bar_tfull(buf));
      }
    }
  } else {
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;           // two warps per quadrant split the column chunks
    const uint32_t stage = stageBase + (uint32_t)(warp - 2) * kStageBytes;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      const uint32_t tph = (tcount >> 1) & 1;
      const long long m = (long long)(tile / p.n_tiles) * 128 + quad * 32 + lane;
      const int n_idx = (tile % p.n_tiles) * p.BN;
      const bool row_ok = m < p.M;
      mbar_wait(bar_tfull(buf), tph, p.err, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * 256u;
      const bool two_acc = p.nsplit == 3;
      auto load_chunk = [&](int c, uint32_t* r) {     // accumulator columns [c, c+32) of this row
        tc_ld32(taddr + c, r);
        if (two_acc) {
          uint32_t s2[32];
          tc_ld32(taddr + 128u + c, s2);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            r[i] = __float_as_uint(add_rn(__uint_as_float(r[i]), __uint_as_float(s2[i])));
        } else {
          tc_wait_ld();
        }
      };
      if (e.pack == PACK_PAIR) {
        for (int c0 = half * 64; c0 < p.BN; c0 += 128) {
          uint32_t r0[32], r1[32];
          load_chunk(c0, r0);
          load_chunk(c0 + 32, r1);
          epilogue_chunk(e, m, row_ok, n_idx + c0, r0, r1, stage, lane);
        }
      } else {
        for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
          uint32_t r0[32];
          load_chunk(c0, r0);
          epilogue_chunk(e, m, row_ok, n_idx + c0, r0, r0, stage, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    const uint32_t ncols = 512;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols)
                 : "memory");
  }
}

int gemm_umma(const void* tmapA, const GemmW& w, const Epi& e, long long M, int K, int nsplit,
              cudaStream_t st) {
  static int smem_max = 0;
  static int* d_err = nullptr;
  if (!smem_max) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (cudaFuncSetAttribute(k_gemm_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max) !=
        cudaSuccess) {
      snprintf(g_umma_err, sizeof g_umma_err, "cudaFuncSetAttribute(max dynamic smem=%d) failed", smem_max);
      smem_max = 0;
      return -1;
    }
    cudaMalloc(&d_err, sizeof(int));
    cudaMemset(d_err, 0, sizeof(int));
  }
  if (!w.tmap || (w.BN % 32) || w.BN > (nsplit == 3 ? 128 : 256) || (e.pack == PACK_PAIR && (w.BN % 64))) {
    snprintf(g_umma_err, sizeof g_umma_err, "gemm_umma: unsupported weight tiling BN=%d", w.BN);
    return -1;
  }
  UmmaParams p;
  p.M = M;
  p.m_tiles = (int)((M + 127) / 128);
  p.n_tiles = (w.ncols + w.BN - 1) / w.BN;
  p.k_blocks = (K + 63) / 64;
  p.BN = w.BN;
  p.nsplit = nsplit;
  const int stage_bytes = nsplit * (kATileBytes + w.BN * 128);
  const int fixed = 1024 + 256 + kEpiWarps * kStageBytes;   // alignment slack + barriers + epilogue staging
  int stages = (smem_max - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 1) {
    snprintf(g_umma_err, sizeof g_umma_err, "gemm_umma: stage of %d bytes does not fit", stage_bytes);
    return -1;
  }
  p.stages = stages;
  p.err = d_err;
  const int smem = stages * stage_bytes + fixed;
  int grid = p.m_tiles * p.n_tiles;
  if (grid > num_sms()) grid = num_sms();
  CUtensorMap ta, tw;
  memcpy(&ta, tmapA, sizeof ta);
  memcpy(&tw, w.tmap, sizeof tw);
  (note_launch(), k_gemm_umma)<<<grid, kThreads, smem, st>>>(ta, tw, e, p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    snprintf(g_umma_err, sizeof g_umma_err, "k_gemm_umma launch: %s", cudaGetErrorString(err));
    return -1;
  }
  return 0;
}

}  // namespace dmc
