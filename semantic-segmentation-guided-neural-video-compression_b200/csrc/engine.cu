// Host side of the DMC engine: builds, per (variant, B, H, W), the static launch program of one
// frame -- buffers, packed weights, TMA descriptors and the ordered list of kernel launches --
// and exposes it through the C ABI of include/dmc_b200.h.
//
// The program follows the reference dataflow:
//   old          src/models/video_model.py:338-388
//   performance  src/refactor/seg_video_model.py:301-365
//   fast         src/refactor/seg_video_model_fast.py:328-411
//   mask_prop    src/refactor/mask_prop_seg_video_model.py:331-417
//   intra        src/models/image_model.py:205-261
// Channel concatenations never exist as copies: producers write straight into column slices
// of a wider S3 buffer (a View), which the consumer reads as one operand.
#include <cuda.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dmc_b200.h"
#include "kernels.h"

using namespace dmc;

namespace {

std::string g_create_error;

[[noreturn]] void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw std::runtime_error(buf);
}
#define CUDA_OK(call)                                                                       \
  do {                                                                                      \
    cudaError_t err__ = (call);                                                             \
    if (err__ != cudaSuccess)                                                               \
      fail("%s failed: %s (chain kernel wait code %d)", #call, cudaGetErrorString(err__),   \
           gemm_s3_trap_code());                                                            \
  } while (0)

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Act {          // an S3 tensor with its geometry (rows = B*H*W)
  View v{nullptr, 0, 0, 0};
  int B = 0, H = 0, W = 0;
  long long M() const { return (long long)B * H * W; }
};

struct Conv {         // dense convolution lowered to a contraction
  GemmW g;
  int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
  CUtensorMap tmap, tmap_half, tmap_s3, tmap_s3_hi;
};
struct DW {
  float* w9c = nullptr;
  float* bias = nullptr;
  int C = 0;
};
struct DCB {          // layers.py:43-79
  Conv* adaptor = nullptr;
  Conv *dc0 = nullptr, *dc3 = nullptr, *ffn0 = nullptr, *ffn2 = nullptr;
  DW* dw = nullptr;
  int cin = 0, cout = 0;
};

struct WSlot {
  std::string key;
  std::vector<int64_t> shape;
  std::function<void(const float*, cudaStream_t)> load;
  bool set = false;
};

struct EpiSpec {
  int act = ACT_NONE;
  const Act* res1 = nullptr;
  const Act* res2 = nullptr;
  const float* scale_table = nullptr;   // (72, C) table, row chosen by qp at launch
  int scale_C = 0;
  float* out_f32 = nullptr;
  int ld_f32 = 0;
  bool clamp01 = false;
  int nsplit = 3;
};

}  // namespace

// Every C-ABI entry that touches an engine runs on the device the engine was created on, whatever the caller's current
// device is (torch: model.to("cuda:1") without torch.cuda.set_device(1)).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev && dev >= 0) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// Process-wide pool of training workspaces (dmc_dcb_train_*): blocks of the same geometry share every activation /
// scratch buffer -- a backward pass recomputes what it needs from x, so nothing has to survive in them between calls --
// and keep only their packed weights to themselves.  Entries are keyed by device, geometry and allocation order and
// counted: the memory goes when the last handle using it does.  Handles that share a workspace must be used from one
// stream at a time (they are: torch's current stream).
struct TrainPool {
  struct Buf { void* p; size_t bytes; int refs; };
  std::map<std::string, Buf> bufs;
  std::mutex mu;
  static TrainPool& get() {
    static TrainPool pool;
    return pool;
  }
};

struct dmc_engine {
  int variant = 0, B = 0, H = 0, W = 0, flags = 0;
  int device = -1;                 // CUDA ordinal the buffers, tensor maps and chains belong to
  std::string error;
  bool finalized = false;

  // per-call state read by the launch closures
  struct Cur {
    const float *x = nullptr, *mask = nullptr, *dpb_frame = nullptr, *dpb_feature = nullptr;
    float *x_hat = nullptr, *feature = nullptr, *bpp3 = nullptr, *mask_pred = nullptr;
    int32_t* finite = nullptr;
    int qp = 0;
  } cur;
  bool cur_after_i = true;
  float* logits_full = nullptr;      // mask_prop: predictor logits at full resolution when the caller passes no mask_pred

  std::vector<void*> allocs;
  std::vector<std::unique_ptr<Conv>> convs;
  std::vector<std::unique_ptr<DW>> dws;
  std::vector<std::unique_ptr<DCB>> dcbs;
  std::vector<std::unique_ptr<Act>> acts;
  std::vector<std::unique_ptr<CUtensorMap>> tmaps;
  std::vector<WSlot> slots;
  std::map<std::string, int> slot_index;
  std::map<std::string, Act> scratch;
  std::map<std::string, Act> taps;               // S3 taps
  struct F32Tap { const float* p; int B, C, H, W; };
  std::map<std::string, F32Tap> ftaps;           // fp32 row-major [M, C] taps

  // One launch (or a few) of the frame program, tagged with the PHASE of the codec it belongs to.  forward() runs
  // every phase; the decoder entry points (dmc_decode_*) run the phases a decoder has, with the entropy decoder of
  // the caller between them.
  enum Phase { PH_FEAT = 0, PH_ENC = 1, PH_PRIOR = 2, PH_STEP0 = 3, PH_SP1 = 4, PH_STEP1 = 5, PH_SP2 = 6, PH_STEP2 = 7,
               PH_SP3 = 8, PH_STEP3 = 9, PH_FIN = 10, PH_DEC = 11, PH_RATE = 12 };
  struct Op {
    std::function<void(cudaStream_t)> fn;
    int phase;
    template <class F>
    Op(F f) : fn(std::move(f)), phase(tl_phase()) {}
    void operator()(cudaStream_t st) const { fn(st); }
  };
  static int& tl_phase() {
    static thread_local int p = PH_FEAT;
    return p;
  }
  void set_phase(int ph) {
    flush_chain();               // (a staged chain belongs to the phase it was staged in)
    tl_phase() = ph;
  }
  std::vector<Op> prog_head_i, prog_head_p, prog_common;
  std::vector<Op>* prog = nullptr;               // where the builder appends

  double* bits_y = nullptr;
  double* bits_z = nullptr;
  float* bpp_scratch = nullptr;

  // ---- caller-owned tensors reach the kernels through device-resident slots (kernels.h: IoSlots), written by one
  // tiny launch before every forward: the launch program itself never contains a caller pointer, so it can be
  // captured once per (after_i, qp, mask present) as a CUDA graph and replayed on any tensors.
  IoSlots* io_dev = nullptr;
  int* finite_dev = nullptr;
  const void* const* slot(size_t off) const {
    return reinterpret_cast<const void* const*>(reinterpret_cast<const char*>(io_dev) + off);
  }
#define IO_SLOT(field) slot(offsetof(IoSlots, field))
  struct GraphEntry { cudaGraphExec_t exec = nullptr; long long launches = 0; };
  std::map<uint64_t, GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;
  bool graphs_on = graphs_default();
  std::string graph_note;                       // why graphs were switched off for this engine (if they were)
  static bool graphs_default() {
    const char* v = getenv("DMC_GRAPH");           // DMC_GRAPH=0: launch every kernel of every forward directly
    return !(v && v[0] == '0');
  }
  // Runs `body` (which only enqueues work on the stream it is given) as a CUDA graph keyed by `key`: captured on the
  // engine's own stream the first time (torch's default stream is the legacy stream, which cannot be captured),
  // instantiated, and from then on launched with one call.  Any failure switches graphs off for this engine and
  // falls back to direct launches.
  template <class Body>
  void run_graph(uint64_t key, cudaStream_t st, Body body) {
    if (graphs_on && !profile) {
      auto it = graphs.find(key);
      if (it == graphs.end()) {
        GraphEntry ge;
        cudaGraph_t g = nullptr;
        const long long l0 = launch_count();
        bool ok = true;
        if (!cap_stream) ok = cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking) == cudaSuccess;
        if (ok) ok = cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          try {
            body(cap_stream);
          } catch (const std::exception& ex) {
            graph_note = ex.what();
            ok = false;
          }
          if (cudaStreamEndCapture(cap_stream, &g) != cudaSuccess || !g) ok = false;
        }
        if (ok) ok = cudaGraphInstantiate(&ge.exec, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        ge.launches = launch_count() - l0;
        if (!ok) {
          if (graph_note.empty()) graph_note = cudaGetErrorString(cudaGetLastError());
          cudaGetLastError();
          graphs_on = false;
          add_launches(-ge.launches);          // counted while capturing, never run
        } else {
          it = graphs.emplace(key, ge).first;
          add_launches(-ge.launches);          // (added back below, like every replay)
        }
      }
      if (graphs_on) {
        CUDA_OK(cudaGraphLaunch(it->second.exec, st));
        add_launches(it->second.launches);
        return;
      }
    }
    body(st);
  }

  // measurement (dmc_profile_*): events around every contraction launch
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  double prof_flops = 0, prof_issued = 0;
  void prof_begin(cudaStream_t st, double flops, double issued) {
    cudaEvent_t a, b;
    CUDA_OK(cudaEventCreate(&a));
    CUDA_OK(cudaEventCreate(&b));
    CUDA_OK(cudaEventRecord(a, st));
    prof_events.emplace_back(a, b);
    prof_flops += flops;
    prof_issued += issued;
  }
  void prof_end(cudaStream_t st) { CUDA_OK(cudaEventRecord(prof_events.back().second, st)); }

  ~dmc_engine() {
    DeviceGuard g(device);
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second.exec);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    for (S3Chain* c : chains) s3_chain_destroy(c);
    for (void* p : allocs) cudaFree(p);
    if (!pool_keys.empty()) {
      TrainPool& pool = TrainPool::get();
      std::lock_guard<std::mutex> lk(pool.mu);
      for (const std::string& k : pool_keys) {
        auto it = pool.bufs.find(k);
        if (it != pool.bufs.end() && --it->second.refs == 0) {
          cudaFree(it->second.p);
          pool.bufs.erase(it);
        }
      }
    }
  }

  bool simt() const { return flags & DMC_FLAG_SIMT_GEMM; }
  bool use_s3 = s3_default();    // route qualifying contractions to the specialised CTA-pair kernel
  static bool s3_default() {
    const char* v = getenv("DMC_GEMM_S3");          // DMC_GEMM_S3=0 keeps every launch on the general kernel (A/B runs)
    return !(v && v[0] == '0');
  }
  bool keep_taps() const { return flags & DMC_FLAG_KEEP_TAPS; }

  // ------------------------------------------------------------ memory
  // `fresh` (optional) tells whether the memory was allocated by this call (pooled buffers may be handed out again)
  std::string pool_prefix;             // non-empty: allocations come from the shared training pool (TrainPool)
  std::vector<std::string> pool_keys;
  void* dalloc(size_t bytes, bool* fresh = nullptr) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (fresh) *fresh = true;
    if (!pool_prefix.empty()) {
      TrainPool& pool = TrainPool::get();
      std::lock_guard<std::mutex> lk(pool.mu);
      const std::string key = pool_prefix + "#" + std::to_string(pool_keys.size()) + ":" + std::to_string(bytes);
      auto it = pool.bufs.find(key);
      if (it != pool.bufs.end()) {
        ++it->second.refs;
        if (fresh) *fresh = false;
        p = it->second.p;
      } else {
        CUDA_OK(cudaMalloc(&p, bytes));
        CUDA_OK(cudaMemset(p, 0, bytes));
        pool.bufs[key] = TrainPool::Buf{p, bytes, 1};
      }
      pool_keys.push_back(key);
      return p;
    }
    CUDA_OK(cudaMalloc(&p, bytes));
    allocs.push_back(p);
    return p;
  }
  Act new_act(int b, int h, int w, int C) {
    Act a;
    a.B = b; a.H = h; a.W = w;
    // tile-blocked planes (common.cuh): [ceil(C/16)][rows padded to 256][16]; padding rows / columns start as
    // zeros (padding columns are only ever rewritten with zeros, padding rows never reach a valid row)
    long long M = a.M();
    long long Mp = (M + 255) / 256 * 256;
    long long bs = Mp * 16;
    long long ps = bs * ((C + 15) / 16);
    bool fresh = true;
    a.v.p = (h16*)dalloc((size_t)ps * kPlanes * sizeof(h16), &fresh);
    a.v.ps = ps; a.v.bs = bs; a.v.C = C;
    if (fresh && pool_prefix.empty()) CUDA_OK(cudaMemset(a.v.p, 0, (size_t)ps * kPlanes * sizeof(h16)));   // (pool: zeroed once)
    return a;
  }
  // scratch buffers are shared by every block of the same geometry (one stream, in-order)
  Act get_scratch(const std::string& tag, int b, int h, int w, int C) {
    char key[160];
    snprintf(key, sizeof key, "%s:%d:%d:%d:%d", tag.c_str(), b, h, w, C);
    auto it = scratch.find(key);
    if (it != scratch.end()) return it->second;
    Act a = new_act(b, h, w, C);
    scratch[key] = a;
    return a;
  }
  std::map<std::string, float*> scratch_f32;
  float* get_scratch_f32(const std::string& tag, size_t n) {
    char key[160];
    snprintf(key, sizeof key, "%s:%zu", tag.c_str(), n);
    auto it = scratch_f32.find(key);
    if (it != scratch_f32.end()) return it->second;
    float* p = new_f32(n);
    scratch_f32[key] = p;
    return p;
  }
  static Act slice(const Act& a, int c0, int C) {
    if (c0 % 16) fail("slice: column offset %d is not a multiple of 16", c0);
    Act s = a;
    s.v.p = a.v.p + (long long)(c0 / 16) * a.v.bs;
    s.v.C = C;
    return s;
  }
  float* new_f32(size_t n) { return (float*)dalloc(n * sizeof(float)); }

  // ------------------------------------------------------------ weights
  void add_slot(const std::string& key, std::vector<int64_t> shape,
                std::function<void(const float*, cudaStream_t)> load) {
    slot_index[key] = (int)slots.size();
    slots.push_back(WSlot{key, std::move(shape), std::move(load), false});
  }
  // raw fp32 copy of a tensor (per-QP tables, tiny convs)
  float* add_table(const std::string& key, std::vector<int64_t> shape) {
    size_t n = 1;
    for (auto s : shape) n *= (size_t)s;
    float* d = new_f32(n);
    add_slot(key, shape, [d, n](const float* src, cudaStream_t st) {
      CUDA_OK(cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    });
    return d;
  }
  Conv* add_conv(const std::string& key, int cin, int cout, int k, int stride, int pad,
                 int pack = PACK_PLAIN) {
    convs.emplace_back(new Conv());
    Conv* c = convs.back().get();
    c->cin = cin; c->cout = cout; c->k = k; c->stride = stride; c->pad = pad;
    GemmW& g = c->g;
    g.N = cout; g.K = cin * k * k; g.pack = pack;
    if (pack == PACK_PAIR) {
      int c2 = cout / 2;
      g.ncols = (c2 + 31) / 32 * 64;
      g.BN = (g.ncols > 64) ? 128 : 64;      // an odd number of 64-column groups leaves half a tile of zero rows
    } else if (pack == PACK_SHUF2) {
      g.Cg = cout / 4;
      g.Cg_pad = round_up(g.Cg, 32);
      g.ncols = 4 * g.Cg_pad;
      g.BN = (g.ncols % 128 == 0) ? 128 : ((g.ncols % 96 == 0) ? 96 : 64);
    } else {
      g.ncols = cout;
      // N tile <= 128: two fp32 accumulators per tile, double buffered, fill the 512 TMEM columns
      if (cout % 128 == 0) g.BN = 128;
      else if (cout % 96 == 0) g.BN = 96;
      else if (cout <= 64) g.BN = 64;
      else g.BN = (cout > 96) ? 128 : 96;
    }
    g.Npad = round_up(round_up(g.ncols, g.BN), 64);
    g.Kld = round_up(g.K, 64);
    g.w = (h16*)dalloc((size_t)kPlanes * g.Npad * g.Kld * sizeof(h16));
    if (!simt()) g.wb = (h16*)dalloc((size_t)kPlanes * g.Npad * g.Kld * sizeof(h16));
    g.bias = new_f32(g.Npad);
    g.tmap = &c->tmap;
    g.tmap_half = &c->tmap_half;
    g.tmap_s3 = &c->tmap_s3;
    g.tmap_s3_hi = &c->tmap_s3_hi;
    if (!simt()) {
      if (make_tmap_s3_weight(&c->tmap_s3, g, kPlanes) != 0 || make_tmap_s3_weight(&c->tmap_s3_hi, g, 1) != 0)
        fail("%s: %s", key.c_str(), gemm_s3_last_error());
      if (make_tmap_weight(&c->tmap, g, g.BN) != 0) fail("%s: %s", key.c_str(), umma_last_error());
      if (make_tmap_weight(&c->tmap_half, g, g.BN / 2) != 0) fail("%s: %s", key.c_str(), umma_last_error());
    }
    add_slot(key + ".weight", {cout, cin, k, k}, [c](const float* src, cudaStream_t st) {
      pack_gemm_weight(src, c->cout, c->cin, c->k, c->k, c->g, st);
    });
    add_slot(key + ".bias", {cout}, [c](const float* src, cudaStream_t st) {
      pack_gemm_bias(src, c->cout, c->g, st);
    });
    return c;
  }
  DW* add_dw(const std::string& key, int C) {
    dws.emplace_back(new DW());
    DW* d = dws.back().get();
    d->C = C;
    d->w9c = new_f32((size_t)9 * C);
    d->bias = new_f32(C);
    add_slot(key + ".weight", {C, 1, 3, 3},
             [d](const float* src, cudaStream_t st) { pack_dw_weight(src, d->w9c, d->C, st); });
    add_slot(key + ".bias", {C}, [d](const float* src, cudaStream_t st) {
      CUDA_OK(cudaMemcpyAsync(d->bias, src, d->C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    });
    return d;
  }
  DCB* add_dcb(const std::string& key, int cin, int cout, bool force_adaptor = false) {
    dcbs.emplace_back(new DCB());
    DCB* b = dcbs.back().get();
    b->cin = cin; b->cout = cout;
    if (cin != cout || force_adaptor) b->adaptor = add_conv(key + ".adaptor", cin, cout, 1, 1, 0);
    b->dc0 = add_conv(key + ".dc.0", cout, cout, 1, 1, 0);
    b->dw = add_dw(key + ".dc.2", cout);
    b->dc3 = add_conv(key + ".dc.3", cout, cout, 1, 1, 0);
    b->ffn0 = add_conv(key + ".ffn.0", cout, cout * 4, 1, 1, 0, PACK_PAIR);
    b->ffn2 = add_conv(key + ".ffn.2", cout * 2, cout, 1, 1, 0);
    return b;
  }

  // ------------------------------------------------------------ op builders
  void op(Op f) {
    flush_chain();
    prog->push_back(std::move(f));
  }
  void set_prog(std::vector<Op>* p) {
    flush_chain();
    prog = p;
  }

  // ---- chain staging (gemm_s3.cu): stages collected by gemm(), turned into one launch by flush_chain()
  std::vector<S3StageDesc> pend;
  long long pend_M = 0;
  double pend_flops = 0, pend_issued = 0;
  std::vector<S3Chain*> chains;
  bool chain_layers = chain_default();
  static bool chain_default() {
    const char* v = getenv("DMC_GEMM_CHAIN");       // DMC_GEMM_CHAIN=0: one launch per layer (A/B runs)
    return !(v && v[0] == '0');
  }
  void flush_chain() {
    if (pend.empty()) return;
    S3Chain* c = s3_chain_create(pend.data(), (int)pend.size(), pend_M);
    if (!c) fail("s3_chain_create: %s", gemm_s3_last_error());
    chains.push_back(c);
    dmc_engine* self = this;
    const double fl = pend_flops, issued = pend_issued;
    prog->push_back([self, c, fl, issued](cudaStream_t st) {
      if (self->profile) self->prof_begin(st, fl, issued);
      if (s3_chain_launch(c, self->cur.qp, st) != 0) fail("s3_chain_launch: %s", gemm_s3_last_error());
      if (self->profile) self->prof_end(st);
    });
    pend.clear();
    pend_flops = pend_issued = 0;
  }

  void tap(const std::string& name, const Act& a) {
    if (!keep_taps()) return;
    Act copy = new_act(a.B, a.H, a.W, a.v.C);
    Act src = a;
    long long M = a.M();
    op([src, copy, M](cudaStream_t st) { copy_view(src.v, copy.v, M, st); });
    taps[name] = copy;
  }

  // contraction of `in` (rows x K) with conv weight `c` into `out` (S3) or spec.out_f32
  void gemm(const Act& in, Conv* c, const Act* out, EpiSpec spec, int shuf_H = 0, int shuf_W = 0) {
    const GemmW* g = &c->g;
    if (in.v.C != g->K) fail("gemm: operand has %d columns, weight expects K=%d", in.v.C, g->K);
    Epi e;
    memset(&e, 0, sizeof e);
    e.bias = g->bias;
    e.act = spec.act;
    e.pack = g->pack;
    if (spec.res1) e.res1 = spec.res1->v;
    if (spec.res2) e.res2 = spec.res2->v;
    if (out) e.out = out->v;
    e.out_f32 = spec.out_f32;
    e.ld_f32 = spec.ld_f32;
    e.n_out = (g->pack == PACK_PAIR) ? g->N / 2 : (g->pack == PACK_SHUF2 ? g->Cg : g->N);
    e.H = shuf_H; e.W = shuf_W; e.Cg = g->Cg; e.Cg_pad = g->Cg_pad;
    if (spec.clamp01) { e.do_clamp = 1; e.clamp_lo = 0.f; e.clamp_hi = 1.f; }
    if (out && out->v.C < e.n_out) fail("gemm: destination has %d columns, need %d", out->v.C, e.n_out);
    long long M = in.M();
    View a = in.v;
    bool aligned = true, out_ok = true;      // blocked planes are always 512-byte aligned
    bool use_umma = !simt() && aligned && out_ok && (g->K % 8 == 0) && (e.n_out % 8 == 0) &&
                    (!spec.out_f32 || (spec.ld_f32 % 4 == 0));
    const float* table = spec.scale_table;
    int sc = spec.scale_C;
    int nsplit = spec.nsplit;
    dmc_engine* self = this;
    auto tma_ok = [](const View& v) { return (uintptr_t)v.p % 512 == 0; };
    if (use_umma && use_s3 && gemm_s3_supports(*g, e, nsplit) && (!out || tma_ok(out->v)) &&
        (!spec.res1 || tma_ok(spec.res1->v)) && (!spec.res2 || tma_ok(spec.res2->v))) {
      CUtensorMap* tm3[5];
      for (auto& t : tm3) {
        tmaps.emplace_back(new CUtensorMap());
        t = tmaps.back().get();
      }
      if (make_tmap_s3_act64(tm3[3], a, M) != 0) fail("gemm A map (64 rows): %s", gemm_s3_last_error());
      if (make_tmap_s3_act(tm3[0], a, M, nsplit != 1 ? kPlanes : 1) != 0) fail("gemm A map: %s", gemm_s3_last_error());
      if (spec.out_f32) {
        if (make_tmap_f32_rows(tm3[1], spec.out_f32, e.n_out, spec.ld_f32, M) != 0)
          fail("gemm fp32 out map: %s", gemm_s3_last_error());
      } else if (make_tmap_s3_rows(tm3[1], out->v, e.n_out, M) != 0) {
        fail("gemm out map: %s", gemm_s3_last_error());
      }
      if (spec.res1 && make_tmap_s3_rows(tm3[2], spec.res1->v, e.n_out, M) != 0)
        fail("gemm residual map: %s", gemm_s3_last_error());
      if (spec.res2 && make_tmap_s3_rows(tm3[4], spec.res2->v, e.n_out, M) != 0)
        fail("gemm second residual map: %s", gemm_s3_last_error());
      // consecutive qualifying contractions over the same rows become stages of one chain launch
      if (!pend.empty() && (pend_M != M || (int)pend.size() >= s3_chain_max_stages() || !chain_layers)) flush_chain();
      S3StageDesc d;
      d.tmA = tm3[0];
      d.tmA64 = tm3[3];
      d.w = g;
      d.e = e;
      d.tmOut = tm3[1];
      d.tmRes = spec.res1 ? tm3[2] : nullptr;
      d.tmRes2 = spec.res2 ? tm3[4] : nullptr;
      d.K = g->K;
      d.nsplit = nsplit;
      d.scale_table = table;
      d.scale_C = sc;
      pend.push_back(d);
      pend_M = M;
      pend_flops += 2.0 * (double)M * g->N * g->K;
      pend_issued += 2.0 * (double)M * g->N * g->K * (nsplit != 1 ? 3 : 1);
    } else if (use_umma) {
      tmaps.emplace_back(new CUtensorMap());
      CUtensorMap* tm = tmaps.back().get();
      if (make_tmap_act(tm, a, M) != 0) fail("gemm A map: %s", umma_last_error());
      int K = g->K;
      op([self, tm, g, e, M, K, nsplit, table, sc](cudaStream_t st) {
        Epi ee = e;
        if (table) ee.scale = table + (size_t)self->cur.qp * sc;
        double fl = 2.0 * (double)M * g->N * g->K;
        if (self->profile) self->prof_begin(st, fl, fl * (nsplit != 1 ? 3 : 1));
        if (gemm_umma(tm, *g, ee, M, K, nsplit, st) != 0) fail("gemm_umma: %s", umma_last_error());
        if (self->profile) self->prof_end(st);
      });
    } else {
      op([self, a, g, e, M, table, sc](cudaStream_t st) {
        Epi ee = e;
        if (table) ee.scale = table + (size_t)self->cur.qp * sc;
        double fl = 2.0 * (double)M * g->N * g->K;
        if (self->profile) self->prof_begin(st, fl, fl);
        gemm_simt(a, *g, ee, M, st);
        if (self->profile) self->prof_end(st);
      });
    }
  }

  // k x k convolution with stride/pad: im2col into a scratch S3 matrix, then the contraction
  void conv_kxk(const Act& in, Conv* c, const Act* out, EpiSpec spec, bool shuf2 = false) {
    int Ho = (in.H + 2 * c->pad - c->k) / c->stride + 1;
    int Wo = (in.W + 2 * c->pad - c->k) / c->stride + 1;
    Act col = get_scratch("im2col", in.B, Ho, Wo, c->k * c->k * c->cin);
    Act src = in;
    int k = c->k, s = c->stride, p = c->pad, cin = c->cin;
    op([src, col, k, s, p, Ho, Wo, cin](cudaStream_t st) {
      im2col(src.v, col.v, src.B, src.H, src.W, k, s, p, Ho, Wo, cin, 0, st);
    });
    gemm(col, c, out, spec, shuf2 ? Ho : 0, shuf2 ? Wo : 0);
  }

  // DepthConvBlock; `out` may alias `x` (every element is read before it is written by the same thread)
  void dcb(DCB* w, const Act& x, const Act& out, bool shortcut, const float* scale_table, int nsplit,
           float* out_f32 = nullptr, int ld_f32 = 0) {
    int C = w->cout;
    Act xin = x;
    EpiSpec s0;
    s0.nsplit = nsplit;
    if (w->adaptor) {
      xin = get_scratch("dcb_x", x.B, x.H, x.W, C);
      gemm(x, w->adaptor, &xin, s0);
    }
    // dc.0's output only feeds the depthwise conv: it stays fp32 rows (4 B/element, no S3 join in the consumer)
    float* t32 = get_scratch_f32("dcb_t32", (size_t)x.M() * C);
    Act t2 = get_scratch("dcb_t2", x.B, x.H, x.W, C);
    Act o1 = get_scratch("dcb_o1", x.B, x.H, x.W, C);
    Act u = get_scratch("dcb_u", x.B, x.H, x.W, 2 * C);
    EpiSpec s1 = s0;
    s1.act = ACT_WSILU;
    {
      EpiSpec sd = s1;
      sd.out_f32 = t32;
      sd.ld_f32 = C;
      gemm(xin, w->dc0, nullptr, sd);
    }
    DW* dw = w->dw;
    op([t32, C, t2, dw](cudaStream_t st) { dwconv3x3_f32(t32, C, dw->w9c, dw->bias, t2.v, t2.B, t2.H, t2.W, st); });
    EpiSpec s2 = s0;
    s2.res1 = &xin;
    gemm(t2, w->dc3, &o1, s2);
    gemm(o1, w->ffn0, &u, s1);
    EpiSpec s3 = s0;
    s3.res1 = &o1;
    if (shortcut) s3.res2 = &xin;
    s3.scale_table = scale_table;
    s3.scale_C = C;
    s3.out_f32 = out_f32;
    s3.ld_f32 = ld_f32;
    gemm(u, w->ffn2, out_f32 ? nullptr : &out, s3);
  }

  void build_p();
  void build_intra();
  void finalize(cudaStream_t st);
  void run(std::vector<Op>& p, cudaStream_t st, uint32_t phases = 0xffffffffu) {
    pdl_set_auto((long long)B * H * W <= (1LL << 20));   // (kernels.cu: launch-bound frames only)
    for (auto& f : p)
      if (phases >> f.phase & 1u) f(st);
  }
  // ---- decoder state (dmc_decode_*): which buffers the entropy decoder of the caller reads / fills
  PriorArgs dec_prior{};          // the prior kernels' arguments (step filled in per call)
  View dec_zhat{nullptr, 0, 0, 0};
  int dec_zC = 0, dec_zH = 0, dec_zW = 0;
  int dec_steps = 0;
};

// ====================================================================== P-frame program
void dmc_engine::build_p() {
  const bool refactor = variant != DMC_VARIANT_OLD;
  // the hyper path works on y replicate-padded to multiples of 4 (models/common_model.py:68-72): Hp16 x Wp16
  const int H8 = H / 8, W8 = W / 8, H16 = H / 16, W16 = W / 16, Hp16 = (H16 + 3) / 4 * 4, Wp16 = (W16 + 3) / 4 * 4,
            H32 = Hp16 / 2, W32 = Wp16 / 2, H64 = Hp16 / 4, W64 = Wp16 / 4;
  const bool padded = Hp16 != H16 || Wp16 != W16;
  const int CD = 256, CY = 128, CZ = 128, CR = 320, QP = 72;
  const int ns = 3;
  // recon_generation_net never feeds a later symbol: plain bf16 operands unless the caller asks
  // for the fp32-grade product there too
  const int ns_recon = (flags & DMC_FLAG_RECON_SPLIT3) ? 3 : 1;
  dmc_engine* self = this;

  // ---- weights, in the reference's registration order (documentation only; lookup is by key)
  float* q_encoder = add_table("q_encoder", {QP, CD, 1, 1});
  float* q_decoder = add_table("q_decoder", {QP, CD, 1, 1});
  float* q_feature = add_table("q_feature", {QP, CD, 1, 1});
  float* q_recon = add_table("q_recon", {QP, CR, 1, 1});
  float* q_sft = (variant == DMC_VARIANT_PERFORMANCE) ? add_table("q_sft", {QP, CD, 1, 1}) : nullptr;
  float* bitparm[11];
  {
    const char* names[11] = {"f1.h", "f1.b", "f1.a", "f2.h", "f2.b", "f2.a",
                             "f3.h", "f3.b", "f3.a", "f4.h", "f4.b"};
    for (int i = 0; i < 11; ++i)
      bitparm[i] = add_table(std::string("bit_estimator_z.") + names[i], {QP, CZ, 1, 1});
  }
  DCB* fa_i = add_dcb("feature_adaptor_i", 192, CD);
  Conv* fa_p = add_conv("feature_adaptor_p", CD, CD, 1, 1, 0);
  DCB* fe1[2];
  DCB* fe2[4];
  for (int i = 0; i < 2; ++i) fe1[i] = add_dcb("feature_extractor.conv1." + std::to_string(i), CD, CD);
  for (int i = 0; i < 4; ++i) fe2[i] = add_dcb("feature_extractor.conv2." + std::to_string(i), CD, CD);
  Conv* enc_conv1 = add_conv("encoder.conv1", 192, CD, 1, 1, 0);
  DCB* enc_b[3];
  enc_b[0] = add_dcb("encoder.conv2.0", 2 * CD, CD);
  enc_b[1] = add_dcb("encoder.conv2.1", CD, CD);
  enc_b[2] = add_dcb(refactor ? "encoder.conv2.2" : "encoder.conv3", CD, CD);
  Conv* enc_down = add_conv("encoder.down", CD, CY, 3, 2, 1);
  DCB* he0 = add_dcb("hyper_encoder.conv.0", CY, CZ);
  Conv* he1d = add_conv("hyper_encoder.conv.1.down", CZ, CZ, 2, 2, 0);
  DCB* he1 = add_dcb("hyper_encoder.conv.1.conv", CZ, CZ);
  Conv* he2d = add_conv("hyper_encoder.conv.2.down", CZ, CZ, 2, 2, 0);
  DCB* he2 = add_dcb("hyper_encoder.conv.2.conv", CZ, CZ);
  Conv* hd0u = add_conv("hyper_decoder.conv.0.up.conv.0", CZ, CZ * 4, 1, 1, 0, PACK_SHUF2);
  DCB* hd0 = add_dcb("hyper_decoder.conv.0.conv", CZ, CZ);
  Conv* hd1u = add_conv("hyper_decoder.conv.1.up.conv.0", CZ, CZ * 4, 1, 1, 0, PACK_SHUF2);
  DCB* hd1 = add_dcb("hyper_decoder.conv.1.conv", CZ, CZ);
  DCB* hd2 = add_dcb("hyper_decoder.conv.2", CZ, CY);
  Conv* tpd = add_conv("temporal_prior_encoder.down", CD, 2 * CY, 2, 2, 0);
  DCB* tpc = add_dcb("temporal_prior_encoder.conv", 2 * CY, 2 * CY);
  DCB* pf[3];
  for (int i = 0; i < 3; ++i) pf[i] = add_dcb("y_prior_fusion.conv." + std::to_string(i), 3 * CY, 3 * CY);
  Conv* pf3 = add_conv("y_prior_fusion.conv.3", 3 * CY, 3 * CY, 1, 1, 0);
  DCB* sp0 = add_dcb("y_spatial_prior.conv.0", 4 * CY, 3 * CY);
  DCB* sp1 = add_dcb("y_spatial_prior.conv.1", 3 * CY, 3 * CY);
  Conv* sp2 = add_conv("y_spatial_prior.conv.2", 3 * CY, 2 * CY, 1, 1, 0);
  Conv* dec_up = add_conv("decoder.up.conv.0", CY, CD * 4, 3, 1, 1, PACK_SHUF2);
  DCB* dec_b[3];
  const std::string dpre = refactor ? "decoder.conv." : "decoder.conv1.";
  dec_b[0] = add_dcb(dpre + "0", 2 * CD, CD);
  dec_b[1] = add_dcb(dpre + "1", CD, CD);
  dec_b[2] = add_dcb(dpre + "2", CD, CD);
  Conv* dec_proj = add_conv(refactor ? "decoder.proj" : "decoder.conv2", CD, CD, 1, 1, 0);
  DCB* rec[4];
  rec[0] = add_dcb("recon_generation_net.conv.0", CD, CR);
  for (int i = 1; i < 4; ++i) rec[i] = add_dcb("recon_generation_net.conv." + std::to_string(i), CR, CR);
  Conv* rec_head = add_conv("recon_generation_net.head", CR, 192, 1, 1, 0);

  Conv* sft_conv1 = nullptr; DCB* sft_b[3] = {nullptr, nullptr, nullptr}; Conv* sft_down = nullptr;
  float *mf_w0 = nullptr, *mf_b0 = nullptr, *mf_w2 = nullptr, *mf_b2 = nullptr;
  float *me_w = nullptr, *me_b = nullptr, *mp4_w = nullptr, *mp4_b = nullptr;
  Conv *mp0 = nullptr, *mp2 = nullptr;
  if (variant == DMC_VARIANT_PERFORMANCE) {
    sft_conv1 = add_conv("mask_sft.conv1", 64, CD, 1, 1, 0);
    for (int i = 0; i < 3; ++i) sft_b[i] = add_dcb("mask_sft.conv2." + std::to_string(i), CD, CD);
    sft_down = add_conv("mask_sft.down", CD, 2 * CY, 3, 2, 1);
  }
  if (variant == DMC_VARIANT_FAST || variant == DMC_VARIANT_MASK_PROP) {
    mf_w0 = add_table("mask_film.net.0.weight", {16, 1, 3, 3});
    mf_b0 = add_table("mask_film.net.0.bias", {16});
    mf_w2 = add_table("mask_film.net.2.weight", {2 * CY, 16, 1, 1});
    mf_b2 = add_table("mask_film.net.2.bias", {2 * CY});
  }
  if (variant == DMC_VARIANT_MASK_PROP) {
    me_w = add_table("mask_predictor.mask_embed.weight", {CD, 1, 3, 3});
    me_b = add_table("mask_predictor.mask_embed.bias", {CD});
    mp0 = add_conv("mask_predictor.net.0", 3 * CD, CD / 4, 3, 1, 1);
    mp2 = add_conv("mask_predictor.net.2", CD / 4, CD / 4, 3, 1, 1);
    mp4_w = add_table("mask_predictor.net.4.weight", {1, CD / 4, 1, 1});
    mp4_b = add_table("mask_predictor.net.4.bias", {1});
  }

  // ---- persistent buffers
  bits_y = (double*)dalloc(sizeof(double) * B);
  bits_z = (double*)dalloc(sizeof(double) * B);
  bpp_scratch = new_f32(3 * B);
  Act X8 = new_act(B, H8, W8, 192);            // pixel_unshuffle(x, 8)
  Act F8 = new_act(B, H8, W8, 192);            // pixel_unshuffle(dpb.frame, 8)
  Act FP = new_act(B, H8, W8, CD);             // dpb.feature
  Act FEAT0 = new_act(B, H8, W8, CD);          // temporal feature
  Act PA = new_act(B, H8, W8, CD), PB = new_act(B, H8, W8, CD);   // ping-pong
  Act X1 = new_act(B, H8, W8, CD);
  Act CTXT = new_act(B, H8, W8, CD);
  Act XC = new_act(B, H8, W8, 2 * CD);         // [encoder.conv1 out / decoder.up out | ctx]
  Act XC_lo = slice(XC, 0, CD), CTX = slice(XC, CD, CD);
  Act Y = new_act(B, H16, W16, CY);
  if (padded && variant == DMC_VARIANT_PERFORMANCE)
    fail("variant performance needs height and width in multiples of 64 (the reference does not pad y there, "
         "seg_video_model.py:331)");
  Act YF = new_act(B, Hp16, Wp16, CY);         // FiLM'd y (performance) / hyper input (fast), on the padded grid
  Act YP = padded ? new_act(B, Hp16, Wp16, CY) : Y;     // y replicate-padded for the hyper path
  Act HA = new_act(B, Hp16, Wp16, CZ);
  Act D32 = new_act(B, H32, W32, CZ), H32a = new_act(B, H32, W32, CZ);
  Act D64 = new_act(B, H64, W64, CZ), Z = new_act(B, H64, W64, CZ), ZH = new_act(B, H64, W64, CZ);
  Act U32 = new_act(B, H32, W32, CZ), G32 = new_act(B, H32, W32, CZ);
  Act U16 = new_act(B, Hp16, Wp16, CZ), G16 = new_act(B, Hp16, Wp16, CZ);
  Act HT = new_act(B, H16, W16, 3 * CY);       // [hier | temporal]
  Act HIER = slice(HT, 0, CY), TEMP = slice(HT, CY, 2 * CY);
  Act TD = new_act(B, H16, W16, 2 * CY);
  Act P0 = new_act(B, H16, W16, 3 * CY), P1 = new_act(B, H16, W16, 3 * CY);
  Act PC = new_act(B, H16, W16, 4 * CY);       // [y_hat (running) | q_dec | sigma0 | mu0]
  Act YH0 = slice(PC, 0, CY), PARAMS = slice(PC, CY, 3 * CY);
  Act S0 = new_act(B, H16, W16, 3 * CY), S1 = new_act(B, H16, W16, 3 * CY);
  Act SP = new_act(B, H16, W16, 2 * CY);
  Act YHAT = new_act(B, H16, W16, CY);
  Act FEAT = new_act(B, H8, W8, CD);
  Act R0 = new_act(B, H8, W8, CR), R1 = new_act(B, H8, W8, CR);
  const long long M8 = (long long)B * H8 * W8, M16 = (long long)B * H16 * W16;
  float* sym = new_f32((size_t)M16 * CY);
  float* sig = new_f32((size_t)M16 * CY);
  float* RF = new_f32((size_t)M8 * 192);

  // ---- head: temporal feature (video_model.py:348-351)
  set_phase(PH_FEAT);
  set_prog(&prog_head_i);
  io_dev = (IoSlots*)dalloc(sizeof(IoSlots));
  finite_dev = (int*)dalloc(sizeof(int));
  op([self, F8](cudaStream_t st) { unshuffle8_in(nullptr, F8.v, self->B, 3, self->H, self->W, st, self->IO_SLOT(dpb_frame)); });
  dcb(fa_i, F8, FEAT0, false, nullptr, ns);
  set_prog(&prog_head_p);
  op([self, FP](cudaStream_t st) { nchw_to_s3(nullptr, FP.v, FP.B, 256, FP.H, FP.W, st, self->IO_SLOT(dpb_feature)); });
  {
    EpiSpec s; s.nsplit = ns;
    gemm(FP, fa_p, &FEAT0, s);
  }

  set_prog(&prog_common);
  double* by = bits_y; double* bz = bits_z; int nb = B;
  set_phase(PH_RATE);
  op([by, bz, nb](cudaStream_t st) {
    CUDA_OK(cudaMemsetAsync(by, 0, sizeof(double) * nb, st));
    CUDA_OK(cudaMemsetAsync(bz, 0, sizeof(double) * nb, st));
  });
  set_phase(PH_FEAT);
  tap("feature_in", FEAT0);
  // ---- feature extractor (video_model.py:23-49)
  dcb(fe1[0], FEAT0, PA, false, nullptr, ns);
  dcb(fe1[1], PA, X1, false, nullptr, ns);
  op([self, X1, CTXT, q_feature, M8](cudaStream_t st) {
    scale_cols(X1.v, q_feature + (size_t)self->cur.qp * 256, CTXT.v, M8, st);
  });
  dcb(fe2[0], X1, PA, false, nullptr, ns);
  dcb(fe2[1], PA, PB, false, nullptr, ns);
  dcb(fe2[2], PB, PA, false, nullptr, ns);
  dcb(fe2[3], PA, CTX, false, nullptr, ns);
  tap("ctx", CTX);
  tap("ctx_t", CTXT);
  // ---- encoder (video_model.py:52-75 / seg_video_model.py:41-59)
  set_phase(PH_ENC);
  op([self, X8](cudaStream_t st) { unshuffle8_in(nullptr, X8.v, self->B, 3, self->H, self->W, st, self->IO_SLOT(x)); });
  {
    EpiSpec s; s.nsplit = ns;
    gemm(X8, enc_conv1, &XC_lo, s);
  }
  dcb(enc_b[0], XC, PA, false, nullptr, ns);
  dcb(enc_b[1], PA, PB, false, nullptr, ns);
  dcb(enc_b[2], PB, PA, false, q_encoder, ns);
  {
    EpiSpec s; s.nsplit = ns;
    conv_kxk(PA, enc_down, &Y, s);
  }
  tap("y_enc", Y);

  // ---- mask conditioning
  Act YQ = Y;          // the y that is quantised
  Act HIN = Y;         // the hyper-encoder input
  if (variant == DMC_VARIANT_PERFORMANCE) {
    // SFT (seg_video_model.py:159-196) and FiLM on y itself (:327-328)
    Act MK = new_act(B, H8, W8, 64);
    Act GB = new_act(B, H16, W16, 2 * CY);
    op([self, MK](cudaStream_t st) {
      if (self->cur.mask) unshuffle8_in(nullptr, MK.v, self->B, 1, self->H, self->W, st, self->IO_SLOT(mask));
      else CUDA_OK(cudaMemsetAsync(MK.v.p, 0, (size_t)MK.v.ps * kPlanes * sizeof(h16), st));
    });
    EpiSpec s; s.nsplit = ns;
    gemm(MK, sft_conv1, &PB, s);
    dcb(sft_b[0], PB, PA, false, nullptr, ns);
    dcb(sft_b[1], PA, PB, false, nullptr, ns);
    dcb(sft_b[2], PB, PA, false, q_sft, ns);
    conv_kxk(PA, sft_down, &GB, s);
    op([Y, GB, YF, M16](cudaStream_t st) { film(Y.v, GB.v, YF.v, M16, 128, st); });
    tap("gamma_beta", GB);
    YQ = YF;
    HIN = YF;
  } else if (variant == DMC_VARIANT_FAST || variant == DMC_VARIANT_MASK_PROP) {
    float* mpool = new_f32((size_t)M16);
    float* logits_full = nullptr;
    if (variant == DMC_VARIANT_MASK_PROP) {
      // MaskPredictor (mask_predictor.py:27-46), only when after_i == 0 and a mask was given
      float* mdown = new_f32((size_t)M8);
      float* logit8 = new_f32((size_t)M8);
      logits_full = new_f32((size_t)B * H * W);
      Act ME = new_act(B, H8, W8, CD);
      Act COL = get_scratch("im2col", B, H8, W8, 9 * 3 * CD);
      Act PM1 = new_act(B, H8, W8, CD / 4), PM2 = new_act(B, H8, W8, CD / 4);
      std::vector<Op> pred;
      std::vector<Op>* saved = prog;
      set_prog(&pred);
      op([self, mdown](cudaStream_t st) { bilinear_down8(nullptr, mdown, self->B, self->H, self->W, st, self->IO_SLOT(mask)); });
      op([mdown, me_w, me_b, ME](cudaStream_t st) { conv3x3_c1(mdown, me_w, me_b, ME.v, ME.B, ME.H, ME.W, 256, st); });
      Act srcs[3] = {ME, CTX, CTXT};
      for (int i = 0; i < 3; ++i) {
        Act s = srcs[i];
        int off = i * CD;
        op([s, COL, off](cudaStream_t st) {
          im2col(s.v, COL.v, s.B, s.H, s.W, 3, 1, 1, s.H, s.W, 3 * 256, off, st);
        });
      }
      EpiSpec sa; sa.nsplit = ns; sa.act = ACT_WSILU;
      gemm(COL, mp0, &PM1, sa);
      conv_kxk(PM1, mp2, &PM2, sa);
      op([PM2, mp4_w, mp4_b, logit8, M8](cudaStream_t st) { conv1x1_to1(PM2.v, mp4_w, mp4_b, logit8, M8, 64, st); });
      self->logits_full = logits_full;
      op([self, logit8, H8, W8](cudaStream_t st) {
        // (the slot holds the caller's mask_pred tensor, or the engine's own buffer when the caller passed none)
        bilinear_up8(logit8, nullptr, self->B, H8, W8, st, self->IO_SLOT(mask_pred));
      });
      set_prog(saved);
      auto pred_ops = std::make_shared<std::vector<Op>>(std::move(pred));
      // guarded: the predictor runs only on non-first P frames with a mask
      // (mask_prop_seg_video_model.py:365-368)
      struct Guard { dmc_engine* e; std::shared_ptr<std::vector<Op>> ops; };
      Guard gd{self, pred_ops};
      op([gd](cudaStream_t st) {
        if (gd.e->cur_after_i || !gd.e->cur.mask) return;
        for (auto& f : *gd.ops) f(st);
      });
      ftaps["mask_logits8"] = F32Tap{logit8, B, 1, H8, W8};
    }
    if (padded) op([Y, YP, H16, W16, Hp16, Wp16](cudaStream_t st) { regrid(Y.v, H16, W16, YP.v, Hp16, Wp16, Y.B, st); });
    op([self, mpool, YP, YF, mf_w0, mf_b0, mf_w2, mf_b2, H16, W16, Hp16, Wp16](cudaStream_t st) {
      // the mask the FiLM sees: the caller's, or (mask_prop, not after_i) the predictor's logits -- slot mask_src
      const float* m = self->cur.mask;
      if (m) avgpool16_clamp(nullptr, mpool, self->B, self->H, self->W, st, self->IO_SLOT(mask_src));
      maskfilm_apply(m ? mpool : nullptr, YP.v, YF.v, mf_w0, mf_b0, mf_w2, mf_b2, self->B, Hp16, Wp16, 128, H16, W16, st);
    });
    HIN = YF;
  }
  if (padded && variant == DMC_VARIANT_OLD) {
    op([Y, YP, H16, W16, Hp16, Wp16](cudaStream_t st) { regrid(Y.v, H16, W16, YP.v, Hp16, Wp16, Y.B, st); });
    HIN = YP;
  }
  tap("y", YQ);
  tap("hyper_in", HIN);

  // ---- hyper encoder (video_model.py:123-133)
  dcb(he0, HIN, HA, false, nullptr, ns);
  {
    EpiSpec s; s.nsplit = ns;
    conv_kxk(HA, he1d, &D32, s);
    dcb(he1, D32, H32a, true, nullptr, ns);
    conv_kxk(H32a, he2d, &D64, s);
    dcb(he2, D64, Z, true, nullptr, ns);
  }
  tap("z", Z);
  {
    int HW64 = H64 * W64;
    op([self, Z, ZH, HW64, bitparm, bz](cudaStream_t st) {
      BitparmRow t;
      for (int i = 0; i < 11; ++i) t.p[i] = bitparm[i] + (size_t)self->cur.qp * 128;
      round_z_bits(Z.v, ZH.v, self->B, HW64, 128, t, bz, st);
    });
  }
  tap("z_hat", ZH);
  // ---- hyper decoder (video_model.py:136-146) + temporal prior + fusion (:236-243,149-160)
  set_phase(PH_PRIOR);
  dec_zhat = ZH.v; dec_zC = CZ; dec_zH = H64; dec_zW = W64;
  {
    EpiSpec s; s.nsplit = ns;
    gemm(ZH, hd0u, &U32, s, H64, W64);
    dcb(hd0, U32, G32, true, nullptr, ns);
    gemm(G32, hd1u, &U16, s, H32, W32);
    dcb(hd1, U16, G16, true, nullptr, ns);
    if (padded) {                 // hierarchical params cropped back to y's grid (video_model.py:238)
      Act HIERP = new_act(B, Hp16, Wp16, CY);
      dcb(hd2, G16, HIERP, false, nullptr, ns);
      op([HIERP, HIER, H16, W16, Hp16, Wp16](cudaStream_t st) { regrid(HIERP.v, Hp16, Wp16, HIER.v, H16, W16, HIER.B, st); });
    } else {
      dcb(hd2, G16, HIER, false, nullptr, ns);
    }
    conv_kxk(CTXT, tpd, &TD, s);
    dcb(tpc, TD, TEMP, true, nullptr, ns);
    tap("hier", HIER);
    tap("temporal", TEMP);
    dcb(pf[0], HT, P0, false, nullptr, ns);
    dcb(pf[1], P0, P1, false, nullptr, ns);
    dcb(pf[2], P1, P0, false, nullptr, ns);
    gemm(P0, pf3, &PARAMS, s);
  }
  tap("params", PARAMS);
  // ---- two-step checkerboard quantisation (models/common_model.py:121-149)
  PriorArgs pa;
  memset(&pa, 0, sizeof pa);
  pa.scheme = 2; pa.B = B; pa.H = H16; pa.W = W16; pa.C = CY;
  pa.y = YQ.v; pa.params = PARAMS.v; pa.sp = SP.v; pa.yh = YH0.v; pa.sym = sym; pa.sig = sig;
  dec_prior = pa;
  dec_steps = 2;
  {
    PriorArgs a0 = pa; a0.step = 0;
    set_phase(PH_STEP0);
    op([a0](cudaStream_t st) { prior_step(a0, st); });
    tap("y_hat_0", YH0);
    set_phase(PH_SP1);
    dcb(sp0, PC, S0, false, nullptr, ns);
    dcb(sp1, S0, S1, false, nullptr, ns);
    EpiSpec s; s.nsplit = ns;
    gemm(S1, sp2, &SP, s);
    tap("spatial_prior", SP);
    PriorArgs a1 = pa; a1.step = 1;
    set_phase(PH_STEP1);
    op([a1](cudaStream_t st) { prior_step(a1, st); });
    int formula = refactor ? 1 : 0;
    set_phase(PH_FIN);
    op([a1, YHAT, formula, by](cudaStream_t st) { prior_finish(a1, YHAT.v, formula, by, st); });
  }
  tap("y_hat", YHAT);
  set_phase(PH_DEC);
  ftaps["y_q"] = F32Tap{sym, B, CY, H16, W16};
  ftaps["scales_hat"] = F32Tap{sig, B, CY, H16, W16};
  // ---- decoder (video_model.py:78-97 / seg_video_model.py:62-77)
  {
    EpiSpec s; s.nsplit = ns;
    if (refactor) { s.scale_table = q_decoder; s.scale_C = CD; }
    conv_kxk(YHAT, dec_up, &XC_lo, s, true);
    dcb(dec_b[0], XC, PA, false, nullptr, ns);
    dcb(dec_b[1], PA, PB, false, nullptr, ns);
    dcb(dec_b[2], PB, PA, false, nullptr, ns);
    EpiSpec s2; s2.nsplit = ns;
    if (!refactor) { s2.scale_table = q_decoder; s2.scale_C = CD; }
    gemm(PA, dec_proj, &FEAT, s2);
  }
  op([self, FEAT](cudaStream_t st) { s3_to_nchw(FEAT.v, nullptr, FEAT.B, 256, FEAT.H, FEAT.W, st, self->IO_SLOT(feature)); });
  // ---- reconstruction (video_model.py:100-120)
  dcb(rec[0], FEAT, R0, false, nullptr, ns_recon);
  dcb(rec[1], R0, R1, false, nullptr, ns_recon);
  dcb(rec[2], R1, R0, false, nullptr, ns_recon);
  dcb(rec[3], R0, R1, false, q_recon, ns_recon);
  {
    EpiSpec s; s.nsplit = ns_recon; s.out_f32 = RF; s.ld_f32 = 192;
    gemm(R1, rec_head, nullptr, s);
  }
  op([self, RF](cudaStream_t st) { shuffle8_out(RF, 192, nullptr, self->B, 3, self->H, self->W, st, self->IO_SLOT(x_hat)); });
  // ---- rate (video_model.py:373-378)
  int pixels = H * W;
  set_phase(PH_RATE);
  op([self, by, bz, pixels](cudaStream_t st) { finalize_bpp(by, bz, nullptr, self->B, pixels, st, self->IO_SLOT(bpp3)); });
  set_phase(PH_DEC);
  {
    // the reference's _finite_check sites (seg_video_model_fast.py:353-371,269-276), checked in ONE launch at the end
    // of the frame instead of twelve host syncs inside it: bit i of the flag names tensor i (include/dmc_b200.h)
    FiniteList fl;
    memset(&fl, 0, sizeof fl);
    const Act* list[8] = {&FEAT0, &CTX, &CTXT, &YQ, &Z, &PARAMS, &YHAT, &FEAT};
    for (int i = 0; i < 8; ++i) { fl.v[i] = list[i]->v; fl.M[i] = list[i]->M(); }
    fl.n = 8;
    op([self, fl](cudaStream_t st) {
      CUDA_OK(cudaMemsetAsync(self->finite_dev, 0, sizeof(int32_t), st));
      finite_check(fl, self->finite_dev, st);
      copy_flag(self->finite_dev, self->IO_SLOT(finite), st);      // (a null caller flag is skipped on the device)
    });
  }
  flush_chain();
}

// ====================================================================== I-frame program
void dmc_engine::build_intra() {
  // the hyper path works on y replicate-padded to multiples of 4 (models/common_model.py:68-72): Hp16 x Wp16
  const int H8 = H / 8, W8 = W / 8, H16 = H / 16, W16 = W / 16, Hp16 = (H16 + 3) / 4 * 4, Wp16 = (W16 + 3) / 4 * 4,
            H32 = Hp16 / 2, W32 = Wp16 / 2, H64 = Hp16 / 4, W64 = Wp16 / 4;
  const bool padded = Hp16 != H16 || Wp16 != W16;
  const int CE = 368, N = 256, CZ = 128, QP = 64;
  const int ns = 3;
  dmc_engine* self = this;

  float* q_enc = add_table("q_scale_enc", {QP, CE, 1, 1});
  float* q_dec = add_table("q_scale_dec", {QP, CE, 1, 1});
  float* bitparm[11];
  {
    const char* names[11] = {"f1.h", "f1.b", "f1.a", "f2.h", "f2.b", "f2.a",
                             "f3.h", "f3.b", "f3.a", "f4.h", "f4.b"};
    for (int i = 0; i < 11; ++i)
      bitparm[i] = add_table(std::string("bit_estimator_z.") + names[i], {QP, CZ, 1, 1});
  }
  DCB* enc1 = add_dcb("enc.enc_1", 192, CE);
  DCB* enc2[6];
  for (int i = 0; i < 6; ++i) enc2[i] = add_dcb("enc.enc_2." + std::to_string(i), CE, CE);
  Conv* enc_down = add_conv("enc.enc_2.6", CE, N, 3, 2, 1);
  DCB* he0 = add_dcb("hyper_enc.0", N, CZ);
  Conv* he1d = add_conv("hyper_enc.1.down", CZ, CZ, 2, 2, 0);
  DCB* he1 = add_dcb("hyper_enc.1.conv", CZ, CZ);
  Conv* he2d = add_conv("hyper_enc.2.down", CZ, CZ, 2, 2, 0);
  DCB* he2 = add_dcb("hyper_enc.2.conv", CZ, CZ);
  Conv* hd0u = add_conv("hyper_dec.0.up.conv.0", CZ, CZ * 4, 1, 1, 0, PACK_SHUF2);
  DCB* hd0 = add_dcb("hyper_dec.0.conv", CZ, CZ);
  Conv* hd1u = add_conv("hyper_dec.1.up.conv.0", CZ, CZ * 4, 1, 1, 0, PACK_SHUF2);
  DCB* hd1 = add_dcb("hyper_dec.1.conv", CZ, CZ);
  DCB* hd2 = add_dcb("hyper_dec.2", CZ, N);
  DCB* pf0 = add_dcb("y_prior_fusion.0", N, 2 * N);
  DCB* pf1 = add_dcb("y_prior_fusion.1", 2 * N, 2 * N);
  DCB* pf2 = add_dcb("y_prior_fusion.2", 2 * N, 2 * N);
  Conv* pf3 = add_conv("y_prior_fusion.3", 2 * N, 2 * N + 2, 1, 1, 0);
  Conv* red = add_conv("y_spatial_prior_reduction", 2 * N + 2, N, 1, 1, 0);
  DCB* ad[3];
  for (int i = 0; i < 3; ++i)
    ad[i] = add_dcb("y_spatial_prior_adaptor_" + std::to_string(i + 1), 2 * N, 2 * N, true);
  DCB* spb[3];
  for (int i = 0; i < 3; ++i) spb[i] = add_dcb("y_spatial_prior." + std::to_string(i), 2 * N, 2 * N);
  Conv* sp3 = add_conv("y_spatial_prior.3", 2 * N, 2 * N, 1, 1, 0);
  Conv* dec_up = add_conv("dec.dec_1.0.up.conv.0", N, CE * 4, 1, 1, 0, PACK_SHUF2);
  DCB* dec0 = add_dcb("dec.dec_1.0.conv", CE, CE);
  DCB* dec1[12];
  for (int i = 0; i < 12; ++i) dec1[i] = add_dcb("dec.dec_1." + std::to_string(i + 1), CE, CE);
  DCB* dec2 = add_dcb("dec.dec_2", CE, 192);

  bits_y = (double*)dalloc(sizeof(double) * B);
  bits_z = (double*)dalloc(sizeof(double) * B);
  Act X8 = new_act(B, H8, W8, 192);
  Act EA = new_act(B, H8, W8, CE), EB = new_act(B, H8, W8, CE);
  Act Y = new_act(B, H16, W16, N);
  Act YP = padded ? new_act(B, Hp16, Wp16, N) : Y;      // y replicate-padded for the hyper path
  Act HA = new_act(B, Hp16, Wp16, CZ);
  Act D32 = new_act(B, H32, W32, CZ), H32a = new_act(B, H32, W32, CZ);
  Act D64 = new_act(B, H64, W64, CZ), Z = new_act(B, H64, W64, CZ), ZH = new_act(B, H64, W64, CZ);
  Act U32 = new_act(B, H32, W32, CZ), G32 = new_act(B, H32, W32, CZ);
  Act U16 = new_act(B, Hp16, Wp16, CZ), G16 = new_act(B, Hp16, Wp16, CZ);
  Act HP = new_act(B, H16, W16, N);
  Act FA = new_act(B, H16, W16, 2 * N), FB = new_act(B, H16, W16, 2 * N);
  Act PARAMS = new_act(B, H16, W16, 2 * N + 2);
  Act Q = new_act(B, H16, W16, 2 * N);          // [y_hat so far | reduced common params]
  Act YHS = slice(Q, 0, N), COMMON = slice(Q, N, N);
  Act SP = new_act(B, H16, W16, 2 * N);
  Act YHAT = new_act(B, H16, W16, N);
  const long long M8 = (long long)B * H8 * W8, M16 = (long long)B * H16 * W16;
  float* sym = new_f32((size_t)M16 * N);
  float* sig = new_f32((size_t)M16 * N);
  float* RF = new_f32((size_t)M8 * 192);

  set_prog(&prog_common);
  double* by = bits_y; double* bz = bits_z; int nb = B;
  set_phase(PH_RATE);
  op([by, bz, nb](cudaStream_t st) {
    CUDA_OK(cudaMemsetAsync(by, 0, sizeof(double) * nb, st));
    CUDA_OK(cudaMemsetAsync(bz, 0, sizeof(double) * nb, st));
  });
  // encoder (image_model.py:16-43)
  set_phase(PH_ENC);
  io_dev = (IoSlots*)dalloc(sizeof(IoSlots));
  op([self, X8](cudaStream_t st) { unshuffle8_in(nullptr, X8.v, self->B, 3, self->H, self->W, st, self->IO_SLOT(x)); });
  dcb(enc1, X8, EA, false, q_enc, ns);
  {
    Act a = EA, b = EB;
    for (int i = 0; i < 6; ++i) { dcb(enc2[i], a, b, false, nullptr, ns); std::swap(a, b); }
    EpiSpec s; s.nsplit = ns;
    conv_kxk(a, enc_down, &Y, s);
  }
  tap("y", Y);
  // hyper path (image_model.py:216-226)
  {
    EpiSpec s; s.nsplit = ns;
    if (padded) op([Y, YP, H16, W16, Hp16, Wp16](cudaStream_t st) { regrid(Y.v, H16, W16, YP.v, Hp16, Wp16, Y.B, st); });
    dcb(he0, YP, HA, false, nullptr, ns);
    conv_kxk(HA, he1d, &D32, s);
    dcb(he1, D32, H32a, true, nullptr, ns);
    conv_kxk(H32a, he2d, &D64, s);
    dcb(he2, D64, Z, true, nullptr, ns);
    tap("z", Z);
    int HW64 = H64 * W64;
    op([self, Z, ZH, HW64, bitparm, bz](cudaStream_t st) {
      BitparmRow t;
      for (int i = 0; i < 11; ++i) t.p[i] = bitparm[i] + (size_t)self->cur.qp * 128;
      round_z_bits(Z.v, ZH.v, self->B, HW64, 128, t, bz, st);
    });
    tap("z_hat", ZH);
    set_phase(PH_PRIOR);
    dec_zhat = ZH.v; dec_zC = CZ; dec_zH = H64; dec_zW = W64;
    gemm(ZH, hd0u, &U32, s, H64, W64);
    dcb(hd0, U32, G32, true, nullptr, ns);
    gemm(G32, hd1u, &U16, s, H32, W32);
    dcb(hd1, U16, G16, true, nullptr, ns);
    if (padded) {                 // the prior fusion runs on the padded grid, params are cropped (image_model.py:224-226)
      Act HPP = new_act(B, Hp16, Wp16, N);
      Act FAP = new_act(B, Hp16, Wp16, 2 * N), FBP = new_act(B, Hp16, Wp16, 2 * N);
      Act PARAMSP = new_act(B, Hp16, Wp16, 2 * N + 2);
      dcb(hd2, G16, HPP, false, nullptr, ns);
      dcb(pf0, HPP, FAP, false, nullptr, ns);
      dcb(pf1, FAP, FBP, false, nullptr, ns);
      dcb(pf2, FBP, FAP, false, nullptr, ns);
      gemm(FAP, pf3, &PARAMSP, s);
      op([PARAMSP, PARAMS, H16, W16, Hp16, Wp16](cudaStream_t st) {
        regrid(PARAMSP.v, Hp16, Wp16, PARAMS.v, H16, W16, PARAMS.B, st);
      });
    } else {
      dcb(hd2, G16, HP, false, nullptr, ns);
      dcb(pf0, HP, FA, false, nullptr, ns);
      dcb(pf1, FA, FB, false, nullptr, ns);
      dcb(pf2, FB, FA, false, nullptr, ns);
      gemm(FA, pf3, &PARAMS, s);
    }
    tap("params", PARAMS);
    gemm(PARAMS, red, &COMMON, s);
  }
  // four-step prior (models/common_model.py:188-248)
  PriorArgs pa;
  memset(&pa, 0, sizeof pa);
  pa.scheme = 4; pa.B = B; pa.H = H16; pa.W = W16; pa.C = N;
  pa.y = Y.v; pa.params = PARAMS.v; pa.sp = SP.v; pa.yh = YHS.v; pa.sym = sym; pa.sig = sig;
  dec_prior = pa;
  dec_steps = 4;
  for (int step = 0; step < 4; ++step) {
    if (step > 0) {
      set_phase(PH_STEP0 + 2 * step - 1);         // PH_SP1 / PH_SP2 / PH_SP3
      dcb(ad[step - 1], Q, FA, false, nullptr, ns);
      dcb(spb[0], FA, FB, false, nullptr, ns);
      dcb(spb[1], FB, FA, false, nullptr, ns);
      dcb(spb[2], FA, FB, false, nullptr, ns);
      EpiSpec s; s.nsplit = ns;
      gemm(FB, sp3, &SP, s);
    }
    PriorArgs a = pa; a.step = step;
    set_phase(PH_STEP0 + 2 * step);               // PH_STEP0 ... PH_STEP3
    op([a](cudaStream_t st) { prior_step(a, st); });
  }
  set_phase(PH_FIN);
  {
    PriorArgs a = pa; a.step = 3;
    op([a, YHAT, by](cudaStream_t st) { prior_finish(a, YHAT.v, 0, by, st); });
  }
  tap("y_hat", YHAT);
  set_phase(PH_DEC);
  ftaps["y_q"] = F32Tap{sym, B, N, H16, W16};
  ftaps["scales_hat"] = F32Tap{sig, B, N, H16, W16};
  // decoder (image_model.py:46-93)
  {
    EpiSpec s; s.nsplit = ns;
    gemm(YHAT, dec_up, &EA, s, H16, W16);
    dcb(dec0, EA, EB, true, nullptr, ns);
    Act a = EB, b = EA;
    for (int i = 0; i < 12; ++i) {
      dcb(dec1[i], a, b, false, i == 11 ? q_dec : nullptr, ns);
      std::swap(a, b);
    }
    dcb(dec2, a, a, false, nullptr, ns, RF, 192);
  }
  op([self, RF](cudaStream_t st) { shuffle8_out(RF, 192, nullptr, self->B, 3, self->H, self->W, st, self->IO_SLOT(x_hat)); });
  int pixels = H * W;
  set_phase(PH_RATE);
  op([self, by, bz, pixels](cudaStream_t st) { finalize_bpp(by, bz, nullptr, self->B, pixels, st, self->IO_SLOT(bpp3)); });
  flush_chain();
}

void dmc_engine::finalize(cudaStream_t st) {
  for (auto& s : slots)
    if (!s.set) fail("weight '%s' was never set", s.key.c_str());
  (void)st;
  finalized = true;
}

// ====================================================================== C ABI
namespace {
template <class F>
int guarded(dmc_engine* e, F f) {
  try {
    f();
    return DMC_OK;
  } catch (const std::exception& ex) {
    if (e) e->error = ex.what();
    else g_create_error = ex.what();
    const char* w = ex.what();
    if (strstr(w, "cuda") || strstr(w, "CUDA")) return DMC_E_CUDA;
    return DMC_E_INVALID;
  }
}
}  // namespace

extern "C" {

int dmc_create(int variant, int batch, int height, int width, int flags, dmc_engine** out) {
  if (!out) return DMC_E_INVALID;
  *out = nullptr;
  dmc_engine* e = nullptr;
  int rc = guarded(nullptr, [&] {
    if (variant < DMC_VARIANT_OLD || variant > DMC_VARIANT_INTRA) fail("unknown variant %d", variant);
    if (batch < 1 || height < 16 || width < 16 || height % 16 || width % 16)
      fail("batch must be >= 1 and height/width multiples of 16 (got %d, %d, %d)", batch, height, width);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      fail("CUDA device required: this engine has no CPU path");
    e = new dmc_engine();
    e->variant = variant; e->B = batch; e->H = height; e->W = width; e->flags = flags;
    CUDA_OK(cudaGetDevice(&e->device));
    try {
      if (variant == DMC_VARIANT_INTRA) e->build_intra();
      else e->build_p();
    } catch (...) {
      delete e;
      e = nullptr;
      throw;
    }
  });
  if (rc == DMC_OK) *out = e;
  return rc;
}

void dmc_destroy(dmc_engine* e) { delete e; }

const char* dmc_last_error(const dmc_engine* e) { return e ? e->error.c_str() : g_create_error.c_str(); }

int dmc_num_weights(const dmc_engine* e) { return e ? (int)e->slots.size() : DMC_E_INVALID; }
const char* dmc_weight_key(const dmc_engine* e, int i) {
  if (!e || i < 0 || i >= (int)e->slots.size()) return nullptr;
  return e->slots[i].key.c_str();
}
int dmc_weight_shape(const dmc_engine* e, int i, int64_t* shape4) {
  if (!e || i < 0 || i >= (int)e->slots.size() || !shape4) return DMC_E_INVALID;
  const auto& s = e->slots[i].shape;
  for (size_t k = 0; k < s.size() && k < 4; ++k) shape4[k] = s[k];
  return (int)s.size();
}

int dmc_set_weight(dmc_engine* e, const char* key, const float* dev_ptr, const int64_t* shape, int ndim,
                   void* stream) {
  if (!e || !key || !dev_ptr) return DMC_E_INVALID;
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    auto it = e->slot_index.find(key);
    if (it == e->slot_index.end()) fail("unknown state_dict key '%s'", key);
    WSlot& s = e->slots[it->second];
    if ((int)s.shape.size() != ndim) fail("%s: expected %d dims, got %d", key, (int)s.shape.size(), ndim);
    for (int i = 0; i < ndim; ++i)
      if (s.shape[i] != shape[i])
        fail("%s: dim %d is %lld, expected %lld", key, i, (long long)shape[i], (long long)s.shape[i]);
    s.load(dev_ptr, (cudaStream_t)stream);
    CUDA_OK(cudaGetLastError());
    s.set = true;
  });
}

int dmc_finalize_weights(dmc_engine* e, void* stream) {
  if (!e) return DMC_E_INVALID;
  int rc = guarded(e, [&] { e->finalize((cudaStream_t)stream); });
  return rc == DMC_E_INVALID ? DMC_E_STATE : rc;
}

int dmc_forward(dmc_engine* e, const float* x, const float* mask, const float* dpb_frame,
                const float* dpb_feature, int qp, int after_i, float* x_hat, float* feature,
                float* bpp3, float* mask_pred, int32_t* finite_flag, void* stream) {
  if (!e) return DMC_E_INVALID;
  if (e->variant == DMC_VARIANT_INTRA) { e->error = "dmc_forward on an intra engine"; return DMC_E_INVALID; }
  if (!e->finalized) { e->error = "weights not finalised"; return DMC_E_STATE; }
  if (!x || !x_hat || !feature || !bpp3 || (after_i ? !dpb_frame : !dpb_feature) || qp < 0 || qp >= 72) {
    e->error = "dmc_forward: null tensor or qp outside [0,72)";
    return DMC_E_INVALID;
  }
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    auto& c = e->cur;
    c.x = x; c.mask = (e->variant == DMC_VARIANT_OLD) ? nullptr : mask;
    c.dpb_frame = dpb_frame; c.dpb_feature = dpb_feature; c.qp = qp;
    c.x_hat = x_hat; c.feature = feature; c.bpp3 = bpp3; c.mask_pred = mask_pred; c.finite = finite_flag;
    e->cur_after_i = after_i != 0;
    cudaStream_t st = (cudaStream_t)stream;
    IoSlots io;
    memset(&io, 0, sizeof io);
    io.x = x; io.mask = c.mask; io.dpb_frame = dpb_frame; io.dpb_feature = dpb_feature;
    io.x_hat = x_hat; io.feature = feature; io.bpp3 = bpp3; io.finite = finite_flag;
    io.mask_pred = mask_pred ? mask_pred : e->logits_full;
    io.mask_src = c.mask;
    if (e->variant == DMC_VARIANT_MASK_PROP && !after_i && c.mask) io.mask_src = io.mask_pred;
    set_io(e->io_dev, io, st);
    // what the program's structure depends on: the head, the per-QP table rows, whether a mask came with the frame
    const uint64_t key = (uint64_t)(after_i != 0) | ((uint64_t)(c.mask != nullptr) << 1) | ((uint64_t)qp << 8);
    e->run_graph(key, st, [&](cudaStream_t s) {
      e->run(after_i ? e->prog_head_i : e->prog_head_p, s);
      e->run(e->prog_common, s);
    });
    CUDA_OK(cudaGetLastError());
  });
}

int dmci_forward(dmc_engine* e, const float* x, int qp, float* x_hat, float* bpp3, void* stream) {
  if (!e) return DMC_E_INVALID;
  if (e->variant != DMC_VARIANT_INTRA) { e->error = "dmci_forward on a P-frame engine"; return DMC_E_INVALID; }
  if (!e->finalized) { e->error = "weights not finalised"; return DMC_E_STATE; }
  if (!x || !x_hat || !bpp3 || qp < 0 || qp >= 64) {
    e->error = "dmci_forward: null tensor or qp outside [0,64)";
    return DMC_E_INVALID;
  }
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    auto& c = e->cur;
    c = dmc_engine::Cur();
    c.x = x; c.qp = qp; c.x_hat = x_hat; c.bpp3 = bpp3;
    cudaStream_t st = (cudaStream_t)stream;
    IoSlots io;
    memset(&io, 0, sizeof io);
    io.x = x; io.x_hat = x_hat; io.bpp3 = bpp3;
    set_io(e->io_dev, io, st);
    e->run_graph((uint64_t)qp << 8, st, [&](cudaStream_t s) { e->run(e->prog_common, s); });
    CUDA_OK(cudaGetLastError());
  });
}

// ---- decoder side: the phases of the frame program a decoder has, with the caller's entropy decoder between them
int dmc_decode_begin(dmc_engine* e, const float* dpb_frame, const float* dpb_feature, int qp, int after_i,
                     const float* z_hat, void* stream) {
  if (!e || !z_hat) return DMC_E_INVALID;
  if (!e->finalized) { e->error = "weights not finalised"; return DMC_E_STATE; }
  const bool intra = e->variant == DMC_VARIANT_INTRA;
  if (qp < 0 || qp >= (intra ? 64 : 72) || (!intra && (after_i ? !dpb_frame : !dpb_feature))) {
    e->error = "dmc_decode_begin: null dpb tensor or qp out of range";
    return DMC_E_INVALID;
  }
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    e->cur = dmc_engine::Cur();
    e->cur.qp = qp;
    e->cur_after_i = after_i != 0;
    IoSlots io;
    memset(&io, 0, sizeof io);
    io.dpb_frame = dpb_frame; io.dpb_feature = dpb_feature;
    set_io(e->io_dev, io, st);
    if (!intra) {
      e->run(after_i ? e->prog_head_i : e->prog_head_p, st, 1u << dmc_engine::PH_FEAT);
      e->run(e->prog_common, st, 1u << dmc_engine::PH_FEAT);
    }
    nchw_to_s3(z_hat, e->dec_zhat, e->B, e->dec_zC, e->dec_zH, e->dec_zW, st);
    e->run(e->prog_common, st, 1u << dmc_engine::PH_PRIOR);
    CUDA_OK(cudaGetLastError());
  });
}

int dmc_decode_sigma(dmc_engine* e, int step, float* sigma_out, void* stream) {
  if (!e || !sigma_out || step < 0 || step >= e->dec_steps) return DMC_E_INVALID;
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (step > 0) e->run(e->prog_common, st, 1u << (dmc_engine::PH_STEP0 + 2 * step - 1));   // the spatial prior of this step
    PriorArgs a = e->dec_prior;
    a.step = step;
    a.mode = 2;
    prior_step(a, st);
    f32rows_to_nchw(a.sig, a.C, sigma_out, a.B, a.C, a.H, a.W, st);
    CUDA_OK(cudaGetLastError());
  });
}

int dmc_decode_symbols(dmc_engine* e, int step, const float* symbols, void* stream) {
  if (!e || !symbols || step < 0 || step >= e->dec_steps) return DMC_E_INVALID;
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    PriorArgs a = e->dec_prior;
    a.step = step;
    a.mode = 1;
    a.sym_in = symbols;
    prior_step(a, (cudaStream_t)stream);
    CUDA_OK(cudaGetLastError());
  });
}

int dmc_decode_finish(dmc_engine* e, float* x_hat, float* feature, void* stream) {
  if (!e || !x_hat || (e->variant != DMC_VARIANT_INTRA && !feature)) return DMC_E_INVALID;
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    IoSlots io;
    memset(&io, 0, sizeof io);
    io.x_hat = x_hat; io.feature = feature;
    set_io(e->io_dev, io, st);
    e->run(e->prog_common, st, (1u << dmc_engine::PH_FIN) | (1u << dmc_engine::PH_DEC));
    CUDA_OK(cudaGetLastError());
  });
}

int dmc_get_tap(dmc_engine* e, const char* name, float* dst, int64_t capacity, int64_t* shape4,
                void* stream) {
  if (!e || !name) return DMC_E_INVALID;
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    auto it = e->taps.find(name);
    if (it != e->taps.end()) {
      const Act& a = it->second;
      if (shape4) { shape4[0] = a.B; shape4[1] = a.v.C; shape4[2] = a.H; shape4[3] = a.W; }
      if (!dst) return;
      if (capacity < a.M() * a.v.C) fail("tap '%s' needs %lld elements", name, a.M() * a.v.C);
      s3_to_nchw(a.v, dst, a.B, a.v.C, a.H, a.W, st);
      return;
    }
    auto jt = e->ftaps.find(name);
    if (jt == e->ftaps.end()) fail("no tap named '%s' (create with DMC_FLAG_KEEP_TAPS)", name);
    const auto& f = jt->second;
    if (shape4) { shape4[0] = f.B; shape4[1] = f.C; shape4[2] = f.H; shape4[3] = f.W; }
    if (!dst) return;
    if (capacity < (int64_t)f.B * f.C * f.H * f.W) fail("tap '%s' too small a buffer", name);
    f32rows_to_nchw(f.p, f.C, dst, f.B, f.C, f.H, f.W, st);
  });
}

int dmc_frame_stats(double* stats7, const float* x_hat, const float* x, const float* mask,
                    const float* bpp3, int batch, int height, int width, void* stream) {
  if (!stats7 || !x_hat || !x) return DMC_E_INVALID;
  frame_stats(stats7, x_hat, x, mask, bpp3, batch, height, width, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? DMC_OK : DMC_E_CUDA;
}

int dmc_frames_from_u8(const uint8_t* img, const uint8_t* mask, float* out, int frames, int height, int width, int top,
                       int left, int crop_h, int crop_w, int out_channels, int bgr, int mask_threshold, void* stream) {
  if (!img || !out || frames < 1 || height < 1 || width < 1 || crop_h < 1 || crop_w < 1 || top < 0 || left < 0 ||
      top + crop_h > height || left + crop_w > width || (out_channels != 3 && out_channels != 4) || frames > 65535 ||
      crop_h > 65535)
    return DMC_E_INVALID;
  frames_from_u8(img, mask, out, frames, height, width, top, left, crop_h, crop_w, out_channels, bgr != 0,
                 mask_threshold, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? DMC_OK : DMC_E_CUDA;
}

int dmc_mask_from_logits(const float* logits, float* mask, int64_t n, void* stream) {
  if (!logits || !mask || n < 0 || ((uintptr_t)logits | (uintptr_t)mask) % 16) return DMC_E_INVALID;
  if (n) mask_from_logits(logits, mask, (long long)n, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? DMC_OK : DMC_E_CUDA;
}

int64_t dmc_kernel_launches(void) { return (int64_t)launch_count(); }
int dmc_profile_enable(dmc_engine* e, int on) {
  if (!e) return DMC_E_INVALID;
  e->profile = on != 0;
  gemm_s3_set_plain_launch(on != 0);
  return DMC_OK;
}
int dmc_profile_read(dmc_engine* e, double* gemm_ms, int64_t* gemm_launches, double* gemm_flops,
                     double* issued_flops) {
  if (!e) return DMC_E_INVALID;
  return guarded(e, [&] {
    DeviceGuard dg(e->device);
    CUDA_OK(cudaDeviceSynchronize());
    double ms = 0;
    for (auto& pr : e->prof_events) {
      float t = 0;
      CUDA_OK(cudaEventElapsedTime(&t, pr.first, pr.second));
      ms += t;
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    if (gemm_ms) *gemm_ms = ms;
    if (gemm_launches) *gemm_launches = (int64_t)e->prof_events.size();
    if (gemm_flops) *gemm_flops = e->prof_flops;
    if (issued_flops) *issued_flops = e->prof_issued;
    e->prof_events.clear();
    e->prof_flops = e->prof_issued = 0;
  });
}
int dmc_num_sms(void) { return num_sms(); }
const char* dmc_version(void) { return "dmc_b200 0.2 (sm_100a)"; }
int dmc_set_acc_comp(float kappa) {
  acc_comp_set_kappa(kappa);
  return DMC_OK;
}
float dmc_get_acc_comp(void) { return acc_comp_kappa(); }

}  // extern "C"

// ---------------------------------------------------------------- single-operator entry points
namespace {
void set_slot(dmc_engine& e, const std::string& key, const float* p, cudaStream_t st) {
  auto it = e.slot_index.find(key);
  if (it == e.slot_index.end()) fail("internal: no slot %s", key.c_str());
  e.slots[it->second].load(p, st);
  e.slots[it->second].set = true;
}
}  // namespace

extern "C" int dmc_op_conv2d(const float* x, const float* weight, const float* bias, float* out, int batch,
                             int cin, int height, int width, int cout, int ksize, int stride, int padding,
                             int groups, int act, int nsplit, int backend, void* stream) {
  return guarded(nullptr, [&] {
    if (!x || !weight || !out) fail("dmc_op_conv2d: null tensor");
    cudaStream_t st = (cudaStream_t)stream;
    dmc_engine e;
    e.variant = -1; e.B = batch; e.H = height; e.W = width;
    e.flags = backend == 1 ? DMC_FLAG_SIMT_GEMM : 0;
    e.prog = &e.prog_common;
    Act in = e.new_act(batch, height, width, cin);
    int Ho = (height + 2 * padding - ksize) / stride + 1, Wo = (width + 2 * padding - ksize) / stride + 1;
    Act o = e.new_act(batch, Ho, Wo, cout);
    nchw_to_s3(x, in.v, batch, cin, height, width, st);
    if (groups == cin && groups > 1) {
      if (ksize != 3 || stride != 1 || padding != 1 || cin != cout) fail("depthwise: only 3x3 s1 p1");
      DW* d = e.add_dw("w", cin);
      set_slot(e, "w.weight", weight, st);
      if (bias) set_slot(e, "w.bias", bias, st);
      else CUDA_OK(cudaMemsetAsync(d->bias, 0, sizeof(float) * cin, st));
      dwconv3x3(in.v, d->w9c, d->bias, o.v, batch, height, width, st);
    } else if (groups == 1) {
      Conv* c = e.add_conv("w", cin, cout, ksize, stride, padding);
      set_slot(e, "w.weight", weight, st);
      if (bias) set_slot(e, "w.bias", bias, st);
      else pack_gemm_bias(nullptr, cout, c->g, st);
      EpiSpec s; s.act = act; s.nsplit = nsplit;
      if (ksize == 1 && stride == 1 && padding == 0) e.gemm(in, c, &o, s);
      else e.conv_kxk(in, c, &o, s);
      e.flush_chain();
      e.run(e.prog_common, st);
    } else {
      fail("groups must be 1 or cin");
    }
    s3_to_nchw(o.v, out, batch, cout, Ho, Wo, st);
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
  });
}

extern "C" int dmc_op_depth_conv_block(const float* x, const float* const* w12, const float* quant_step,
                                       float* out, int batch, int cin, int cout, int height, int width,
                                       int shortcut, int nsplit, int backend, void* stream) {
  return guarded(nullptr, [&] {
    if (!x || !w12 || !out) fail("dmc_op_depth_conv_block: null tensor");
    cudaStream_t st = (cudaStream_t)stream;
    dmc_engine e;
    e.variant = -1; e.B = batch; e.H = height; e.W = width;
    e.flags = backend == 1 ? DMC_FLAG_SIMT_GEMM : 0;
    e.prog = &e.prog_common;
    DCB* b = e.add_dcb("b", cin, cout, w12[0] != nullptr);
    if ((cin != cout) && !w12[0]) fail("adaptor weights required when cin != cout");
    const char* names[6] = {"b.adaptor", "b.dc.0", "b.dc.2", "b.dc.3", "b.ffn.0", "b.ffn.2"};
    for (int i = 0; i < 6; ++i) {
      if (i == 0 && !b->adaptor) continue;
      if (!w12[2 * i] || !w12[2 * i + 1]) fail("missing weight %s", names[i]);
      set_slot(e, std::string(names[i]) + ".weight", w12[2 * i], st);
      set_slot(e, std::string(names[i]) + ".bias", w12[2 * i + 1], st);
    }
    float* table = nullptr;
    if (quant_step) {
      table = e.new_f32(cout);
      CUDA_OK(cudaMemcpyAsync(table, quant_step, sizeof(float) * cout, cudaMemcpyDeviceToDevice, st));
    }
    Act in = e.new_act(batch, height, width, cin);
    Act o = e.new_act(batch, height, width, cout);
    nchw_to_s3(x, in.v, batch, cin, height, width, st);
    e.cur.qp = 0;
    e.dcb(b, in, o, shortcut != 0, table, nsplit);
    e.flush_chain();
    e.run(e.prog_common, st);
    s3_to_nchw(o.v, out, batch, cout, height, width, st);
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
  });
}

extern "C" int dmc_op_gaussian_bits(const float* sym, const float* sigma, float* bits, int64_t n, int formula,
                                    void* stream) {
  if (!sym || !sigma || !bits || n < 0) return DMC_E_INVALID;
  gaussian_bits(sym, sigma, bits, n, formula, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? DMC_OK : DMC_E_CUDA;
}

// ---------------------------------------------------------------- training mode: DepthConvBlock forward / backward
// SURVEY 8f rank 2 ("autograd through fused DCB").  The forward program is the block of the frame engine (dcb():
// chain launches + the depthwise kernel); the backward program recomputes the block's intermediates from x -- only x
// is kept between the two passes -- and walks the block in reverse:
//   data gradients  = contractions with the transposed weights on the same tcgen05 chain kernel (3-term product),
//                     residual gradients folded into their epilogues;
//   weight gradients = k_wgrad_s3 over the pixel axis (train.cu), bias gradients = column sums;
//   WSiLU' / chunk-add' / depthwise gradients = the elementwise kernels of train.cu.
struct dmc_dcb_train {
  dmc_engine e;
  // backward = head (reads the caller's tensors) + body (works on the handle's own buffers only: replayed as ONE CUDA
  // graph per set of requested gradients) + tail (writes the caller's tensors)
  std::vector<dmc_engine::Op> prog_bwd_head, prog_bwd, prog_bwd_tail;
  float* gflat = nullptr;        // the twelve parameter gradients, in order, as the body writes them
  size_t goff[13] = {};
  int gmask = 0;                 // bit i: gradient i is wanted (this call)
  long long bwd_calls = 0;
  int cin = 0, cout = 0, shortcut = 0, has_qs = 0, terms = 3;
  DCB* blk = nullptr;
  Conv *T_ad = nullptr, *T_dc0 = nullptr, *T_dc3 = nullptr, *T_ffn0 = nullptr, *T_ffn2 = nullptr, *P_ffn0 = nullptr;
  bool fwd_packed = false, bwd_packed = false;
  float *w9c_flip = nullptr, *zero_bias = nullptr, *qs_table = nullptr, *part = nullptr;
  // Weight / bias gradients hang off the chain of data gradients as leaves: with DMC_TRAIN_SIDE_STREAM=1 they run on a
  // second stream (forked and joined with events, inside the graph too); partial-sum buffers are per stream: partS for the side stream's own
  // kernels, partA / partB for the column sums the main-stream kernels k_chunkadd_fwd_bwd / k_wsilu_bwd leave behind
  float *partS = nullptr, *partA = nullptr, *partB = nullptr;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool use_side = false;
  size_t part_floats = 0;
  // caller tensors of the current call (read by the launch closures)
  const float *x = nullptr, *gout = nullptr, *yout = nullptr;
  float* scale2 = nullptr;       // {gradient scale, its reciprocal}, written on the device at the start of every backward
  float *out = nullptr, *gx = nullptr, *gqs = nullptr;
  float* gw[12] = {};            // caller's destinations (this call)
  float* gint(int i) const { return (gmask >> i & 1) ? gflat + goff[i] : nullptr; }
  ~dmc_dcb_train() {
    DeviceGuard dg(e.device);
    cudaDeviceSynchronize();
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
  }
};

namespace {
// the incoming gradient is scaled so that max |g| = 2^this before it enters the fp16 split planes
float grad_peak_log2() {
  static float v = -1000.0f;
  if (v < -999.0f) {
    const char* e = getenv("DMC_TRAIN_GRAD_PEAK_LOG2");
    v = e ? (float)atof(e) : 8.0f;
  }
  return v;
}
const char* const kDcbNames[6] = {"b.adaptor", "b.dc.0", "b.dc.2", "b.dc.3", "b.ffn.0", "b.ffn.2"};

// Packs the caller's parameters for the forward program (`backward` false) or for both.  With `unchanged` set the
// caller vouches that the parameter VALUES are the ones this handle packed last: only what is still missing is packed
// (a training step calls forward, then backward, on the same weights).
void dcb_train_load(dmc_dcb_train& t, const float* const* w12, const float* quant_step, bool backward, bool unchanged,
                    cudaStream_t st) {
  dmc_engine& e = t.e;
  if (!w12) fail("null weight list");
  if (t.blk->adaptor && (!w12[0] || !w12[1])) fail("adaptor weights required (cin != cout or force_adaptor)");
  for (int i = 0; i < 6; ++i) {
    if (i == 0 && !t.blk->adaptor) continue;
    if (!w12[2 * i] || !w12[2 * i + 1]) fail("missing weight %s", kDcbNames[i]);
  }
  if (t.has_qs) {
    if (!quant_step) fail("quant_step required: the block was created with has_quant_step");
    CUDA_OK(cudaMemcpyAsync(t.qs_table, quant_step, sizeof(float) * t.cout, cudaMemcpyDeviceToDevice, st));
  } else if (quant_step) {
    fail("quant_step given but the block was created without has_quant_step");
  }
  if (!unchanged) t.fwd_packed = t.bwd_packed = false;
  if (!t.fwd_packed) {
    for (int i = 0; i < 6; ++i) {
      if (i == 0 && !t.blk->adaptor) continue;
      set_slot(e, std::string(kDcbNames[i]) + ".weight", w12[2 * i], st);
      set_slot(e, std::string(kDcbNames[i]) + ".bias", w12[2 * i + 1], st);
    }
    t.fwd_packed = true;
  }
  if (!backward || t.bwd_packed) return;
  const int C = t.cout;
  set_slot(e, "P.ffn0.weight", w12[8], st);
  set_slot(e, "P.ffn0.bias", w12[9], st);
  // data-gradient weights: the transposes, packed straight from the forward weights
  struct TW { Conv* c; const float* w; };
  const TW tw[5] = {{t.T_ad, w12[0]}, {t.T_dc0, w12[2]}, {t.T_dc3, w12[6]}, {t.T_ffn0, w12[8]}, {t.T_ffn2, w12[10]}};
  for (const TW& w : tw)
    if (w.c) pack_gemm_weight(w.w, w.c->cout, w.c->cin, 1, 1, w.c->g, st, true);
  flip_dw_weight(t.blk->dw->w9c, t.w9c_flip, C, st);
  t.bwd_packed = true;
}
}  // namespace

extern "C" int dmc_dcb_train_create(int batch, int height, int width, int cin, int cout, int force_adaptor, int shortcut,
                                    int has_quant_step, int terms, dmc_dcb_train** out) {
  if (!out) return DMC_E_INVALID;
  *out = nullptr;
  dmc_dcb_train* t = nullptr;
  int rc = guarded(nullptr, [&] {
    if (batch < 1 || height < 1 || width < 1) fail("dmc_dcb_train_create: bad geometry");
    if (cin < 16 || cout < 32 || cin % 16 || cout % 16) fail("dmc_dcb_train_create: channels must be multiples of 16 (cin >= 16, cout >= 32)");
    if (terms != 1 && terms != 3) fail("dmc_dcb_train_create: terms must be 1 or 3");
    t = new dmc_dcb_train();
    dmc_engine& e = t->e;
    CUDA_OK(cudaGetDevice(&e.device));
    e.variant = -1; e.B = batch; e.H = height; e.W = width;
    t->cin = cin; t->cout = cout; t->shortcut = shortcut != 0; t->has_qs = has_quant_step != 0; t->terms = terms;
    const int C = cout, B = batch, H = height, W = width;
    const long long M = (long long)B * H * W;
    t->blk = e.add_dcb("b", cin, cout, force_adaptor != 0);
    const bool ad = t->blk->adaptor != nullptr;
    t->P_ffn0 = e.add_conv("P.ffn0", C, 4 * C, 1, 1, 0);
    if (ad) t->T_ad = e.add_conv("T.ad", C, cin, 1, 1, 0);
    t->T_dc0 = e.add_conv("T.dc0", C, C, 1, 1, 0);
    t->T_dc3 = e.add_conv("T.dc3", C, C, 1, 1, 0);
    t->T_ffn0 = e.add_conv("T.ffn0", 4 * C, C, 1, 1, 0);
    t->T_ffn2 = e.add_conv("T.ffn2", C, 2 * C, 1, 1, 0);
    for (Conv* c : {t->T_ad, t->T_dc0, t->T_dc3, t->T_ffn0, t->T_ffn2})
      if (c) pack_gemm_bias(nullptr, c->cout, c->g, nullptr);      // data gradients carry no bias
    t->w9c_flip = e.new_f32((size_t)9 * C);
    t->zero_bias = e.new_f32((size_t)4 * C);
    CUDA_OK(cudaMemset(t->zero_bias, 0, sizeof(float) * 4 * C));
    t->qs_table = e.new_f32(C);
    t->scale2 = e.new_f32(2);
    {
      const size_t sz[12] = {(size_t)C * cin, (size_t)C, (size_t)C * C, (size_t)C, (size_t)C * 9, (size_t)C, (size_t)C * C, (size_t)C,
                             (size_t)4 * C * C, (size_t)4 * C, (size_t)2 * C * C, (size_t)C};
      for (int i = 0; i < 12; ++i) t->goff[i + 1] = t->goff[i] + sz[i];
      t->gflat = e.new_f32(t->goff[12]);
    }
    // everything allocated from here on is workspace, shared with the other blocks of this geometry (TrainPool)
    {
      const char* v = getenv("DMC_TRAIN_SHARED_WORKSPACE");      // =0: every handle keeps its own buffers
      if (!(v && v[0] == '0')) {
        char key[160];
        snprintf(key, sizeof key, "dev%d:%dx%dx%d:%d>%d:a%d:s%d:q%d:t%d", e.device, B, H, W, cin, cout, ad ? 1 : 0,
                 t->shortcut, t->has_qs, terms);
        e.pool_prefix = key;
      }
    }
    // partial sums: the largest of the weight-gradient splits, the column sums and the depthwise partial rows
    const int max_parts = 8 * num_sms();
    const int ca_parts = chunkadd_parts(M, 2 * C);
    size_t pf = (size_t)max_parts * 4 * C;                      // column sums of the widest tensor
    pf = std::max(pf, (size_t)max_parts * C * 10);              // depthwise partial rows
    const int shapes[5][2] = {{C, cin}, {C, C}, {C, C}, {4 * C, C}, {C, 2 * C}};
    for (auto& sh : shapes) pf = std::max(pf, (size_t)wgrad_splits(M, sh[0], sh[1]) * sh[0] * sh[1]);
    t->partS = e.new_f32(pf);
    t->partA = e.new_f32((size_t)ca_parts * 4 * C);
    t->partB = e.new_f32((size_t)2 * max_parts * C);             // k_wsilu_bwd's rows; the head's scratch (absmax, nchw_dot)
    t->part = t->partB;
    t->part_floats = pf;
    {
      // =1: the leaves on a second stream.  Measured neutral (0.42 against 0.39 ms for a 20x30 block, 31.7 against 31.5 ms
      // for a full-size training step): small blocks are bound by the host side of a call, large ones fill the machine
      // with every kernel -- so the default keeps everything on the caller's stream
      const char* v = getenv("DMC_TRAIN_SIDE_STREAM");
      t->use_side = v && v[0] == '1';
      if (t->use_side) {
        CUDA_OK(cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
      }
    }

    // ---- forward program
    dmc_dcb_train* self = t;
    e.prog = &e.prog_common;
    {
      Act fin = e.new_act(B, H, W, cin), fout = e.new_act(B, H, W, C);
      e.op([self, fin, B, H, W, cin](cudaStream_t st) { nchw_to_s3(self->x, fin.v, B, cin, H, W, st); });
      e.cur.qp = 0;
      e.dcb(t->blk, fin, fout, t->shortcut, t->has_qs ? t->qs_table : nullptr, terms);
      e.op([self, fout, B, H, W, C](cudaStream_t st) { s3_to_nchw(fout.v, self->out, B, C, H, W, st); });
      e.flush_chain();
    }

    // ---- backward program
    e.set_prog(&t->prog_bwd_head);
    Act xs = e.new_act(B, H, W, cin);
    Act a = ad ? e.new_act(B, H, W, C) : xs;
    float* t0 = e.new_f32((size_t)M * C);
    float* t1 = e.new_f32((size_t)M * C);
    float* u0 = e.new_f32((size_t)M * 4 * C);
    float* gt2 = e.new_f32((size_t)M * C);
    Act t2 = e.new_act(B, H, W, C), o1 = e.new_act(B, H, W, C), v = e.new_act(B, H, W, 2 * C);
    Act g = e.new_act(B, H, W, C), gv = e.new_act(B, H, W, 2 * C), gu = e.new_act(B, H, W, 4 * C);
    Act go1 = e.new_act(B, H, W, C), gt1 = e.new_act(B, H, W, C), gt0 = e.new_act(B, H, W, C);
    Act ga = e.new_act(B, H, W, C);
    float* scale2 = t->scale2;
    EpiSpec plain;
    plain.nsplit = terms;
    // head: everything that reads a caller tensor.  The incoming gradient gets a power-of-two scale into fp16's range
    // (everything below is linear in it) and quant_step
    e.op([self, xs, B, H, W, cin](cudaStream_t st) { nchw_to_s3(self->x, xs.v, B, cin, H, W, st); });
    e.op([self, g, B, H, W, C, M, scale2, max_parts](cudaStream_t st) {
      grad_scale(self->gout, M * C, grad_peak_log2(), self->part, scale2, st);
      if (self->has_qs && self->gqs) {
        // out = out_pre * quant_step  ->  d/d quant_step[c] = sum g * out_pre = (sum g * out) / quant_step[c]
        if (!self->yout) fail("dmc_dcb_train_backward: the forward output is needed for grad_quant_step");
        int S = nchw_dot(self->gout, self->yout, B, C, (long long)H * W, self->part, max_parts, st);
        reduce_div(self->part, S, C, self->qs_table, self->gqs, st);
      }
      nchw_to_s3_scaled(self->gout, g.v, B, C, H, W, self->has_qs ? self->qs_table : nullptr, scale2, st);
    });
    // body: recompute (the forward pass keeps nothing but x and its own output), then the block in reverse
    e.set_prog(&t->prog_bwd);
    if (ad) e.gemm(xs, t->blk->adaptor, &a, plain);
    {
      EpiSpec s = plain;
      s.out_f32 = t0; s.ld_f32 = C;
      e.gemm(a, t->blk->dc0, nullptr, s);
    }
    DW* dw = t->blk->dw;
    e.op([t0, t1, M, C](cudaStream_t st) { wsilu_rows(t0, t1, M * C, st); });
    e.op([t1, C, t2, dw](cudaStream_t st) { dwconv3x3_f32(t1, C, dw->w9c, dw->bias, t2.v, t2.B, t2.H, t2.W, st); });
    {
      EpiSpec s = plain;
      s.res1 = &a;
      e.gemm(t2, t->blk->dc3, &o1, s);
    }
    {
      EpiSpec s = plain;
      s.out_f32 = u0; s.ld_f32 = 4 * C;
      e.gemm(o1, t->P_ffn0, nullptr, s);
    }
    // A leaf of the backward pass (nothing on the chain of data gradients waits for it): on the side stream, after
    // everything launched so far on the main one.  Every tensor of a pass is written once, so the only hazards are the
    // partial-sum buffers, and those are per stream.
    auto leaf = [&](std::function<void(cudaStream_t)> fn) {
      e.op([self, fn](cudaStream_t st) {
        if (!self->use_side) return fn(st);
        CUDA_OK(cudaEventRecord(self->ev_fork, st));
        CUDA_OK(cudaStreamWaitEvent(self->side, self->ev_fork, 0));
        fn(self->side);
      });
    };
    // one weight + bias gradient: G [M, N] against the layer input X [M, K]; `bias_done`: the column sums of G were
    // already reduced by the kernel that produced G
    auto wgrad = [&](const Act& G, const Act& X, int iw, bool bias_done) {
      const int terms_ = terms;
      leaf([self, G, X, M, iw, terms_, max_parts, bias_done, scale2](cudaStream_t st) {
        const int N = G.v.C, K = X.v.C;
        if (self->gint(iw)) {
          int S = wgrad_s3(G.v, X.v, M, terms_, self->partS, st);
          if (S < 1) fail("weight gradient launch: %s", wgrad_umma_last_error());
          reduce_partials(self->partS, (long long)N * K, S, self->gint(iw), (long long)N * K, scale2 + 1, 1.0f, st);
        }
        if (self->gint(iw + 1) && !bias_done) {
          int S = colsum_s3(G.v, nullptr, M, self->partS, N, max_parts, st);
          reduce_partials(self->partS, N, S, self->gint(iw + 1), N, scale2 + 1, 1.0f, st);
        }
      });
    };
    // ffn.2: data gradient first, then ONE pass over the pre-activations gives v (operand of ffn.2's weight gradient),
    // the gradient of the pre-activations and ffn.0's bias gradient
    e.gemm(g, t->T_ffn2, &gv, plain);
    e.op([self, u0, gv, v, gu, M, C, ca_parts](cudaStream_t st) {
      int S = chunkadd_fwd_bwd(u0, 4 * C, gv.v, v.v, gu.v, M, self->partA, 4 * C, ca_parts, st);
      if (S != ca_parts) fail("chunkadd_fwd_bwd: partial buffer too small");
    });
    leaf([self, C, ca_parts, scale2](cudaStream_t st) {
      if (self->gint(9)) reduce_partials(self->partA, 4 * C, ca_parts, self->gint(9), 4 * C, scale2 + 1, 1.0f, st);
    });
    wgrad(g, v, 10, false);
    // ffn.0 (+ the residual around the ffn)
    wgrad(gu, o1, 8, true);
    {
      EpiSpec s = plain;
      s.res1 = &g;
      e.gemm(gu, t->T_ffn0, &go1, s);
    }
    // dc.3: its data gradient feeds the depthwise kernels only -> fp32 rows
    wgrad(go1, t2, 6, false);
    {
      EpiSpec s = plain;
      s.out_f32 = gt2; s.ld_f32 = C;
      e.gemm(go1, t->T_dc3, nullptr, s);
    }
    // depthwise 3x3
    e.op([self, gt2, gt1, C, B, H, W](cudaStream_t st) {
      dwconv3x3_f32(gt2, C, self->w9c_flip, self->zero_bias, gt1.v, B, H, W, st);
    });
    leaf([self, gt2, t1, C, B, H, W, max_parts, scale2](cudaStream_t st) {
      if (self->gint(4) || self->gint(5)) {
        int S = dw_wgrad(gt2, C, C, t1, C, B, H, W, self->partS, C * 10, max_parts, st);
        reduce_dw(self->partS, C * 10, S, self->gint(4), self->gint(5), C, scale2 + 1, 1.0f, st);
      }
    });
    const int wb_parts = wsilu_bwd_parts(M, C, 2 * max_parts);
    e.op([self, gt1, t0, gt0, M, C, max_parts, wb_parts](cudaStream_t st) {
      int S = wsilu_bwd(gt1.v, t0, C, gt0.v, M, self->partB, C, 2 * max_parts, st);
      if (S != wb_parts) fail("wsilu_bwd: unexpected number of partial rows");
    });
    leaf([self, C, wb_parts, scale2](cudaStream_t st) {
      if (self->gint(3)) reduce_partials(self->partB, C, wb_parts, self->gint(3), C, scale2 + 1, 1.0f, st);
    });
    // dc.0 (+ the residual around dc, + the shortcut)
    wgrad(gt0, a, 2, true);
    {
      EpiSpec s = plain;
      s.res1 = &go1;
      if (t->shortcut) s.res2 = &g;
      e.gemm(gt0, t->T_dc0, &ga, s);
    }
    Act gxs = ad ? e.new_act(B, H, W, cin) : ga;
    if (ad) {
      wgrad(ga, xs, 0, false);
      e.gemm(ga, t->T_ad, &gxs, plain);
    }
    // join: the pass is over when both streams are
    e.op([self](cudaStream_t st) {
      if (!self->use_side) return;
      CUDA_OK(cudaEventRecord(self->ev_join, self->side));
      CUDA_OK(cudaStreamWaitEvent(st, self->ev_join, 0));
    });
    e.set_prog(&t->prog_bwd_tail);
    {
      const int cx = ad ? cin : C;
      e.op([self, gxs, B, H, W, cx, scale2](cudaStream_t st) {
        if (self->gx) s3_to_nchw_scaled(gxs.v, self->gx, B, cx, H, W, scale2 + 1, st);
      });
    }
    e.flush_chain();
    e.prog = &e.prog_common;
    CUDA_OK(cudaDeviceSynchronize());
  });
  if (rc != DMC_OK) {
    delete t;
    return rc;
  }
  *out = t;
  return DMC_OK;
}

extern "C" void dmc_dcb_train_destroy(dmc_dcb_train* t) { delete t; }

extern "C" const char* dmc_dcb_train_last_error(const dmc_dcb_train* t) {
  return t ? t->e.error.c_str() : g_create_error.c_str();
}

extern "C" int dmc_dcb_train_forward(dmc_dcb_train* t, const float* x, const float* const* weights12, const float* quant_step,
                                     float* out, int weights_unchanged, void* stream) {
  if (!t) return DMC_E_INVALID;
  return guarded(&t->e, [&] {
    if (!x || !out) fail("dmc_dcb_train_forward: null tensor");
    DeviceGuard dg(t->e.device);
    cudaStream_t st = (cudaStream_t)stream;
    dcb_train_load(*t, weights12, quant_step, false, weights_unchanged != 0, st);
    t->x = x; t->out = out;
    t->e.cur.qp = 0;
    t->e.run(t->e.prog_common, st);
    CUDA_OK(cudaGetLastError());
  });
}

extern "C" int dmc_dcb_train_backward(dmc_dcb_train* t, const float* x, const float* const* weights12, const float* quant_step,
                                      const float* out, const float* grad_out, float* grad_x,
                                      float* const* grad_weights12, float* grad_quant_step, int weights_unchanged,
                                      void* stream) {
  if (!t) return DMC_E_INVALID;
  return guarded(&t->e, [&] {
    if (!x || !grad_out) fail("dmc_dcb_train_backward: null tensor");
    DeviceGuard dg(t->e.device);
    cudaStream_t st = (cudaStream_t)stream;
    dcb_train_load(*t, weights12, quant_step, true, weights_unchanged != 0, st);
    t->x = x; t->gout = grad_out; t->yout = out; t->gx = grad_x; t->gqs = t->has_qs ? grad_quant_step : nullptr;
    if (t->gqs && !out) fail("dmc_dcb_train_backward: `out` (the forward output) is required for grad_quant_step");
    if (((uintptr_t)grad_out | (uintptr_t)x) % 16) fail("dmc_dcb_train_backward: tensors must be 16-byte aligned");
    for (int i = 0; i < 12; ++i) t->gw[i] = grad_weights12 ? grad_weights12[i] : nullptr;
    if (!t->blk->adaptor) t->gw[0] = t->gw[1] = nullptr;
    t->gmask = 0;
    for (int i = 0; i < 12; ++i)
      if (t->gw[i]) t->gmask |= 1 << i;
    dmc_engine& e = t->e;
    e.cur.qp = 0;
    e.run(t->prog_bwd_head, st);
    // the body touches only the handle's buffers: one graph per set of requested gradients, from the second call on
    // (the first call runs directly: one-time allocations inside the launchers are not capturable)
    if (t->bwd_calls++ > 0) e.run_graph(0x100000ull | (uint64_t)t->gmask, st, [&](cudaStream_t s2) { e.run(t->prog_bwd, s2); });
    else e.run(t->prog_bwd, st);
    e.run(t->prog_bwd_tail, st);
    // gradients leave in as few copies as the caller's layout allows (training.py: one flat tensor in parameter order)
    for (int i = 0; i < 12;) {
      if (!t->gw[i]) { ++i; continue; }
      int j = i;
      while (j + 1 < 12 && t->gw[j + 1] && t->gw[j + 1] == t->gw[j] + (t->goff[j + 1] - t->goff[j])) ++j;
      CUDA_OK(cudaMemcpyAsync(t->gw[i], t->gflat + t->goff[i], (t->goff[j + 1] - t->goff[i]) * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
      i = j + 1;
    }
    CUDA_OK(cudaGetLastError());
  });
}

// ---------------------------------------------------------------- training mode: a plain 1x1 convolution
// The 1x1 convolutions of the models OUTSIDE the DepthConvBlocks (feature_adaptor_p, encoder.conv1, decoder.proj, the
// recon head, the sub-pixel convolutions' 1x1 kernels, the prior heads): same pieces as the block -- forward = one
// contraction of the frame engine, data gradient = the contraction with the transposed weight, weight gradient =
// k_wgrad_umma, bias gradient = a column sum -- so that a training step is fp32-grade end to end without paying
// cuDNN's fp32 (TF32 off) rate for them.
struct dmc_conv1x1_train {
  dmc_engine e;
  std::vector<dmc_engine::Op> prog_bwd_head, prog_bwd, prog_bwd_tail;
  int cin = 0, cout = 0, terms = 3;
  bool has_bias = true, packed = false, packed_T = false;
  Conv *fw = nullptr, *T = nullptr;
  float *gflat = nullptr, *scale2 = nullptr, *partS = nullptr, *partB = nullptr;
  int gmask = 0;                 // bit 0: weight gradient, bit 1: bias gradient, bit 2: input gradient
  long long bwd_calls = 0;
  const float *x = nullptr, *gout = nullptr;
  float *out = nullptr, *gx = nullptr, *gw = nullptr, *gb = nullptr;
  ~dmc_conv1x1_train() {
    DeviceGuard dg(e.device);
    cudaDeviceSynchronize();
  }
};

extern "C" int dmc_conv1x1_train_create(int batch, int height, int width, int cin, int cout, int has_bias, int terms,
                                        dmc_conv1x1_train** out) {
  if (!out) return DMC_E_INVALID;
  *out = nullptr;
  dmc_conv1x1_train* t = nullptr;
  int rc = guarded(nullptr, [&] {
    if (batch < 1 || height < 1 || width < 1) fail("dmc_conv1x1_train_create: bad geometry");
    if (cin < 32 || cout < 32 || cin % 16 || cout % 16) fail("dmc_conv1x1_train_create: channels must be multiples of 16, at least 32");
    if (terms != 1 && terms != 3) fail("dmc_conv1x1_train_create: terms must be 1 or 3");
    t = new dmc_conv1x1_train();
    dmc_engine& e = t->e;
    CUDA_OK(cudaGetDevice(&e.device));
    e.variant = -1; e.B = batch; e.H = height; e.W = width;
    t->cin = cin; t->cout = cout; t->terms = terms; t->has_bias = has_bias != 0;
    const int B = batch, H = height, W = width;
    const long long M = (long long)B * H * W;
    t->fw = e.add_conv("c", cin, cout, 1, 1, 0);
    t->T = e.add_conv("T", cout, cin, 1, 1, 0);
    pack_gemm_bias(nullptr, cin, t->T->g, nullptr);
    pack_gemm_bias(nullptr, cout, t->fw->g, nullptr);
    t->gflat = e.new_f32((size_t)cout * cin + cout);
    t->scale2 = e.new_f32(2);
    {
      const char* v = getenv("DMC_TRAIN_SHARED_WORKSPACE");
      if (!(v && v[0] == '0')) {
        char key[160];
        snprintf(key, sizeof key, "conv1x1:dev%d:%dx%dx%d:%d>%d:t%d", e.device, B, H, W, cin, cout, terms);
        e.pool_prefix = key;
      }
    }
    const int max_parts = 8 * num_sms();
    t->partS = e.new_f32(std::max((size_t)max_parts * cout, (size_t)wgrad_splits(M, cout, cin) * cout * cin));
    t->partB = e.new_f32((size_t)max_parts);
    dmc_conv1x1_train* self = t;
    EpiSpec plain;
    plain.nsplit = terms;
    // ---- forward
    e.prog = &e.prog_common;
    {
      Act fin = e.new_act(B, H, W, cin), fout = e.new_act(B, H, W, cout);
      e.op([self, fin, B, H, W, cin](cudaStream_t st) { nchw_to_s3(self->x, fin.v, B, cin, H, W, st); });
      e.gemm(fin, t->fw, &fout, plain);
      e.op([self, fout, B, H, W, cout](cudaStream_t st) { s3_to_nchw(fout.v, self->out, B, cout, H, W, st); });
      e.flush_chain();
    }
    // ---- backward: head (caller tensors -> planes), body (the handle's buffers only: one graph), tail
    Act xs = e.new_act(B, H, W, cin), g = e.new_act(B, H, W, cout), gxs = e.new_act(B, H, W, cin);
    float* scale2 = t->scale2;
    e.set_prog(&t->prog_bwd_head);
    e.op([self, xs, g, B, H, W, cin, cout, M, scale2](cudaStream_t st) {
      if (self->gmask & 1) nchw_to_s3(self->x, xs.v, B, cin, H, W, st);
      grad_scale(self->gout, M * cout, grad_peak_log2(), self->partB, scale2, st);
      nchw_to_s3_scaled(self->gout, g.v, B, cout, H, W, nullptr, scale2, st);
    });
    e.set_prog(&t->prog_bwd);
    {
      const int terms_ = terms;
      e.op([self, g, xs, M, cin, cout, terms_, max_parts, scale2](cudaStream_t st) {
        if (self->gmask & 1) {
          int S = wgrad_s3(g.v, xs.v, M, terms_, self->partS, st);
          if (S < 1) fail("weight gradient launch: %s", wgrad_umma_last_error());
          reduce_partials(self->partS, (long long)cout * cin, S, self->gflat, (long long)cout * cin, scale2 + 1, 1.0f, st);
        }
        if (self->gmask & 2) {
          int S = colsum_s3(g.v, nullptr, M, self->partS, cout, max_parts, st);
          reduce_partials(self->partS, cout, S, self->gflat + (size_t)cout * cin, cout, scale2 + 1, 1.0f, st);
        }
      });
    }
    // (the data gradient is staged unconditionally; the body graph is keyed by the mask and skips it when unwanted)
    e.gemm(g, t->T, &gxs, plain);
    e.set_prog(&t->prog_bwd_tail);
    e.op([self, gxs, B, H, W, cin, scale2](cudaStream_t st) {
      if (self->gx) s3_to_nchw_scaled(gxs.v, self->gx, B, cin, H, W, scale2 + 1, st);
    });
    e.flush_chain();
    e.prog = &e.prog_common;
    CUDA_OK(cudaDeviceSynchronize());
  });
  if (rc != DMC_OK) {
    delete t;
    return rc;
  }
  *out = t;
  return DMC_OK;
}

extern "C" void dmc_conv1x1_train_destroy(dmc_conv1x1_train* t) { delete t; }
extern "C" const char* dmc_conv1x1_train_last_error(const dmc_conv1x1_train* t) {
  return t ? t->e.error.c_str() : g_create_error.c_str();
}

namespace {
void conv1x1_train_load(dmc_conv1x1_train& t, const float* weight, const float* bias, bool backward, bool need_T,
                        bool unchanged, cudaStream_t st) {
  if (!unchanged) t.packed = t.packed_T = false;
  if (backward && !need_T) return;
  if (!weight) fail("null weight");
  if (t.has_bias && !backward && !bias) fail("bias required: the handle was created with has_bias");
  if (!backward && !t.packed) {
    pack_gemm_weight(weight, t.cout, t.cin, 1, 1, t.fw->g, st);
    if (t.has_bias) pack_gemm_bias(bias, t.cout, t.fw->g, st);
    t.packed = true;
  }
  if (backward && !t.packed_T) {
    pack_gemm_weight(weight, t.cin, t.cout, 1, 1, t.T->g, st, true);
    t.packed_T = true;
  }
}
}  // namespace

extern "C" int dmc_conv1x1_train_forward(dmc_conv1x1_train* t, const float* x, const float* weight, const float* bias,
                                         float* out, int weights_unchanged, void* stream) {
  if (!t) return DMC_E_INVALID;
  return guarded(&t->e, [&] {
    if (!x || !out) fail("dmc_conv1x1_train_forward: null tensor");
    DeviceGuard dg(t->e.device);
    cudaStream_t st = (cudaStream_t)stream;
    conv1x1_train_load(*t, weight, bias, false, false, weights_unchanged != 0, st);
    t->x = x; t->out = out;
    t->e.cur.qp = 0;
    t->e.run(t->e.prog_common, st);
    CUDA_OK(cudaGetLastError());
  });
}

extern "C" int dmc_conv1x1_train_backward(dmc_conv1x1_train* t, const float* x, const float* weight, const float* grad_out,
                                          float* grad_x, float* grad_weight, float* grad_bias, int weights_unchanged,
                                          void* stream) {
  if (!t) return DMC_E_INVALID;
  return guarded(&t->e, [&] {
    if (!grad_out || (grad_weight && !x)) fail("dmc_conv1x1_train_backward: null tensor");
    if ((uintptr_t)grad_out % 16) fail("dmc_conv1x1_train_backward: grad_out must be 16-byte aligned");
    DeviceGuard dg(t->e.device);
    cudaStream_t st = (cudaStream_t)stream;
    conv1x1_train_load(*t, weight, nullptr, true, grad_x != nullptr, weights_unchanged != 0, st);
    t->x = x; t->gout = grad_out; t->gx = grad_x; t->gw = grad_weight; t->gb = grad_bias;
    t->gmask = (grad_weight ? 1 : 0) | (grad_bias ? 2 : 0) | (grad_x ? 4 : 0);
    dmc_engine& e = t->e;
    e.cur.qp = 0;
    e.run(t->prog_bwd_head, st);
    // body: [0] = weight / bias gradients, [1] = the data-gradient chain (skipped when the input needs no gradient)
    auto body = [&](cudaStream_t s2) {
      for (size_t i = 0; i < t->prog_bwd.size(); ++i)
        if (i == 0 || (t->gmask & 4)) t->prog_bwd[i](s2);
    };
    if (t->bwd_calls++ > 0) e.run_graph(0x200000ull | (uint64_t)t->gmask, st, body);
    else body(st);
    e.run(t->prog_bwd_tail, st);
    const size_t wn = (size_t)t->cout * t->cin;
    if (grad_weight && grad_bias && grad_bias == grad_weight + wn) {
      CUDA_OK(cudaMemcpyAsync(grad_weight, t->gflat, (wn + t->cout) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      if (grad_weight) CUDA_OK(cudaMemcpyAsync(grad_weight, t->gflat, wn * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if (grad_bias) CUDA_OK(cudaMemcpyAsync(grad_bias, t->gflat + wn, t->cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    CUDA_OK(cudaGetLastError());
  });
}

// ---------------------------------------------------------------- training mode: k x k convolutions
// encoder.down / mask_sft.down (3x3 stride 2), the sub-pixel 3x3 of decoder.up, the 2x2 stride-2 downs of the hyper path
// and the temporal prior: forward = im2col + one contraction (as in the frame engine); backward through the same im2col
// view: weight gradient = k_wgrad_umma over (gradient rows, im2col rows), data gradient = the contraction with the
// transposed weight followed by the gather that undoes im2col (k_col2im_nchw).
struct dmc_convkxk_train {
  dmc_engine e;
  std::vector<dmc_engine::Op> prog_bwd_head, prog_bwd, prog_bwd_tail;
  int cin = 0, cout = 0, k = 1, stride = 1, pad = 0, terms = 3;
  bool has_bias = true, packed = false, packed_T = false;
  Conv *fw = nullptr, *T = nullptr;
  float *gflat = nullptr, *scale2 = nullptr, *partS = nullptr, *partB = nullptr, *wt = nullptr;
  int gmask = 0;                 // bit 0: weight gradient, bit 1: bias gradient, bit 2: input gradient
  long long bwd_calls = 0;
  const float *x = nullptr, *gout = nullptr;
  float *out = nullptr, *gx = nullptr;
  ~dmc_convkxk_train() {
    DeviceGuard dg(e.device);
    cudaDeviceSynchronize();
  }
};

extern "C" int dmc_convkxk_train_create(int batch, int height, int width, int cin, int cout, int ksize, int stride, int padding,
                                        int has_bias, int terms, dmc_convkxk_train** out) {
  if (!out) return DMC_E_INVALID;
  *out = nullptr;
  dmc_convkxk_train* t = nullptr;
  int rc = guarded(nullptr, [&] {
    if (batch < 1 || height < 1 || width < 1) fail("dmc_convkxk_train_create: bad geometry");
    if (ksize < 2 || ksize > 3 || stride < 1 || stride > 2 || padding < 0 || padding > 1) fail("dmc_convkxk_train_create: kernel 2 or 3, stride 1 or 2, padding 0 or 1");
    if (cin < 16 || cout < 32 || cin % 16 || cout % 16) fail("dmc_convkxk_train_create: channels must be multiples of 16 (cout >= 32)");
    if (terms != 1 && terms != 3) fail("dmc_convkxk_train_create: terms must be 1 or 3");
    const int B = batch, H = height, W = width, k = ksize;
    const int Ho = (H + 2 * padding - k) / stride + 1, Wo = (W + 2 * padding - k) / stride + 1;
    if (Ho < 1 || Wo < 1) fail("dmc_convkxk_train_create: empty output");
    t = new dmc_convkxk_train();
    dmc_engine& e = t->e;
    CUDA_OK(cudaGetDevice(&e.device));
    e.variant = -1; e.B = B; e.H = H; e.W = W;
    t->cin = cin; t->cout = cout; t->k = k; t->stride = stride; t->pad = padding; t->terms = terms; t->has_bias = has_bias != 0;
    const long long Mo = (long long)B * Ho * Wo;
    const int K = k * k * cin;
    t->fw = e.add_conv("c", cin, cout, k, stride, padding);
    t->T = e.add_conv("T", cout, K, 1, 1, 0);
    pack_gemm_bias(nullptr, K, t->T->g, nullptr);
    pack_gemm_bias(nullptr, cout, t->fw->g, nullptr);
    t->gflat = e.new_f32((size_t)cout * K + cout);
    t->wt = e.new_f32((size_t)cout * K);
    t->scale2 = e.new_f32(2);
    {
      const char* v = getenv("DMC_TRAIN_SHARED_WORKSPACE");
      if (!(v && v[0] == '0')) {
        char key[200];
        snprintf(key, sizeof key, "convkxk:dev%d:%dx%dx%d:%d>%d:k%ds%dp%d:t%d", e.device, B, H, W, cin, cout, k, stride, padding, terms);
        e.pool_prefix = key;
      }
    }
    const int max_parts = 8 * num_sms();
    t->partS = e.new_f32(std::max((size_t)max_parts * cout, (size_t)wgrad_splits(Mo, cout, K) * cout * K));
    t->partB = e.new_f32((size_t)max_parts);
    dmc_convkxk_train* self = t;
    EpiSpec plain;
    plain.nsplit = terms;
    // ---- forward
    e.prog = &e.prog_common;
    {
      Act fin = e.new_act(B, H, W, cin), fout = e.new_act(B, Ho, Wo, cout);
      e.op([self, fin, B, H, W, cin](cudaStream_t st) { nchw_to_s3(self->x, fin.v, B, cin, H, W, st); });
      e.conv_kxk(fin, t->fw, &fout, plain);
      e.op([self, fout, B, Ho, Wo, cout](cudaStream_t st) { s3_to_nchw(fout.v, self->out, B, cout, Ho, Wo, st); });
      e.flush_chain();
    }
    // ---- backward
    Act xs = e.new_act(B, H, W, cin), g = e.new_act(B, Ho, Wo, cout), gcol = e.new_act(B, Ho, Wo, K);
    Act col = e.get_scratch("im2col", B, Ho, Wo, K);          // (the forward's scratch: same geometry)
    float* scale2 = t->scale2;
    e.set_prog(&t->prog_bwd_head);
    e.op([self, xs, g, B, H, W, Ho, Wo, cin, cout, Mo, scale2](cudaStream_t st) {
      if (self->gmask & 1) nchw_to_s3(self->x, xs.v, B, cin, H, W, st);
      grad_scale(self->gout, Mo * cout, grad_peak_log2(), self->partB, scale2, st);
      nchw_to_s3_scaled(self->gout, g.v, B, cout, Ho, Wo, nullptr, scale2, st);
    });
    e.set_prog(&t->prog_bwd);
    {
      const int terms_ = terms;
      e.op([self, g, xs, col, B, H, W, Ho, Wo, Mo, cin, cout, k, stride, padding, K, terms_, max_parts, scale2](cudaStream_t st) {
        if (self->gmask & 1) {
          im2col(xs.v, col.v, B, H, W, k, stride, padding, Ho, Wo, cin, 0, st);
          int S = wgrad_s3(g.v, col.v, Mo, terms_, self->partS, st);
          if (S < 1) fail("weight gradient launch: %s", wgrad_umma_last_error());
          reduce_wgrad_kxk(self->partS, (long long)cout * K, S, self->gflat, cout, cin, k, scale2 + 1, st);
        }
        if (self->gmask & 2) {
          int S = colsum_s3(g.v, nullptr, Mo, self->partS, cout, max_parts, st);
          reduce_partials(self->partS, cout, S, self->gflat + (size_t)cout * K, cout, scale2 + 1, 1.0f, st);
        }
      });
    }
    e.gemm(g, t->T, &gcol, plain);
    e.set_prog(&t->prog_bwd_tail);
    e.op([self, gcol, B, H, W, Ho, Wo, cin, k, stride, padding, scale2](cudaStream_t st) {
      if (self->gx) col2im_nchw(gcol.v, self->gx, B, cin, H, W, k, stride, padding, Ho, Wo, scale2 + 1, st);
    });
    e.flush_chain();
    e.prog = &e.prog_common;
    CUDA_OK(cudaDeviceSynchronize());
  });
  if (rc != DMC_OK) {
    delete t;
    return rc;
  }
  *out = t;
  return DMC_OK;
}

extern "C" void dmc_convkxk_train_destroy(dmc_convkxk_train* t) { delete t; }
extern "C" const char* dmc_convkxk_train_last_error(const dmc_convkxk_train* t) {
  return t ? t->e.error.c_str() : g_create_error.c_str();
}

extern "C" int dmc_convkxk_train_forward(dmc_convkxk_train* t, const float* x, const float* weight, const float* bias,
                                         float* out, int weights_unchanged, void* stream) {
  if (!t) return DMC_E_INVALID;
  return guarded(&t->e, [&] {
    if (!x || !out || !weight) fail("dmc_convkxk_train_forward: null tensor");
    if (t->has_bias && !bias) fail("bias required: the handle was created with has_bias");
    DeviceGuard dg(t->e.device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!weights_unchanged) t->packed = t->packed_T = false;
    if (!t->packed) {
      pack_gemm_weight(weight, t->cout, t->cin, t->k, t->k, t->fw->g, st);
      if (t->has_bias) pack_gemm_bias(bias, t->cout, t->fw->g, st);
      t->packed = true;
    }
    t->x = x; t->out = out;
    t->e.cur.qp = 0;
    t->e.run(t->e.prog_common, st);
    CUDA_OK(cudaGetLastError());
  });
}

extern "C" int dmc_convkxk_train_backward(dmc_convkxk_train* t, const float* x, const float* weight, const float* grad_out,
                                          float* grad_x, float* grad_weight, float* grad_bias, int weights_unchanged,
                                          void* stream) {
  if (!t) return DMC_E_INVALID;
  return guarded(&t->e, [&] {
    if (!grad_out || (grad_weight && !x) || (grad_x && !weight)) fail("dmc_convkxk_train_backward: null tensor");
    if ((uintptr_t)grad_out % 16) fail("dmc_convkxk_train_backward: grad_out must be 16-byte aligned");
    DeviceGuard dg(t->e.device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!weights_unchanged) t->packed = t->packed_T = false;
    const int K = t->k * t->k * t->cin;
    if (grad_x && !t->packed_T) {
      wt_kxk(weight, t->wt, t->cout, t->cin, t->k, st);
      pack_gemm_weight(t->wt, K, t->cout, 1, 1, t->T->g, st);
      t->packed_T = true;
    }
    t->x = x; t->gout = grad_out; t->gx = grad_x;
    t->gmask = (grad_weight ? 1 : 0) | (grad_bias ? 2 : 0) | (grad_x ? 4 : 0);
    dmc_engine& e = t->e;
    e.cur.qp = 0;
    e.run(t->prog_bwd_head, st);
    auto body = [&](cudaStream_t s2) {
      for (size_t i = 0; i < t->prog_bwd.size(); ++i)
        if (i == 0 || (t->gmask & 4)) t->prog_bwd[i](s2);
    };
    if (t->bwd_calls++ > 0) e.run_graph(0x300000ull | (uint64_t)t->gmask, st, body);
    else body(st);
    e.run(t->prog_bwd_tail, st);
    const size_t wn = (size_t)t->cout * K;
    if (grad_weight && grad_bias && grad_bias == grad_weight + wn) {
      CUDA_OK(cudaMemcpyAsync(grad_weight, t->gflat, (wn + t->cout) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      if (grad_weight) CUDA_OK(cudaMemcpyAsync(grad_weight, t->gflat, wn * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if (grad_bias) CUDA_OK(cudaMemcpyAsync(grad_bias, t->gflat + wn, t->cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    CUDA_OK(cudaGetLastError());
  });
}

extern "C" int dmc_op_quant_train(const float* x, const float* noise, float* out, int64_t n, int mode, void* stream) {
  if (!x || !out || n < 0 || (mode != 0 && mode != 1) || (mode == 1 && !noise)) return DMC_E_INVALID;
  if (n) quant_train(x, noise, out, (long long)n, mode, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? DMC_OK : DMC_E_CUDA;
}

extern "C" int dmc_op_gaussian_bits_backward(const float* sym, const float* sigma, const float* grad_bits, float* grad_sym,
                                             float* grad_sigma, int64_t n, int formula, void* stream) {
  if (!sym || !sigma || !grad_bits || !grad_sym || !grad_sigma || n < 0) return DMC_E_INVALID;
  if (n) gaussian_bits_bwd(sym, sigma, grad_bits, grad_sym, grad_sigma, (long long)n, formula, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? DMC_OK : DMC_E_CUDA;
}

extern "C" int dmc_bench_gemm(int rows, int k, int n, int mode, int nsplit, int pair, int iters, int probe,
                              float* ms_per_launch) {
  return guarded(nullptr, [&] {
    if (rows < 1 || k < 8 || n < 8 || iters < 1 || !ms_per_launch) fail("dmc_bench_gemm: bad arguments");
    dmc_engine e;
    e.variant = -1; e.B = 1; e.H = 1; e.W = rows;
    e.prog = &e.prog_common;
    e.use_s3 = pair == 2;
    Conv* c = e.add_conv("w", k, n, 1, 1, 0, mode == 3 ? PACK_PAIR : PACK_PLAIN);
    CUDA_OK(cudaMemset(c->g.w, 0, (size_t)kPlanes * c->g.Npad * c->g.Kld * sizeof(h16)));
    CUDA_OK(cudaMemset(c->g.wb, 0, (size_t)kPlanes * c->g.Npad * c->g.Kld * sizeof(h16)));
    CUDA_OK(cudaMemset(c->g.bias, 0, sizeof(float) * c->g.Npad));
    Act in = e.new_act(1, 1, rows, k);
    Act out = e.new_act(1, 1, rows, mode == 3 ? n / 2 : n);
    Act res = e.new_act(1, 1, rows, n);
    CUDA_OK(cudaMemset(in.v.p, 0, (size_t)in.v.ps * kPlanes * sizeof(h16)));
    CUDA_OK(cudaMemset(res.v.p, 0, (size_t)res.v.ps * kPlanes * sizeof(h16)));
    EpiSpec s;
    s.nsplit = nsplit;
    if (mode == 1 || mode == 3) s.act = ACT_WSILU;
    if (mode == 2) s.res1 = &res;
    gemm_s3_set_debug(probe);          // (the blob-store probe is decided when the chain is built)
    e.gemm(in, c, &out, s);
    e.flush_chain();
    umma_set_pair(pair != 0);
    umma_set_debug(probe);
    cudaEvent_t a, b;
    CUDA_OK(cudaEventCreate(&a));
    CUDA_OK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) e.run(e.prog_common, 0);
    CUDA_OK(cudaEventRecord(a, 0));
    for (int i = 0; i < iters; ++i) e.run(e.prog_common, 0);
    CUDA_OK(cudaEventRecord(b, 0));
    cudaError_t err = cudaEventSynchronize(b);
    umma_set_debug(0);
    gemm_s3_set_debug(0);
    umma_set_pair(true);
    CUDA_OK(err);
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, a, b));
    *ms_per_launch = ms / iters;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  });
}

extern "C" int dmc_bench_dwconv(int batch, int height, int width, int channels, int f32_in, int iters,
                                float* ms_per_launch) {
  return guarded(nullptr, [&] {
    if (batch < 1 || height < 1 || width < 1 || channels % 8 || iters < 1 || !ms_per_launch)
      fail("dmc_bench_dwconv: bad arguments");
    dmc_engine e;
    e.variant = -1; e.B = batch; e.H = height; e.W = width;
    Act in = e.new_act(batch, height, width, channels), out = e.new_act(batch, height, width, channels);
    CUDA_OK(cudaMemset(in.v.p, 0, (size_t)in.v.ps * kPlanes * sizeof(h16)));
    float* w = e.new_f32((size_t)9 * channels);
    float* b = e.new_f32(channels);
    CUDA_OK(cudaMemset(w, 0, sizeof(float) * 9 * channels));
    CUDA_OK(cudaMemset(b, 0, sizeof(float) * channels));
    cudaEvent_t t0, t1;
    CUDA_OK(cudaEventCreate(&t0));
    CUDA_OK(cudaEventCreate(&t1));
    float* in32 = e.new_f32((size_t)in.M() * channels);
    CUDA_OK(cudaMemset(in32, 0, sizeof(float) * in.M() * channels));
    auto run = [&] {
      if (f32_in) dwconv3x3_f32(in32, channels, w, b, out.v, batch, height, width, 0);
      else dwconv3x3(in.v, w, b, out.v, batch, height, width, 0);
    };
    for (int i = 0; i < 3; ++i) run();
    CUDA_OK(cudaEventRecord(t0, 0));
    for (int i = 0; i < iters; ++i) run();
    CUDA_OK(cudaEventRecord(t1, 0));
    CUDA_OK(cudaEventSynchronize(t1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, t0, t1));
    *ms_per_launch = ms / iters;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
  });
}

extern "C" int dmc_bench_dcb(int batch, int height, int width, int cin, int cout, int blocks, int iters,
                             float* ms_per_block) {
  return guarded(nullptr, [&] {
    if (batch < 1 || height < 1 || width < 1 || cin % 8 || cout % 8 || blocks < 1 || iters < 1 || !ms_per_block)
      fail("dmc_bench_dcb: bad arguments");
    dmc_engine e;
    e.variant = -1; e.B = batch; e.H = height; e.W = width;
    e.prog = &e.prog_common;
    Act a = e.new_act(batch, height, width, cin), b = e.new_act(batch, height, width, cout),
        c = e.new_act(batch, height, width, cout);
    CUDA_OK(cudaMemset(a.v.p, 0, (size_t)a.v.ps * kPlanes * sizeof(h16)));
    for (int i = 0; i < blocks; ++i) {
      DCB* w = e.add_dcb("b" + std::to_string(i), i == 0 ? cin : cout, cout);
      e.dcb(w, i == 0 ? a : (i % 2 ? b : c), i % 2 ? c : b, false, nullptr, 3);
    }
    e.flush_chain();
    for (auto& cv : e.convs) {
      CUDA_OK(cudaMemset(cv->g.w, 0, (size_t)kPlanes * cv->g.Npad * cv->g.Kld * sizeof(h16)));
      CUDA_OK(cudaMemset(cv->g.wb, 0, (size_t)kPlanes * cv->g.Npad * cv->g.Kld * sizeof(h16)));
      CUDA_OK(cudaMemset(cv->g.bias, 0, sizeof(float) * cv->g.Npad));
    }
    for (auto& d : e.dws) {
      CUDA_OK(cudaMemset(d->w9c, 0, sizeof(float) * 9 * d->C));
      CUDA_OK(cudaMemset(d->bias, 0, sizeof(float) * d->C));
    }
    cudaEvent_t t0, t1;
    CUDA_OK(cudaEventCreate(&t0));
    CUDA_OK(cudaEventCreate(&t1));
    for (int i = 0; i < 2; ++i) e.run(e.prog_common, 0);
    CUDA_OK(cudaEventRecord(t0, 0));
    for (int i = 0; i < iters; ++i) e.run(e.prog_common, 0);
    CUDA_OK(cudaEventRecord(t1, 0));
    CUDA_OK(cudaEventSynchronize(t1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, t0, t1));
    *ms_per_block = ms / iters / blocks;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
  });
}
