// Host engine behind the C ABI of include/ld_sampler.h: weight registry + repacking, the UNet
// "plan" (a fixed launch sequence over a liveness-managed workspace), the per-timestep sampler
// loop with CUDA-graph replay, and the extern "C" entry points.
//
// Reference behaviour restated here (no code shared): Unet.__init__/forward ddpm.py:286-451,
// ResUnet unet_model.py:91-137, p_sample_loop / p_sample / p_mean_variance / model_predictions
// ddpm.py:668-860, 929-977.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ld_sampler.h"
#include "ld_conv_tc.h"
#include "ld_kernels.h"
#include "ld_linattn_tc.h"
#include "ld_attn_tc.h"
#include "ld_conv7_tc.h"
#include "ld_launch.cuh"

namespace ld {

int& pdl_flag() {
  // measured in-process on the bench workload (tools/gpu_ab.py): every kernel (1): 5.21 ms per timestep without, 5.55 ms with -- the
  // early-scheduled CTAs of the next kernel start at different times on different SMs and skew the static tile split of the persistent
  // kernels; only the launches that follow a tiny kernel (2, ld_launch.cuh): 4.71 -> 4.61, 4.66 -> 4.67, 4.67 -> 4.68 ms in three
  // alternating runs (minima 1 % lower each time): inside the noise, so the default stays off.
  static int v = [] { const char* e = getenv("LD_PDL"); return e ? atoi(e) : 0; }();
  return v;
}

static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(LD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// -------------------------------------------------------------------------------------------------
// weight registry
// -------------------------------------------------------------------------------------------------
struct WSpec {
  std::string key;
  std::vector<int64_t> shape;
  std::vector<float> host;
  bool loaded = false;
  size_t numel() const { size_t n = 1; for (auto s : shape) n *= (size_t)s; return n; }
};

struct DevArr { float* p = nullptr; size_t n = 0; };

// a packed convolution ready for the kernels
struct ConvW {
  int Cin = 0, Cout = 0, ks = 1, stride = 1, pad = 0;
  float* w = nullptr;     // fp32 [taps][Cin][Cout]  (CUDA-core path, c1 and cout1 kernels)
  float* bias = nullptr;  // fp32 [Cout] or null
  ConvTcW tc;             // bf16 tcgen05 packing (empty in fp32 mode)
  ConvTcW tc_up2;         // nearest-x2 + 3x3 folded into one low-resolution convolution with pixel-shuffle output (conv_tc_pack_up2)
};

struct ResW { ConvW c1, c2, res; ConvTcW c1_dual; bool has_res = false, has_film = false; int film_off = 0; float *g1, *b1, *g2, *b2; int Cin, Cout; };
struct AttnW { bool full; int C; float* g; ConvW qkv; ConvW out; float* g2 = nullptr; LinAttnTcW la; };
struct CondW { ConvW a, b, id; float *ga, *ba, *gb, *bb, *gi, *bi; int Cin, Cmid, Cout; };

// -------------------------------------------------------------------------------------------------
// plan: fixed launch sequence
// -------------------------------------------------------------------------------------------------
struct Ten { void* p = nullptr; int N = 0, H = 0, W = 0, C = 0; int buf = -1; };

struct Plan {
  std::vector<std::function<int(cudaStream_t)>> ops;
  std::vector<void*> bufs; std::vector<size_t> cap; std::vector<char> busy;
  void* zero_arena = nullptr; size_t zero_bytes = 0, zero_used = 0;
  size_t total_bytes = 0;
  int N = 0, H = 0, W = 0;
  std::vector<std::pair<std::string, Ten>> tags;  // debug taps (only with option debug_keep)
  // external I/O of the plan
  const float* x = nullptr; const float* cond = nullptr; const void* cond_feat = nullptr; float* out = nullptr;
  ~Plan() {
    for (void* b : bufs) cudaFree(b);
    if (zero_arena) cudaFree(zero_arena);
  }
};

struct Staged {
  std::unique_ptr<Plan> cond, unet;
  float *x = nullptr, *c = nullptr, *o = nullptr; void* feat = nullptr; int64_t* t = nullptr;
  void free_all() { cudaFree(x); cudaFree(c); cudaFree(o); cudaFree(feat); cudaFree(t); x = c = o = nullptr; feat = nullptr; t = nullptr; }
};

struct Engine {
  ld_model_desc d{};
  int device = 0;
  bool bf = false;
  bool use_tc = false;
  std::vector<WSpec> specs;
  std::unordered_map<std::string, int> index;
  bool finalized = false;
  std::vector<void*> dev_allocs;
  std::map<std::string, float*> vec;  // small fp32 vectors (GN gamma/beta, RMSNorm g, biases)
  // model
  std::vector<int> dims;
  ConvW init_conv, final_conv;
  Conv7TcW init_tc;
  float *tw1, *tb1, *tw2, *tb2, *film_w, *film_b;
  int film_total = 0;
  std::vector<ResW> res;                  // all resnet blocks by name index
  std::map<std::string, int> res_index;
  std::map<std::string, AttnW> attn;
  std::vector<CondW> cond_blocks;
  std::map<std::string, ConvW> samp;      // down/up sampling convs
  int cond_C = 0, cond_div = 8;
  // schedule
  int T = 0;
  float *coef1 = nullptr, *coef2 = nullptr, *sigma = nullptr;
  float *obj_a = nullptr, *obj_b = nullptr;   // pred_noise / pred_v: x0 = a[t] x_t - b[t] out (null: pred_x0)
  // runtime
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  int64_t launches = 0;
  int64_t opt_use_graph = 1, opt_debug_keep = 0, opt_la_exact = 0, opt_attn_simt = 0, opt_async = 0, opt_up2 = 1;
  bool pending = false; ld_sample_desc pend_sd{}; bool pend_fuse = false;   // opt_async: checks deferred to ld_sample_finish
  unsigned int* la_flag = nullptr;        // soft-max underflow counter of the fused LinearAttention
  // sampler: FiLM rows of every timestep [film_tab_T][film_total] -- time MLP + block MLPs depend on t only (ddpm.py:339-344,191-194),
  // so the loop gathers a row instead of running them
  float* film_tab = nullptr; int film_tab_T = 0;
  std::map<std::string, std::unique_ptr<Plan>> plans;
  std::map<std::string, Staged> staged;   // stand-alone entry points (ld_unet_forward / ld_cond_encode)
  // sampler-owned state
  struct SampState {
    int B = 0, H = 0, W = 0, tloop = 0;
    float *xs = nullptr, *o = nullptr, *bm = nullptr, *cond = nullptr, *cond_out = nullptr, *cond_in = nullptr;
    void *feat_pair = nullptr, *feat_full = nullptr;
    unsigned int* counters = nullptr;
    int* t_dev = nullptr;
    int64_t* t64 = nullptr;
    std::vector<void*> allocs;
    void* ddim_blk = nullptr; size_t ddim_bytes = 0;   // DDIM: step index, times, coefficients
    cudaGraphExec_t g_branch = nullptr, g_single = nullptr;
    std::string key;
  } ss;

  ~Engine() {
    free_samp();
    plans.clear();
    for (auto& kv : staged) { kv.second.cond.reset(); kv.second.unet.reset(); kv.second.free_all(); }
    staged.clear();
    for (void* p : dev_allocs) cudaFree(p);
    for (auto& kv : attn) linattn_tc_free(&kv.second.la);
    conv7_tc_free(&init_tc);
    if (la_flag) cudaFree(la_flag);
    if (film_tab) cudaFree(film_tab);
    if (coef1) cudaFree(coef1);
    if (coef2) cudaFree(coef2);
    if (sigma) cudaFree(sigma);
    if (obj_a) cudaFree(obj_a);
    if (obj_b) cudaFree(obj_b);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    if (own_stream) cudaStreamDestroy(own_stream);
  }
  void free_samp() {
    if (ss.g_branch) cudaGraphExecDestroy(ss.g_branch);
    if (ss.g_single) cudaGraphExecDestroy(ss.g_single);
    for (void* p : ss.allocs) cudaFree(p);
    if (ss.ddim_blk) cudaFree(ss.ddim_blk);
    ss = SampState();
  }
  size_t esz() const { return bf ? 2 : 4; }
};

static void drop_plans(Engine& E) {
  E.free_samp();
  E.plans.clear();
  for (auto& kv : E.staged) { kv.second.cond.reset(); kv.second.unet.reset(); kv.second.free_all(); }
  E.staged.clear();
}

// -------------------------------------------------------------------------------------------------
// spec generation (state_dict keys + shapes)
// -------------------------------------------------------------------------------------------------
static void add_spec(Engine& E, const std::string& key, std::vector<int64_t> shape) {
  WSpec s; s.key = key; s.shape = std::move(shape);
  E.index[key] = (int)E.specs.size();
  E.specs.push_back(std::move(s));
}
static void spec_conv(Engine& E, const std::string& p, int co, int ci, int k, bool bias = true) {
  add_spec(E, p + ".weight", {co, ci, k, k});
  if (bias) add_spec(E, p + ".bias", {co});
}
static void spec_norm(Engine& E, const std::string& p, int c) { add_spec(E, p + ".weight", {c}); add_spec(E, p + ".bias", {c}); }
static void spec_res(Engine& E, const std::string& p, int ci, int co, int td) {
  add_spec(E, p + ".mlp.1.weight", {2 * co, td});
  add_spec(E, p + ".mlp.1.bias", {2 * co});
  spec_conv(E, p + ".block1.proj", co, ci, 3); spec_norm(E, p + ".block1.norm", co);
  spec_conv(E, p + ".block2.proj", co, co, 3); spec_norm(E, p + ".block2.norm", co);
  if (ci != co) spec_conv(E, p + ".res_conv", co, ci, 1);
}
static void spec_attn(Engine& E, const std::string& p, int c, bool full) {
  const int hid = E.d.attn_heads * E.d.attn_dim_head;
  add_spec(E, p + ".norm.g", {1, c, 1, 1});
  spec_conv(E, p + ".to_qkv", 3 * hid, c, 1, false);
  if (full) spec_conv(E, p + ".to_out", c, hid, 1);
  else { spec_conv(E, p + ".to_out.0", c, hid, 1); add_spec(E, p + ".to_out.1.g", {1, c, 1, 1}); }
}
static void spec_cond(Engine& E, const std::string& p, int ci, int cm, int co) {
  spec_conv(E, p + ".convblock.0", cm, ci, 3); spec_norm(E, p + ".convblock.1", cm);
  spec_conv(E, p + ".convblock.3", co, cm, 3); spec_norm(E, p + ".convblock.4", co);
  spec_conv(E, p + ".identity.0", co, ci, 3); spec_norm(E, p + ".identity.1", co);
}
static void build_specs(Engine& E) {
  const ld_model_desc& d = E.d;
  const int L = d.n_levels, td = 4 * d.dim;
  E.dims.clear(); E.dims.push_back(d.init_dim);
  for (int i = 0; i < L; ++i) E.dims.push_back(d.dim * d.dim_mults[i]);
  spec_cond(E, "cond_model.residual_conv1.0", 1, 32, 32);
  spec_cond(E, "cond_model.residual_conv2.0", 32, 32, 64);
  spec_cond(E, "cond_model.residual_conv3.0", 64, 64, 128);
  if (d.cond_mode == LD_COND_MRI) spec_cond(E, "cond_model.mid_conv.0", 128, 128, 256);
  spec_conv(E, "init_conv", d.init_dim, d.channels, 7);
  add_spec(E, "time_mlp.1.weight", {td, d.dim}); add_spec(E, "time_mlp.1.bias", {td});
  add_spec(E, "time_mlp.3.weight", {td, td}); add_spec(E, "time_mlp.3.bias", {td});
  char b[64];
  for (int i = 0; i < L; ++i) {
    const int di = E.dims[i], dn = E.dims[i + 1];
    snprintf(b, sizeof b, "downs.%d", i); std::string p = b;
    spec_res(E, p + ".0", di, di, td); spec_res(E, p + ".1", di, di, td);
    spec_attn(E, p + ".2", di, d.full_attn[i] != 0);
    if (i < L - 1) spec_conv(E, p + ".3.1", dn, 4 * di, 1); else spec_conv(E, p + ".3", dn, di, 3);
  }
  for (int i = 0; i < L; ++i) {
    const int di = E.dims[L - 1 - i], dn = E.dims[L - i];  // (dim_in, dim_out) reversed
    snprintf(b, sizeof b, "ups.%d", i); std::string p = b;
    spec_res(E, p + ".0", dn + di, dn, td); spec_res(E, p + ".1", dn + di, dn, td);
    spec_attn(E, p + ".2", dn, d.full_attn[L - 1 - i] != 0);
    if (i < L - 1) spec_conv(E, p + ".3.1", di, dn, 3); else spec_conv(E, p + ".3", di, dn, 3);
  }
  const int mid = E.dims[L];
  spec_res(E, "mid_block1", mid, mid, td);
  spec_attn(E, "mid_attn", mid, true);
  spec_res(E, "mid_block2", mid, mid, td);
  spec_res(E, "conv_fusion", 2 * mid, mid, td);
  spec_res(E, "final_res_block", 2 * d.dim, d.dim, td);
  spec_conv(E, "final_conv", d.channels, d.dim, 1);
}

// -------------------------------------------------------------------------------------------------
// weight upload / packing
// -------------------------------------------------------------------------------------------------
static const WSpec& W(const Engine& E, const std::string& key) { return E.specs[E.index.at(key)]; }
static bool has(const Engine& E, const std::string& key) { return E.index.count(key) != 0; }

static int upload(Engine& E, const std::vector<float>& h, float** out) {
  float* p = nullptr;
  CK(cudaMalloc(&p, h.size() * sizeof(float)));
  E.dev_allocs.push_back(p);
  CK(cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = p;
  return 0;
}
static int upload_key(Engine& E, const std::string& key, float** out) { return upload(E, W(E, key).host, out); }

// [Cout,Cin,k,k] -> [tap][Cin][Cout];  unshuffle=true: [Cout,4*Cs,1,1] -> 2x2/stride-2 taps (p1,p2) over Cs
static int pack_conv(Engine& E, const std::string& p, bool bias, bool unshuffle, ConvW* cw) {
  const WSpec& w = W(E, p + ".weight");
  const int co = (int)w.shape[0];
  int ci = (int)w.shape[1], k = (int)w.shape[2];
  std::vector<float> pk(w.host.size());
  if (unshuffle) {
    const int cs = ci / 4;
    for (int o = 0; o < co; ++o)
      for (int c = 0; c < cs; ++c)
        for (int q = 0; q < 4; ++q) pk[((size_t)q * cs + c) * co + o] = w.host[(size_t)o * ci + c * 4 + q];
    cw->Cin = cs; cw->ks = 2; cw->stride = 2; cw->pad = 0;
  } else {
    for (int o = 0; o < co; ++o)
      for (int c = 0; c < ci; ++c)
        for (int t = 0; t < k * k; ++t) pk[((size_t)t * ci + c) * co + o] = w.host[((size_t)o * ci + c) * k * k + t];
    cw->Cin = ci; cw->ks = k; cw->stride = 1; cw->pad = k / 2;
  }
  cw->Cout = co;
  int r = upload(E, pk, &cw->w);
  if (r) return r;
  cw->bias = nullptr;
  if (bias) { r = upload_key(E, p + ".bias", &cw->bias); if (r) return r; }
  if (E.use_tc && cw->Cin >= 32 && cw->Cout >= 32) {
    std::vector<float> bh;
    if (bias) bh = W(E, p + ".bias").host;
    // the 2x2/stride-2 (pixel-unshuffle) conv is a 1x1 conv over the (p1, p2, c) gather: same [tap][Cin][Cout] memory
    if (unshuffle) r = conv_tc_pack(pk.data(), bias ? bh.data() : nullptr, 4 * cw->Cin, cw->Cout, 1, 1, 0, &cw->tc);
    else r = conv_tc_pack(pk.data(), bias ? bh.data() : nullptr, cw->Cin, cw->Cout, cw->ks, cw->stride, cw->pad, &cw->tc);
    if (r) return fail(LD_ERR_CUDA, "conv_tc_pack(%s) failed", p.c_str());
  }
  return 0;
}
static int vecp(Engine& E, const std::string& key, float** out) {
  auto it = E.vec.find(key);
  if (it != E.vec.end()) { *out = it->second; return 0; }
  float* p; int r = upload_key(E, key, &p); if (r) return r;
  E.vec[key] = p; *out = p; return 0;
}

static int pack_res(Engine& E, const std::string& p, bool film, std::vector<float>& fw, std::vector<float>& fb) {
  ResW r;
  int rc;
  if ((rc = pack_conv(E, p + ".block1.proj", true, false, &r.c1))) return rc;
  if ((rc = pack_conv(E, p + ".block2.proj", true, false, &r.c2))) return rc;
  r.Cin = r.c1.Cin; r.Cout = r.c1.Cout;
  r.has_res = has(E, p + ".res_conv.weight");
  if (r.has_res && (rc = pack_conv(E, p + ".res_conv", true, false, &r.res))) return rc;
  if (r.has_res && E.use_tc && r.c1.tc.ready && r.c1.tc.ntile <= 64 && r.c1.Cout == r.c1.tc.ntile) {
    // block1.proj and res_conv read the same input (ddpm.py:207,212): one launch, two accumulators (ConvTcW::dual)
    const WSpec& w3 = W(E, p + ".block1.proj.weight"); const WSpec& w1 = W(E, p + ".res_conv.weight");
    const int co = r.c1.Cout, ci = r.c1.Cin;
    std::vector<float> p3((size_t)9 * ci * co), p1((size_t)ci * co);
    for (int o = 0; o < co; ++o)
      for (int c = 0; c < ci; ++c) {
        for (int t = 0; t < 9; ++t) p3[((size_t)t * ci + c) * co + o] = w3.host[((size_t)o * ci + c) * 9 + t];
        p1[(size_t)c * co + o] = w1.host[(size_t)o * ci + c];
      }
    if (conv_tc_pack(p3.data(), W(E, p + ".block1.proj.bias").host.data(), ci, co, 3, 1, 1, &r.c1_dual, p1.data(),
                     W(E, p + ".res_conv.bias").host.data()))
      return fail(LD_ERR_CUDA, "conv_tc_pack(dual %s) failed", p.c_str());
  }
  if ((rc = vecp(E, p + ".block1.norm.weight", &r.g1))) return rc;
  if ((rc = vecp(E, p + ".block1.norm.bias", &r.b1))) return rc;
  if ((rc = vecp(E, p + ".block2.norm.weight", &r.g2))) return rc;
  if ((rc = vecp(E, p + ".block2.norm.bias", &r.b2))) return rc;
  r.has_film = film;
  if (film) {
    r.film_off = (int)fb.size();
    const WSpec& w = W(E, p + ".mlp.1.weight"); const WSpec& b = W(E, p + ".mlp.1.bias");
    fw.insert(fw.end(), w.host.begin(), w.host.end());
    fb.insert(fb.end(), b.host.begin(), b.host.end());
  }
  E.res_index[p] = (int)E.res.size();
  E.res.push_back(r);
  return 0;
}
static int pack_attn(Engine& E, const std::string& p, int C, bool full) {
  AttnW a; a.full = full; a.C = C;
  int rc;
  if ((rc = vecp(E, p + ".norm.g", &a.g))) return rc;
  if ((rc = pack_conv(E, p + ".to_qkv", false, false, &a.qkv))) return rc;
  if (full) { if ((rc = pack_conv(E, p + ".to_out", true, false, &a.out))) return rc; }
  else {
    if ((rc = pack_conv(E, p + ".to_out.0", true, false, &a.out))) return rc;
    if ((rc = vecp(E, p + ".to_out.1.g", &a.g2))) return rc;
    if (E.use_tc && (E.d.attn_heads == 4 || E.d.attn_heads == 8) && E.d.attn_dim_head == 32 &&
        linattn_tc_pack(W(E, p + ".to_qkv.weight").host.data(), W(E, p + ".norm.g").host.data(), W(E, p + ".to_out.0.weight").host.data(),
                        W(E, p + ".to_out.0.bias").host.data(), W(E, p + ".to_out.1.g").host.data(), C, E.d.attn_heads, &a.la))
      return fail(LD_ERR_CUDA, "linattn_tc_pack(%s) failed", p.c_str());
  }
  E.attn[p] = a;
  return 0;
}
static int pack_cond(Engine& E, const std::string& p) {
  CondW c; int rc;
  if ((rc = pack_conv(E, p + ".convblock.0", true, false, &c.a))) return rc;
  if ((rc = pack_conv(E, p + ".convblock.3", true, false, &c.b))) return rc;
  if ((rc = pack_conv(E, p + ".identity.0", true, false, &c.id))) return rc;
  if ((rc = vecp(E, p + ".convblock.1.weight", &c.ga))) return rc;
  if ((rc = vecp(E, p + ".convblock.1.bias", &c.ba))) return rc;
  if ((rc = vecp(E, p + ".convblock.4.weight", &c.gb))) return rc;
  if ((rc = vecp(E, p + ".convblock.4.bias", &c.bb))) return rc;
  if ((rc = vecp(E, p + ".identity.1.weight", &c.gi))) return rc;
  if ((rc = vecp(E, p + ".identity.1.bias", &c.bi))) return rc;
  c.Cin = c.a.Cin; c.Cmid = c.a.Cout; c.Cout = c.b.Cout;
  E.cond_blocks.push_back(c);
  return 0;
}

static int finalize(Engine& E) {
  for (auto& s : E.specs)
    if (!s.loaded) return fail(LD_ERR_STATE, "weight '%s' was never loaded", s.key.c_str());
  const int L = E.d.n_levels;
  int rc;
  std::vector<float> fw, fb;
  if ((rc = pack_cond(E, "cond_model.residual_conv1.0"))) return rc;
  if ((rc = pack_cond(E, "cond_model.residual_conv2.0"))) return rc;
  if ((rc = pack_cond(E, "cond_model.residual_conv3.0"))) return rc;
  if (E.d.cond_mode == LD_COND_MRI) { if ((rc = pack_cond(E, "cond_model.mid_conv.0"))) return rc; E.cond_C = 256; E.cond_div = 8; }
  else { E.cond_C = 128; E.cond_div = 4; }
  if ((rc = pack_conv(E, "init_conv", true, false, &E.init_conv))) return rc;
  if (E.use_tc && E.d.channels == 1) {   // 7x7 init conv on tensor cores: weights [Cout][1][7][7] -> [tap][Cout]
    const WSpec& w = W(E, "init_conv.weight");
    const int co = (int)w.shape[0];
    std::vector<float> wt((size_t)49 * co);
    for (int o = 0; o < co; ++o)
      for (int t = 0; t < 49; ++t) wt[(size_t)t * co + o] = w.host[(size_t)o * 49 + t];
    if (conv7_tc_pack(wt.data(), W(E, "init_conv.bias").host.data(), co, &E.init_tc)) return fail(LD_ERR_CUDA, "conv7_tc_pack failed");
  }
  if ((rc = pack_conv(E, "final_conv", true, false, &E.final_conv))) return rc;
  if ((rc = upload_key(E, "time_mlp.1.weight", &E.tw1))) return rc;
  if ((rc = upload_key(E, "time_mlp.1.bias", &E.tb1))) return rc;
  if ((rc = upload_key(E, "time_mlp.3.weight", &E.tw2))) return rc;
  if ((rc = upload_key(E, "time_mlp.3.bias", &E.tb2))) return rc;
  char b[64];
  for (int i = 0; i < L; ++i) {
    snprintf(b, sizeof b, "downs.%d", i); std::string p = b;
    if ((rc = pack_res(E, p + ".0", true, fw, fb))) return rc;
    if ((rc = pack_res(E, p + ".1", true, fw, fb))) return rc;
    if ((rc = pack_attn(E, p + ".2", E.dims[i], E.d.full_attn[i] != 0))) return rc;
    ConvW cw;
    if (i < L - 1) rc = pack_conv(E, p + ".3.1", true, true, &cw); else rc = pack_conv(E, p + ".3", true, false, &cw);
    if (rc) return rc;
    E.samp[p + ".3"] = cw;
  }
  if ((rc = pack_res(E, "mid_block1", true, fw, fb))) return rc;
  if ((rc = pack_attn(E, "mid_attn", E.dims[L], true))) return rc;
  if ((rc = pack_res(E, "mid_block2", true, fw, fb))) return rc;
  if ((rc = pack_res(E, "conv_fusion", false, fw, fb))) return rc;  // ddpm.py:436: no time embedding
  for (int i = 0; i < L; ++i) {
    snprintf(b, sizeof b, "ups.%d", i); std::string p = b;
    if ((rc = pack_res(E, p + ".0", true, fw, fb))) return rc;
    if ((rc = pack_res(E, p + ".1", true, fw, fb))) return rc;
    if ((rc = pack_attn(E, p + ".2", E.dims[L - i], E.d.full_attn[L - 1 - i] != 0))) return rc;
    ConvW cw;
    if (i < L - 1) rc = pack_conv(E, p + ".3.1", true, false, &cw); else rc = pack_conv(E, p + ".3", true, false, &cw);
    if (rc) return rc;
    if (i < L - 1 && E.use_tc && cw.tc.ready) {   // Upsample (ddpm.py:114-118): fold the nearest x2 into the filter
      const WSpec& w = W(E, p + ".3.1.weight");
      const int co = (int)w.shape[0], ci = (int)w.shape[1];
      std::vector<float> pk((size_t)9 * ci * co);
      for (int o = 0; o < co; ++o)
        for (int c = 0; c < ci; ++c)
          for (int t = 0; t < 9; ++t) pk[((size_t)t * ci + c) * co + o] = w.host[((size_t)o * ci + c) * 9 + t];
      if (conv_tc_pack_up2(pk.data(), W(E, p + ".3.1.bias").host.data(), ci, co, &cw.tc_up2)) return fail(LD_ERR_CUDA, "conv_tc_pack_up2(%s) failed", p.c_str());
    }
    E.samp[p + ".3"] = cw;
  }
  if ((rc = pack_res(E, "final_res_block", true, fw, fb))) return rc;
  if (E.use_tc && attn_tc_configure()) return fail(LD_ERR_CUDA, "attn_tc_configure failed");
  CK(cudaMalloc(&E.la_flag, sizeof(unsigned int)));
  CK(cudaMemset(E.la_flag, 0, sizeof(unsigned int)));
  E.film_total = (int)fb.size();
  if ((rc = upload(E, fw, &E.film_w))) return rc;
  if ((rc = upload(E, fb, &E.film_b))) return rc;
  for (auto& s : E.specs) { std::vector<float>().swap(s.host); }
  E.finalized = true;
  return 0;
}

// -------------------------------------------------------------------------------------------------
// plan builder
// -------------------------------------------------------------------------------------------------
struct Builder {
  Engine& E; Plan& P; int err = 0;
  float* film = nullptr;  // [N][film_total] (row stride film_stride; 0 when the whole batch shares one timestep)
  int film_stride = 0;
  Builder(Engine& e, Plan& p) : E(e), P(p) {}

  Ten alloc(int N, int H, int W, int C, size_t elem) {
    const size_t bytes = (size_t)N * H * W * C * elem;
    int best = -1;
    for (size_t i = 0; i < P.bufs.size(); ++i)
      if (!P.busy[i] && P.cap[i] >= bytes && (best < 0 || P.cap[i] < P.cap[best])) best = (int)i;
    if (best < 0) {
      void* p = nullptr;
      if (cudaMalloc(&p, bytes) != cudaSuccess) { err = fail(LD_ERR_CUDA, "workspace cudaMalloc(%zu) failed", bytes); return Ten(); }
      P.bufs.push_back(p); P.cap.push_back(bytes); P.busy.push_back(0); P.total_bytes += bytes;
      best = (int)P.bufs.size() - 1;
    }
    P.busy[best] = 1;
    Ten t; t.p = P.bufs[best]; t.N = N; t.H = H; t.W = W; t.C = C; t.buf = best;
    return t;
  }
  Ten act(int N, int H, int W, int C) { return alloc(N, H, W, C, E.esz()); }
  void release(Ten& t) { if (t.buf >= 0 && !E.opt_debug_keep) P.busy[t.buf] = 0; t.buf = -1; }
  void tag(const std::string& name, const Ten& t) { if (E.opt_debug_keep) P.tags.emplace_back(name, t); }
  // zero-initialised scratch (GN statistics, linear-attention accumulators): one memset per forward
  void* zalloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (P.zero_used == 0) P.zero_used = 256;  // offset 0 is reserved: a null offset means "absent"
    size_t off = P.zero_used; P.zero_used += bytes;
    return (void*)off;  // resolved after the arena is allocated
  }
  template <typename F> void op(F f) { P.ops.emplace_back(std::move(f)); }

  // ---- convolution dispatch -----------------------------------------------------------------
  // GroupNorm (+FiLM) + activation of the SOURCE tensor, applied while the conv stages its input
  struct ProSpec { double* st; const float* g; const float* b; int G; int film_off; int act; };
  // conv with optional fused prologue (normalise-on-load) and fused GroupNorm statistics of the output.
  // `stats_off` / `pro->st` are zero-arena offsets.  Whatever the selected kernel cannot fuse is run as a
  // separate pass, so callers are uniform across the fp32 (CUDA-core) and bf16 (tcgen05) paths.
  Ten conv(const ConvW& cw, const Ten& a_in, const Ten* b, bool up, const Ten* resid, int outH, int outW,
           const ProSpec* pro = nullptr, double* stats_off = nullptr, int stats_G = 0) {
    const bool bf = E.bf;
    Plan* pp = &P;
    Ten a = a_in;
    bool own_a = false;
    ConvTcArgs ta;
    bool use_tc = false;
    static int no_up2 = -1;   // env LD_CONV_NO_UP2=1: keep the replicate-on-load up-sampling path (A/B aid)
    if (no_up2 < 0) { const char* e = getenv("LD_CONV_NO_UP2"); no_up2 = e ? atoi(e) : 0; }
    if (E.use_tc && up && !b && !resid && !pro && !stats_off && cw.tc_up2.ready && !no_up2 && E.opt_up2) {
      ConvTcArgs tu;
      tu.src0 = a.p; tu.C0 = a.C; tu.N = a.N; tu.H = a.H; tu.W = a.W; tu.Hin = a.H; tu.Win = a.W; tu.ps = cw.Cout;
      if (outH == 2 * a.H && outW == 2 * a.W && conv_tc_supports(cw.tc_up2, tu)) {
        Ten o = act(a.N, outH, outW, cw.Cout);
        if (err) return o;
        tu.dst = o.p;
        const ConvTcW* w = &cw.tc_up2;
        op([tu, w](cudaStream_t s) { return conv_tc_launch(*w, tu, s); });
        return o;
      }
    }
    if (E.use_tc && cw.tc.ready) {
      ta.src0 = a.p; ta.C0 = a.C; ta.src1 = b ? b->p : nullptr; ta.C1 = b ? b->C : 0;
      ta.N = a.N; ta.H = outH; ta.W = outW; ta.Hin = a.H; ta.Win = a.W; ta.up = up ? 1 : 0;
      ta.res = resid ? resid->p : nullptr;
      ta.ds = cw.ks == 2 ? 1 : 0;
      use_tc = conv_tc_supports(cw.tc, ta);
    }
    bool pro_fused = false, stats_fused = false;
    if (use_tc && pro) {
      ConvTcArgs t2 = ta;
      t2.pro_ab = (const float*)8; t2.pro_act = pro->act;
      pro_fused = conv_tc_supports(cw.tc, t2);
    }
    if (use_tc && stats_off) {
      ConvTcArgs t2 = ta;
      t2.stats = (double*)8; t2.stats_G = stats_G;
      stats_fused = conv_tc_supports(cw.tc, t2);
    }
    if (pro && !pro_fused) {  // materialise the normalised activation
      a = gn_apply(a_in, pro->st, pro->g, pro->b, pro->G, pro->film_off, pro->act, nullptr, 0);
      own_a = true;
      ta.src0 = a.p;
    }
    Ten o = act(a.N, outH, outW, cw.Cout);
    if (err) return o;
    if (use_tc) {
      ta.dst = o.p;
      const ConvTcW* w = &cw.tc;
      const size_t pro_off = pro_fused ? (size_t)pro->st : 0, st_off = stats_fused ? (size_t)stats_off : 0;
      Ten abt;
      if (pro_fused) {   // y = a x + b table of the source tensor's GroupNorm (+FiLM), one tiny kernel
        abt = alloc(a.N, 1, 1, 2 * a.C, 4);
        const float* g = pro->g; const float* bb = pro->b; const int G = pro->G, C = a.C, N = a.N; const long long HW = (long long)a.H * a.W;
        const float* fl = pro->film_off >= 0 ? film + pro->film_off : nullptr; const int fs = film_stride;
        float* ab = (float*)abt.p;
        op([pp, pro_off, g, bb, fl, fs, G, C, N, HW, ab](cudaStream_t s) {
          return gn_coef_launch((const double*)((char*)pp->zero_arena + pro_off), g, bb, fl, fs, G, C, N, HW, 1e-5f, ab, s);
        });
        ta.pro_ab = ab; ta.pro_act = pro->act;
      }
      if (stats_fused) ta.stats_G = stats_G;
      op([ta, w, pp, stats_fused, st_off](cudaStream_t s) {
        ConvTcArgs q = ta;
        if (stats_fused) q.stats = (double*)((char*)pp->zero_arena + st_off);
        return conv_tc_launch(*w, q, s);
      });
      if (pro_fused) release(abt);
    } else {
      ConvP p{};
      p.src0 = a.p; p.C0 = a.C; p.src1 = b ? b->p : nullptr; p.C1 = b ? b->C : 0;
      p.N = a.N; p.H = outH; p.W = outW; p.Hin = a.H; p.Win = a.W;
      p.ks = cw.ks; p.stride = cw.stride; p.pad = cw.pad; p.up = up ? 1 : 0;
      p.w = cw.w; p.bias = cw.bias; p.Cout = cw.Cout; p.dst = o.p; p.res = resid ? resid->p : nullptr;
      p.M = (long long)a.N * outH * outW;
      op([p, bf](cudaStream_t s) { return launch_conv_simt(p, bf, s); });
    }
    if (own_a) release(a);
    if (stats_off && !stats_fused) stats_into(o, stats_G, stats_off);
    return o;
  }
  // block1.proj (3x3, GroupNorm statistics of its output) and res_conv (1x1) of the same input in one launch
  bool conv_dual(const ConvTcW& w, const Ten& a, const Ten* b, double* stats_off, int stats_G, Ten* o1, Ten* o2) {
    if (!E.use_tc || !w.ready || !w.dual) return false;
    ConvTcArgs ta;
    ta.src0 = a.p; ta.C0 = a.C; ta.src1 = b ? b->p : nullptr; ta.C1 = b ? b->C : 0;
    ta.N = a.N; ta.H = a.H; ta.W = a.W; ta.Hin = a.H; ta.Win = a.W;
    ta.stats = (double*)8; ta.stats_G = stats_G; ta.dst2 = (void*)8;
    if (!conv_tc_supports(w, ta)) return false;
    *o1 = act(a.N, a.H, a.W, w.Cout);
    *o2 = act(a.N, a.H, a.W, w.Cout);
    if (err) return true;
    ta.dst = o1->p; ta.dst2 = o2->p; ta.stats = nullptr;
    const ConvTcW* wp = &w; Plan* pp = &P; const size_t st_off = (size_t)stats_off;
    op([ta, wp, pp, st_off](cudaStream_t s) {
      ConvTcArgs q = ta;
      q.stats = (double*)((char*)pp->zero_arena + st_off);
      return conv_tc_launch(*wp, q, s);
    });
    return true;
  }
  Ten conv_same(const ConvW& cw, const Ten& a, const Ten* b = nullptr, const Ten* resid = nullptr,
                const ProSpec* pro = nullptr, double* stats_off = nullptr, int stats_G = 0) {
    return conv(cw, a, b, false, resid, a.H, a.W, pro, stats_off, stats_G);
  }
  double* stats_alloc(int N, int G) { return (double*)zalloc((size_t)N * G * 2 * sizeof(double)); }
  void stats_into(const Ten& x, int G, double* st) {
    Plan* pp = &P; const bool bf = E.bf; Ten xx = x;
    op([pp, st, xx, G, bf](cudaStream_t s) {
      return launch_gn_stats(xx.p, (double*)((char*)pp->zero_arena + (size_t)st), xx.N, xx.H * xx.W, xx.C, G, bf, s);
    });
  }
  // out = act(GN(xa)*film) (+ xb variants), see GnApplyP
  Ten gn_apply(const Ten& xa, double* stA, const float* gA, const float* bA, int GA, int film_off, int act,
               const Ten* xb, int modeB, double* stB = nullptr, const float* gB = nullptr, const float* bB = nullptr,
               int GB = 1, const ConvW* dot = nullptr, float* dot_out = nullptr) {
    Ten o;
    if (!dot) o = act_t(xa);
    GnApplyP p{};
    if (dot) { p.dot_w = dot->w; p.dot_b = dot->bias; p.dot_out = dot_out; }
    p.xa = xa.p; p.statsA = stA; p.gA = gA; p.bA = bA; p.GA = GA;
    p.xb = xb ? xb->p : nullptr; p.statsB = stB; p.gB = gB; p.bB = bB; p.GB = GB; p.modeB = modeB;
    p.film = film_off >= 0 ? film + film_off : nullptr; p.film_stride = film_stride;
    p.act = act; p.out = o.p; p.N = xa.N; p.HW = xa.H * xa.W; p.C = xa.C; p.eps = 1e-5f;
    Plan* pp = &P; const bool bf = E.bf;
    op([pp, p, bf](cudaStream_t s) {
      GnApplyP q = p;
      q.statsA = (const double*)((char*)pp->zero_arena + (size_t)p.statsA);
      if (p.statsB) q.statsB = (const double*)((char*)pp->zero_arena + (size_t)p.statsB);
      return launch_gn_apply(q, bf, s);
    });
    return o;
  }
  Ten act_t(const Ten& like) { return act(like.N, like.H, like.W, like.C); }

  // ---- ResnetBlock (ddpm.py:200-212) -----------------------------------------------------------
  // block1: conv (+ statistics in its epilogue); block2: conv whose staging applies GN1 + FiLM + SiLU and whose
  // epilogue gathers the GN2 statistics; one elementwise pass: SiLU(GN2(h2)) + res_conv(x).
  // `dot` / `dot_out`: fold the 1x1 convolution to one fp32 channel that consumes the block's output (final_conv, ddpm.py:398)
  // into the output pass; the block's own output tensor is then never written
  Ten resblock(const std::string& name, Ten& a, Ten* b, const ConvW* dot = nullptr, float* dot_out = nullptr) {
    const ResW& r = E.res[E.res_index.at(name)];
    const int G = E.d.resnet_groups;
    double* s1 = stats_alloc(a.N, G);
    Ten h1, rs;
    const bool dual = r.has_res && conv_dual(r.c1_dual, a, b, s1, G, &h1, &rs);
    if (!dual) h1 = conv_same(r.c1, a, b, nullptr, nullptr, s1, G);
    ProSpec pr{s1, r.g1, r.b1, G, r.has_film ? r.film_off : -1, 1};
    double* s2 = stats_alloc(a.N, G);
    Ten h2 = conv_same(r.c2, h1, nullptr, nullptr, &pr, s2, G);
    release(h1);
    Ten o;
    if (r.has_res) {
      if (!dual) rs = conv_same(r.res, a, b);
      o = gn_apply(h2, s2, r.g2, r.b2, G, -1, 1, &rs, 1, nullptr, nullptr, nullptr, 1, dot, dot_out);
      release(rs);
    } else {
      o = gn_apply(h2, s2, r.g2, r.b2, G, -1, 1, &a, 1, nullptr, nullptr, nullptr, 1, dot, dot_out);
    }
    release(h2);
    return o;
  }
  // ---- attention + residual (ddpm.py:425,431,444) ------------------------------------------------
  Ten attention(const std::string& name, Ten& x) {
    const AttnW& a = E.attn.at(name);
    const bool bf = E.bf;
    const int heads = E.d.attn_heads, hid = heads * 32;
    if (!a.full && a.la.ready && !E.opt_la_exact) {
      // fused tcgen05 LinearAttention: x is read twice, the result written once (ld_linattn_tc.cu)
      Ten out = act_t(x);
      Ten mn = alloc(x.N, 1, 1, hid * x.C, 2);
      LinAttnTcArgs la;
      la.x = x.p; la.out = out.p; la.N = x.N; la.HW = x.H * x.W; la.Mn = mn.p; la.flag = E.la_flag;
      const size_t ctx_off = (size_t)zalloc((size_t)x.N * hid * x.C * 4), ks_off = (size_t)zalloc((size_t)x.N * hid * 4);
      const LinAttnTcW* w = &a.la;
      Plan* pp = &P;
      op([pp, la, w, ctx_off, ks_off](cudaStream_t s) {
        LinAttnTcArgs q = la;
        q.Z = (float*)((char*)pp->zero_arena + ctx_off);
        q.ksum = (float*)((char*)pp->zero_arena + ks_off);
        return linattn_tc_launch(*w, q, s);
      });
      release(mn);
      return out;
    }
    Ten xn = act_t(x);
    {
      Ten xx = x, o = xn; const float* g = a.g;
      op([xx, o, g, bf](cudaStream_t s) { return launch_rmsnorm(xx.p, g, nullptr, o.p, (long long)xx.N * xx.H * xx.W, xx.C, bf, s); });
    }
    Ten qkv = conv_same(a.qkv, xn);
    release(xn);
    Ten out;
    if (a.full) {
      Ten ao = act(x.N, x.H, x.W, hid);
      if (E.use_tc && !E.opt_attn_simt) {
        Ten sc = alloc(1, 1, 1, (int)((attn_tc_scratch_bytes(x.N, x.H * x.W, heads) + 255) / 256), 256);
        Ten q = qkv, o = ao;
        op([q, o, sc, heads](cudaStream_t s) { return attn_tc_launch(q.p, o.p, sc.p, q.N, q.H * q.W, heads, s); });
        release(sc);
      } else {
        Ten q = qkv, o = ao;
        op([q, o, heads, bf](cudaStream_t s) { return launch_attention_simt(q.p, o.p, q.N, q.H * q.W, heads, bf, s); });
      }
      out = conv_same(a.out, ao, nullptr, &x);
      release(ao);
    } else {
      out = act_t(x);
      LinAttnP p{};
      p.qkv = qkv.p; p.N = x.N; p.HW = x.H * x.W; p.heads = heads; p.C = x.C;
      p.chunks = std::max(1, std::min(64, p.HW / 1024));
      Ten t_part = alloc(x.N, 1, 1, p.chunks * hid, 4), t_max = alloc(x.N, 1, 1, hid, 4), t_mn = alloc(x.N, 1, 1, hid * x.C, 4);
      p.kmax_part = (float*)t_part.p; p.kmax = (float*)t_max.p; p.Mn = (float*)t_mn.p;
      p.ctx = (float*)zalloc((size_t)x.N * hid * 32 * 4); p.ksum = (float*)zalloc((size_t)x.N * hid * 4);
      p.wout = a.out.w; p.bout = a.out.bias; p.g2 = a.g2; p.x = x.p; p.out = out.p;
      Plan* pp = &P;
      op([pp, p, bf](cudaStream_t s) {
        LinAttnP q = p;
        q.ctx = (float*)((char*)pp->zero_arena + (size_t)p.ctx);
        q.ksum = (float*)((char*)pp->zero_arena + (size_t)p.ksum);
        return launch_linear_attention(q, bf, s);
      });
      release(t_part); release(t_max); release(t_mn);
    }
    release(qkv);
    return out;
  }
  // ---- conditional encoder block (unet_model.py:37-51) -------------------------------------------
  Ten cond_block(const CondW& c, const Ten* xin, const float* img, int N, int H, int W) {
    const bool bf = E.bf;
    Ten a, id;
    double *sa = stats_alloc(N, 16), *si = stats_alloc(N, 16), *sb = stats_alloc(N, 16);
    if (c.Cin == 1) {
      a = act(N, H, W, c.Cmid); id = act(N, H, W, c.Cout);
      const ConvW *ca = &c.a, *ci = &c.id; Ten aa = a, ii = id;
      op([=](cudaStream_t s) { return launch_conv_c1(img, ca->w, ca->bias, aa.p, N, H, W, ca->Cout, 3, bf, s); });
      op([=](cudaStream_t s) { return launch_conv_c1(img, ci->w, ci->bias, ii.p, N, H, W, ci->Cout, 3, bf, s); });
      stats_into(a, 16, sa); stats_into(id, 16, si);
    } else {
      a = conv_same(c.a, *xin, nullptr, nullptr, nullptr, sa, 16);
      id = conv_same(c.id, *xin, nullptr, nullptr, nullptr, si, 16);
    }
    ProSpec pr{sa, c.ga, c.ba, 16, -1, 2};
    Ten b2 = conv_same(c.b, a, nullptr, nullptr, &pr, sb, 16);
    release(a);
    Ten o = gn_apply(b2, sb, c.gb, c.bb, 16, -1, 2, &id, 2, si, c.gi, c.bi, 16);
    release(b2); release(id);
    return o;
  }
  Ten maxpool(Ten& x) {
    Ten o = act(x.N, x.H / 2, x.W / 2, x.C);
    Ten xx = x, oo = o; const bool bf = E.bf;
    op([xx, oo, bf](cudaStream_t s) { return launch_maxpool2(xx.p, oo.p, xx.N, xx.H, xx.W, xx.C, bf, s); });
    release(x);
    return o;
  }
  int finish() {
    if (err) return err;
    if (P.zero_used) {
      if (cudaMalloc(&P.zero_arena, P.zero_used) != cudaSuccess) return fail(LD_ERR_CUDA, "zero arena cudaMalloc failed");
      P.zero_bytes = P.zero_used; P.total_bytes += P.zero_used;
    }
    return 0;
  }
};

// cond encoder plan: cond fp32 [N,H,W] -> feat T [N,H/f,W/f,Cf]  (unet_model.py:122-137)
static int build_cond_plan(Engine& E, Plan& P, int N, int H, int W, const float* cond, void* feat_out) {
  Builder B(E, P);
  P.N = N; P.H = H; P.W = W;
  Ten x = B.cond_block(E.cond_blocks[0], nullptr, cond, N, H, W);
  x = B.maxpool(x);
  Ten y = B.cond_block(E.cond_blocks[1], &x, nullptr, N, x.H, x.W); B.release(x);
  y = B.maxpool(y);
  Ten z = B.cond_block(E.cond_blocks[2], &y, nullptr, N, y.H, y.W); B.release(y);
  if (E.d.cond_mode == LD_COND_MRI) {
    z = B.maxpool(z);
    Ten w = B.cond_block(E.cond_blocks[3], &z, nullptr, N, z.H, z.W); B.release(z);
    z = w;
  }
  if (B.err) return B.err;
  const size_t bytes = (size_t)N * z.H * z.W * z.C * E.esz();
  void* src = z.p;
  B.op([src, feat_out, bytes](cudaStream_t s) {
    return cudaMemcpyAsync(feat_out, src, bytes, cudaMemcpyDeviceToDevice, s) == cudaSuccess ? 0 : -1;
  });
  return B.finish();
}

// UNet plan (ddpm.py:404-451) with the conditional features supplied (hoisted out of the loop)
static int build_unet_plan(Engine& E, Plan& P, int N, int H, int W, const float* x, const void* cond_feat,
                           const int64_t* t64, const int* t_scalar, float* out) {
  Builder B(E, P);
  const bool bf = E.bf;
  const int L = E.d.n_levels;
  P.N = N; P.H = H; P.W = W;
  // time embedding + FiLM vectors
  Ten st = B.alloc(N, 1, 1, 4 * E.d.dim, 4), fl = B.alloc(N, 1, 1, E.film_total, 4);
  B.film = (float*)fl.p;
  B.film_stride = t_scalar ? 0 : E.film_total;
  {
    TimeP tp{};
    tp.t = t64; tp.t_scalar = t_scalar; tp.N = t_scalar ? 1 : N; tp.dim = E.d.dim; tp.theta = E.d.sinusoidal_theta;
    tp.neg_step = (float)(-(std::log((double)E.d.sinusoidal_theta) / (double)(E.d.dim / 2 - 1)));
    tp.w1 = E.tw1; tp.b1 = E.tb1; tp.w2 = E.tw2; tp.b2 = E.tb2; tp.st = (float*)st.p;
    tp.wf = E.film_w; tp.bf_ = E.film_b; tp.total = E.film_total; tp.film = (float*)fl.p;
    Engine* Ep = &E;
    B.op([tp, Ep](cudaStream_t s) {
      // sampler loop: one shared timestep whose row was precomputed (ensure_film_table); else run the two MLPs
      if (tp.t_scalar && Ep->film_tab && (Ep->film_total & 3) == 0) return launch_film_gather(Ep->film_tab, tp.total, tp.t_scalar, tp.film, s);
      return launch_time_film(tp, s);
    });
  }
  Ten h = B.act(N, H, W, E.d.init_dim);
  {
    const ConvW* c = &E.init_conv; Ten hh = h;
    const Conv7TcW* c7 = &E.init_tc;
    if (E.use_tc && c7->ready) B.op([=](cudaStream_t s) { return conv7_tc_launch(*c7, x, hh.p, N, H, W, s); });
    else B.op([=](cudaStream_t s) { return launch_conv_c1(x, c->w, c->bias, hh.p, N, H, W, c->Cout, 7, bf, s); });
  }
  B.tag("init_conv", h);
  Ten r = h;  // ddpm.py:414 (clone not needed: h is never written again)
  std::vector<Ten> skips;
  Ten cur = h;
  bool cur_is_r = true;
  char nb[64];
  for (int i = 0; i < L; ++i) {
    snprintf(nb, sizeof nb, "downs.%d", i); std::string p = nb;
    Ten a = B.resblock(p + ".0", cur, nullptr);
    if (!cur_is_r) B.release(cur);
    cur_is_r = false;
    skips.push_back(a);
    B.tag(p + ".0", a);
    Ten b = B.resblock(p + ".1", a, nullptr);
    B.tag(p + ".1", b);
    Ten c = B.attention(p + ".2", b);
    B.tag(p + ".2", c);
    B.release(b);
    skips.push_back(c);
    const ConvW& dw = E.samp.at(p + ".3");
    if (i < L - 1) cur = B.conv(dw, c, nullptr, false, nullptr, c.H / 2, c.W / 2);
    else cur = B.conv_same(dw, c);
    B.tag(p + ".3", cur);
  }
  {
    Ten a = B.resblock("mid_block1", cur, nullptr); B.release(cur);
    B.tag("mid_block1", a);
    Ten b = B.attention("mid_attn", a); B.release(a);
    B.tag("mid_attn", b);
    Ten c = B.resblock("mid_block2", b, nullptr); B.release(b);
    B.tag("mid_block2", c);
    Ten cf; cf.p = const_cast<void*>(cond_feat); cf.N = N; cf.H = c.H; cf.W = c.W; cf.C = E.cond_C;
    cur = B.resblock("conv_fusion", c, &cf); B.release(c);
    B.tag("conv_fusion", cur);
  }
  for (int i = 0; i < L; ++i) {
    snprintf(nb, sizeof nb, "ups.%d", i); std::string p = nb;
    Ten s1 = skips.back(); skips.pop_back();
    Ten a = B.resblock(p + ".0", cur, &s1); B.release(cur); B.release(s1);
    B.tag(p + ".0", a);
    Ten s2 = skips.back(); skips.pop_back();
    Ten b = B.resblock(p + ".1", a, &s2); B.release(a); B.release(s2);
    B.tag(p + ".1", b);
    Ten c = B.attention(p + ".2", b); B.release(b);
    B.tag(p + ".2", c);
    const ConvW& uw = E.samp.at(p + ".3");
    if (i < L - 1) cur = B.conv(uw, c, nullptr, true, nullptr, c.H * 2, c.W * 2);
    else cur = B.conv_same(uw, c);
    B.tag(p + ".3", cur);
    B.release(c);
  }
  // final ResnetBlock + final_conv (ddpm.py:449-451): on the bf16 path the 1x1 convolution to the single output channel rides on the
  // block's output pass (the 32-channel tensor is never written); debug taps and the fp32 parity path keep the two steps apart
  if (gn_apply_can_dot(E.d.dim, bf) && !E.opt_debug_keep) {
    B.resblock("final_res_block", cur, &r, &E.final_conv, out); B.release(cur); B.release(r);
    if (B.err) return B.err;
    return B.finish();
  }
  Ten f = B.resblock("final_res_block", cur, &r); B.release(cur); B.release(r);
  B.tag("final_res_block", f);
  if (B.err) return B.err;
  {
    const ConvW* c = &E.final_conv; Ten ff = f;
    B.op([=](cudaStream_t s) { return launch_conv_cout1(ff.p, c->w, c->bias, out, (long long)N * H * W, ff.C, bf, s); });
  }
  B.release(f);
  return B.finish();
}

static int run_plan(Engine& E, Plan& P, cudaStream_t s) {
  if (P.zero_arena && cudaMemsetAsync(P.zero_arena, 0, P.zero_bytes, s) != cudaSuccess)
    return fail(LD_ERR_CUDA, "memset of the zero arena failed");
  // development aid (env LD_PROFILE_OPS=1): per-op device time of this plan, CUDA events on the launch stream
  static int prof = -1;
  if (prof < 0) { const char* e = getenv("LD_PROFILE_OPS"); prof = e ? atoi(e) : 0; }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (prof) cudaStreamIsCapturing(s, &cap);
  const bool timed = prof && cap == cudaStreamCaptureStatusNone && P.ops.size() >= (size_t)prof;
  std::vector<cudaEvent_t> evs;
  std::vector<int> nl;
  if (timed) { evs.resize(P.ops.size() + 1); for (auto& ev : evs) cudaEventCreate(&ev); cudaEventRecord(evs[0], s); }
  size_t oi = 0;
  for (auto& f : P.ops) {
    int n = f(s);
    if (n < 0) return fail(LD_ERR_CUDA, "kernel launch failed in plan");
    E.launches += n;
    if (timed) { cudaEventRecord(evs[++oi], s); nl.push_back(n); }
  }
  if (timed) {
    cudaStreamSynchronize(s);
    float tot = 0;
    for (size_t i = 0; i < P.ops.size(); ++i) {
      float ms = 0; cudaEventElapsedTime(&ms, evs[i], evs[i + 1]); tot += ms;
      fprintf(stderr, "LDPROF %3zu n=%d %9.1f us\n", i, nl[i], ms * 1000.f);
    }
    fprintf(stderr, "LDPROF total %9.1f us over %zu ops (N=%d %dx%d)\n", tot * 1000.f, P.ops.size(), P.N, P.H, P.W);
    for (auto& ev : evs) cudaEventDestroy(ev);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "launch error: %s", cudaGetErrorString(e));
  return 0;
}

static int check_shape(const Engine& E, int H, int W) {
  const int f = 1 << (E.d.n_levels - 1);
  if (H % f || W % f)  // ddpm.py:405
    return fail(LD_ERR_INVALID, "your input dimensions (%d, %d) need to be divisible by %d, given the unet", H, W, f);
  if (H % E.cond_div || W % E.cond_div) return fail(LD_ERR_INVALID, "input dimensions must be divisible by %d for the conditional encoder", E.cond_div);
  if ((H >> (E.d.n_levels - 1)) != H / E.cond_div)
    return fail(LD_ERR_INVALID, "UNet bottleneck (S/%d) and conditional encoder (S/%d) resolutions differ", f, E.cond_div);
  return 0;
}

// device scratch that is released on every exit path
struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
  template <typename T> cudaError_t get(T** p, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e == cudaSuccess) { ptrs.push_back(q); *p = (T*)q; }
    return e;
  }
};

}  // namespace ld

// =================================================================================================
// C ABI
// =================================================================================================
using namespace ld;
struct ld_handle { Engine E; };

extern "C" {

const char* ld_last_error(void) { return g_err; }
const char* ld_version(void) { return "ld_sampler 0.1 sm_100a"; }

int ld_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

int ld_create(const ld_model_desc* d, int device, ld_handle** out) {
  if (!d || !out) return fail(LD_ERR_INVALID, "null argument");
  if (d->n_levels < 1 || d->n_levels > LD_MAX_LEVELS) return fail(LD_ERR_INVALID, "n_levels out of range");
  if (d->channels != 1) return fail(LD_ERR_INVALID, "only channels == 1 is supported on this path");
  if (d->attn_dim_head != 32) return fail(LD_ERR_INVALID, "attn_dim_head must be 32");
  if (d->dim % 32 || d->init_dim % 32) return fail(LD_ERR_INVALID, "dim and init_dim must be multiples of 32");
  if (d->dim * d->dim_mults[d->n_levels - 1] != (d->cond_mode == LD_COND_MRI ? 256 : 128))
    return fail(LD_ERR_INVALID, "dim*dim_mults[-1] must equal the conditional encoder width (ddpm.py:380, unet_model.py:100)");
  if ((1 << (d->n_levels - 1)) != (d->cond_mode == LD_COND_MRI ? 8 : 4))
    return fail(LD_ERR_INVALID, "number of levels does not match the conditional encoder depth");
  ld_handle* h = new ld_handle();
  h->E.d = *d; h->E.device = device;
  h->E.bf = d->precision == LD_PREC_BF16;
  h->E.use_tc = h->E.bf;
  if (const char* e = getenv("LD_USE_GRAPH")) h->E.opt_use_graph = atoi(e);   // development aid: eager launches (A/B against graph replay)
  build_specs(h->E);
  *out = h;
  return 0;
}

int ld_destroy(ld_handle* h) {
  if (!h) return 0;
  if (ld_device_count() > 0) cudaSetDevice(h->E.device);
  delete h;
  return 0;
}

int ld_num_weights(const ld_handle* h) { return h ? (int)h->E.specs.size() : 0; }
int ld_weight_info(const ld_handle* h, int i, const char** key, int64_t shape[4], int* ndim) {
  if (!h || i < 0 || i >= (int)h->E.specs.size()) return fail(LD_ERR_INVALID, "weight index out of range");
  const WSpec& s = h->E.specs[i];
  if (key) *key = s.key.c_str();
  if (ndim) *ndim = (int)s.shape.size();
  if (shape) for (size_t k = 0; k < s.shape.size() && k < 4; ++k) shape[k] = s.shape[k];
  return 0;
}
int ld_load_weight(ld_handle* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  if (!h || !key || !data) return fail(LD_ERR_INVALID, "null argument");
  Engine& E = h->E;
  if (E.finalized) return fail(LD_ERR_STATE, "weights already finalized");
  auto it = E.index.find(key);
  if (it == E.index.end()) return fail(LD_ERR_KEY, "unexpected key '%s'", key);
  WSpec& s = E.specs[it->second];
  if (s.loaded) return fail(LD_ERR_KEY, "key '%s' loaded twice", key);
  if (ndim != (int)s.shape.size()) return fail(LD_ERR_KEY, "size mismatch for '%s' (ndim %d vs %zu)", key, ndim, s.shape.size());
  for (int i = 0; i < ndim; ++i)
    if (shape[i] != s.shape[i]) return fail(LD_ERR_KEY, "size mismatch for '%s' (dim %d: %lld vs %lld)", key, i, (long long)shape[i], (long long)s.shape[i]);
  s.host.assign(data, data + s.numel());
  s.loaded = true;
  return 0;
}
int ld_finalize_weights(ld_handle* h) {
  if (!h) return fail(LD_ERR_INVALID, "null handle");
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  Engine& E = h->E;
  if (E.finalized) return fail(LD_ERR_STATE, "weights already finalized");
  CK(cudaSetDevice(E.device));
  if (!E.own_stream) {
    CK(cudaStreamCreateWithFlags(&E.own_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&E.ev_in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&E.ev_out, cudaEventDisableTiming));
  }
  return finalize(E);
}

int ld_set_schedule(ld_handle* h, int T, const float* c1, const float* c2, const float* logvar, const float* sigma) {
  if (!h || T <= 0 || !c1 || !c2 || (!logvar && !sigma)) return fail(LD_ERR_INVALID, "bad schedule");
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  Engine& E = h->E;
  CK(cudaSetDevice(E.device));
  std::vector<float> sg(T);
  for (int i = 0; i < T; ++i) sg[i] = sigma ? sigma[i] : expf(0.5f * logvar[i]);  // ddpm.py:853
  if (E.coef1) { cudaFree(E.coef1); cudaFree(E.coef2); cudaFree(E.sigma); }
  CK(cudaMalloc(&E.coef1, T * sizeof(float))); CK(cudaMalloc(&E.coef2, T * sizeof(float))); CK(cudaMalloc(&E.sigma, T * sizeof(float)));
  CK(cudaMemcpy(E.coef1, c1, T * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(E.coef2, c2, T * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(E.sigma, sg.data(), T * sizeof(float), cudaMemcpyHostToDevice));
  E.T = T;
  return 0;
}

// Objective of the denoiser (ddpm.py:534-536): pred_x0 (a == NULL), or pred_noise / pred_v through the per-timestep pair (a, b) with
// x0 = a[t] * x_t - b[t] * model_output: (sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod) resp. (sqrt_alphas_cumprod,
// sqrt_one_minus_alphas_cumprod) (ddpm.py:631-653).  Only single-trajectory sampling supports them -- like the reference, whose branch
// path dies with UnboundLocalError for anything but pred_x0 (ddpm.py:731-761).
int ld_set_objective(ld_handle* h, int T, const float* a, const float* b) {
  if (!h) return fail(LD_ERR_INVALID, "null handle");
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  Engine& E = h->E;
  CK(cudaSetDevice(E.device));
  if (E.obj_a) { cudaFree(E.obj_a); cudaFree(E.obj_b); E.obj_a = E.obj_b = nullptr; }
  if (!a) return 0;
  if (!b || T <= 0) return fail(LD_ERR_INVALID, "bad objective tables");
  if (E.T && T != E.T) return fail(LD_ERR_INVALID, "objective tables must have the schedule's length (%d)", E.T);
  CK(cudaMalloc(&E.obj_a, T * sizeof(float))); CK(cudaMalloc(&E.obj_b, T * sizeof(float)));
  CK(cudaMemcpy(E.obj_a, a, T * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(E.obj_b, b, T * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

static int need_ready(Engine& E) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  if (!E.finalized) return fail(LD_ERR_STATE, "weights not finalized");
  if (cudaSetDevice(E.device) != cudaSuccess) return fail(LD_ERR_CUDA, "cudaSetDevice failed");
  return 0;
}

// plans for the stand-alone entry points read/write engine-owned staging buffers
static int get_staged(ld_handle* h, int N, int H, int W, Staged** out) {
  Engine& E = h->E;
  char k[64]; snprintf(k, sizeof k, "%dx%dx%d", N, H, W);
  auto it = E.staged.find(k);
  if (it != E.staged.end()) { *out = &it->second; return 0; }
  // keep at most one staged shape per handle (the workspace can be large)
  for (auto& kv : E.staged) { kv.second.cond.reset(); kv.second.unet.reset(); kv.second.free_all(); }
  E.staged.clear();
  Staged& S = E.staged[k];
  const size_t img = (size_t)N * H * W * sizeof(float);
  const int fh = H / E.cond_div, fw = W / E.cond_div;
  CK(cudaMalloc(&S.x, img)); CK(cudaMalloc(&S.c, img)); CK(cudaMalloc(&S.o, img));
  CK(cudaMalloc(&S.feat, (size_t)N * fh * fw * E.cond_C * E.esz()));
  CK(cudaMalloc(&S.t, N * sizeof(int64_t)));
  S.cond.reset(new Plan()); S.unet.reset(new Plan());
  int rc = build_cond_plan(E, *S.cond, N, H, W, S.c, S.feat);
  if (rc) return rc;
  rc = build_unet_plan(E, *S.unet, N, H, W, S.x, S.feat, S.t, nullptr, S.o);
  if (rc) return rc;
  *out = &S;
  return 0;
}

int ld_unet_forward(ld_handle* h, const float* x, const float* cond, const int64_t* t, float* out, int N, int H, int W, void* stream) {
  if (!h || !x || !cond || !t || !out) return fail(LD_ERR_INVALID, "null argument");
  Engine& E = h->E;
  int rc = need_ready(E); if (rc) return rc;
  if ((rc = check_shape(E, H, W))) return rc;
  Staged* S; if ((rc = get_staged(h, N, H, W, &S))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t img = (size_t)N * H * W * sizeof(float);
  CK(cudaMemcpyAsync(S->x, x, img, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(S->c, cond, img, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(S->t, t, N * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  if ((rc = run_plan(E, *S->cond, s))) return rc;
  if ((rc = run_plan(E, *S->unet, s))) return rc;
  CK(cudaMemcpyAsync(out, S->o, img, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int ld_cond_encode(ld_handle* h, const float* cond, float* feat, int N, int H, int W, void* stream) {
  if (!h || !cond || !feat) return fail(LD_ERR_INVALID, "null argument");
  Engine& E = h->E;
  int rc = need_ready(E); if (rc) return rc;
  if ((rc = check_shape(E, H, W))) return rc;
  Staged* S; if ((rc = get_staged(h, N, H, W, &S))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(S->c, cond, (size_t)N * H * W * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if ((rc = run_plan(E, *S->cond, s))) return rc;
  const int fh = H / E.cond_div, fw = W / E.cond_div;
  E.launches += launch_nhwc_to_nchw_f32(S->feat, feat, N, fh * fw, E.cond_C, E.bf, s);
  return 0;
}

// ---- sampler ----------------------------------------------------------------------------------
static int dalloc(Engine& E, void** p, size_t bytes) {
  CK(cudaMalloc(p, bytes));
  E.ss.allocs.push_back(*p);
  return 0;
}

// Sampler state of one (batch, height, width): buffers + the single-trajectory plans.  The branched plans are built on
// demand and kept next to them, so alternating anomalous masks (branched) with all-ones masks (vanilla DDPM, ddpm.py:1110-1117),
// as the reference's test loop does, re-uses everything.
static int prepare_sampler(Engine& E, const ld_sample_desc& sd) {
  char k[128];
  const bool pair_unet = sd.branch_out && !(sd.mask_x && sd.ood_uses_cond);
  snprintf(k, sizeof k, "%dx%dx%d", sd.batch, sd.height, sd.width);
  auto& S = E.ss;
  const size_t n = (size_t)sd.batch * sd.height * sd.width;
  const int B = sd.batch, H = sd.height, W = sd.width;
  int rc;
  if (S.key != k) {
    E.free_samp(); E.plans.clear();
    S.B = B; S.H = H; S.W = W;
    const int fh = H / E.cond_div, fw = W / E.cond_div;
    const size_t featB = (size_t)B * fh * fw * E.cond_C * E.esz();
    if ((rc = dalloc(E, (void**)&S.xs, 2 * n * 4))) return rc;
    if ((rc = dalloc(E, (void**)&S.o, 2 * n * 4))) return rc;
    if ((rc = dalloc(E, (void**)&S.bm, n * 4))) return rc;
    if ((rc = dalloc(E, (void**)&S.cond, n * 4))) return rc;
    if ((rc = dalloc(E, (void**)&S.cond_out, 2 * n * 4))) return rc;  // [cond_out; cond_in] contiguous
    S.cond_in = S.cond_out + n;
    if ((rc = dalloc(E, &S.feat_pair, 2 * featB))) return rc;
    if ((rc = dalloc(E, &S.feat_full, featB))) return rc;
    if ((rc = dalloc(E, (void**)&S.counters, 8 * sizeof(unsigned int)))) return rc;   // [0..3] assertion counters, [4] step ticket
    if ((rc = dalloc(E, (void**)&S.t_dev, sizeof(int)))) return rc;
    auto p = std::unique_ptr<Plan>(new Plan());
    if ((rc = build_cond_plan(E, *p, B, H, W, S.cond, S.feat_full))) return rc;
    E.plans["cond_full"] = std::move(p);
    p.reset(new Plan());
    if ((rc = build_unet_plan(E, *p, B, H, W, S.xs, S.feat_full, nullptr, S.t_dev, S.o))) return rc;
    E.plans["unet_full"] = std::move(p);
    S.key = k;
  }
  if (sd.branch_out) {
    // conditional-feature plans are hoisted out of the loop (the encoder does not depend on t);
    // branched UNet: x = [x_out; x_in], or x_in only when the OOD output is discarded (ddpm.py:704-708)
    const std::string ck = pair_unet ? "cond_pair/2" : "cond_pair/1", uk = pair_unet ? "unet_pair/2" : "unet_pair/1";
    if (!E.plans.count(uk)) {
      auto p = std::unique_ptr<Plan>(new Plan());
      if (pair_unet) rc = build_cond_plan(E, *p, 2 * B, H, W, S.cond_out, S.feat_pair);
      else rc = build_cond_plan(E, *p, B, H, W, S.cond_in, S.feat_pair);
      if (rc) return rc;
      E.plans[ck] = std::move(p);
      p.reset(new Plan());
      if (pair_unet) rc = build_unet_plan(E, *p, 2 * B, H, W, S.xs, S.feat_pair, nullptr, S.t_dev, S.o);
      else rc = build_unet_plan(E, *p, B, H, W, S.xs + n, S.feat_pair, nullptr, S.t_dev, S.o + n);
      if (rc) return rc;
      E.plans[uk] = std::move(p);
    }
  }
  return 0;
}

// FiLM rows for t = 0 .. T-1 (once per handle and T): the same two kernels the per-step path runs, on T "images"
static int ensure_film_table(Engine& E, int T, cudaStream_t s) {
  if (E.film_tab && E.film_tab_T >= T) return 0;
  if (E.film_tab) { cudaFree(E.film_tab); E.film_tab = nullptr; E.film_tab_T = 0; }
  if (E.film_total & 3) return 0;   // (row gathers are 16-byte vectors)
  float* tab = nullptr; float* st = nullptr; int64_t* tt = nullptr;
  const int td = 4 * E.d.dim;
  CK(cudaMalloc(&tab, (size_t)T * E.film_total * sizeof(float)));
  CK(cudaMalloc(&st, (size_t)T * td * sizeof(float)));
  CK(cudaMalloc(&tt, (size_t)T * sizeof(int64_t)));
  std::vector<int64_t> th(T);
  for (int i = 0; i < T; ++i) th[i] = i;
  CK(cudaMemcpyAsync(tt, th.data(), (size_t)T * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  TimeP tp{};
  tp.t = tt; tp.t_scalar = nullptr; tp.N = T; tp.dim = E.d.dim; tp.theta = E.d.sinusoidal_theta;
  tp.neg_step = (float)(-(std::log((double)E.d.sinusoidal_theta) / (double)(E.d.dim / 2 - 1)));
  tp.w1 = E.tw1; tp.b1 = E.tb1; tp.w2 = E.tw2; tp.b2 = E.tb2; tp.st = st;
  tp.wf = E.film_w; tp.bf_ = E.film_b; tp.total = E.film_total; tp.film = tab;
  E.launches += launch_time_film(tp, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(st); cudaFree(tt);
  if (e != cudaSuccess) { cudaFree(tab); return fail(LD_ERR_CUDA, "FiLM table failed: %s", cudaGetErrorString(e)); }
  E.film_tab = tab; E.film_tab_T = T;
  return 0;
}

static StepP make_step(Engine& E, const ld_sample_desc& sd, int kind, const float* noise, float* x0_trace) {
  auto& S = E.ss;
  const long long n = (long long)sd.batch * sd.height * sd.width;
  StepP p{};
  p.kind = kind;
  p.o_out = S.o; p.o_in = S.o + n;
  p.x_out = S.xs; p.x_in = S.xs + n;
  p.bm = S.bm; p.cond_out = S.cond_out; p.z = noise;
  p.t_ptr = S.t_dev; p.coef1 = E.coef1; p.coef2 = E.coef2; p.sigma = E.sigma;
  p.mask_x = sd.mask_x; p.ood_uses_cond = sd.ood_uses_cond; p.lo = sd.min_val; p.hi = sd.max_val;
  p.n = n; p.z_stride = n; p.tloop = sd.num_timesteps; p.counters = S.counters;
  p.x0_trace = x0_trace; p.trace_stride = 2 * n;
  p.ticket = S.counters + 4;   // the step kernel also does `t -= 1` (ddpm.py:951)
  p.ca = E.obj_a; p.cb = E.obj_b;
  return p;
}

static int capture(Engine& E, cudaGraphExec_t* exec, const std::function<int(cudaStream_t)>& body) {
  cudaStream_t s = E.own_stream;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = body(s);
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(s, &g);
  if (rc) { if (g) cudaGraphDestroy(g); return rc; }
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  return 0;
}

// The asserts of the reference (ddpm.py:698, 790) and the soft-max shift check of the fused LinearAttention, from the counters
// the loop left on the host.  *la_retry is set when the call has to be repeated on the exact-max LinearAttention path.
static int finish_checks(Engine& E, const ld_sample_desc& sd, bool will_fuse, const unsigned int* cnt, unsigned int la_under, bool* la_retry) {
  if (la_retry) *la_retry = false;
  if (la_under) {
    cudaMemset(E.la_flag, 0, sizeof(unsigned int));
    if (la_retry && !E.opt_la_exact) { *la_retry = true; return 0; }
    return fail(LD_ERR_STATE, "LinearAttention soft-max shift underflowed (%u rows)", la_under);
  }
  if (sd.branch_out && sd.mask_x && (cnt[0] == 0 || cnt[1] == 0)) return fail(LD_ERR_MASK, "mask should be binary");
  if (will_fuse && !(cnt[2] > 0 && cnt[3] > 0)) return fail(LD_ERR_MASK, "x_out and x_in should be masked");
  return 0;
}

// The fused LinearAttention shifts its soft-max over the pixels by an analytic bound instead of the true maximum (ld_linattn_tc.cu);
// with extreme to_qkv weights a whole row of weights can underflow.  The kernels count such rows; the loop looks at the counter
// after its first timestep and at its end, and on a hit the engine switches to the exact-max path for good (plans rebuilt) and
// repeats the call -- no user action, no wrong result.
static int switch_to_exact_linattn(Engine& E) {
  E.opt_la_exact = 1;
  drop_plans(E);
  return 0;
}

static int sample_impl(ld_handle* h, const ld_sample_desc& sd, const float* cond, const float* mask, const float* noise, float* out,
                       float* x0_trace, cudaStream_t cs, bool allow_retry) {
  Engine& E = h->E;
  int rc;
  if ((rc = prepare_sampler(E, sd))) return rc;
  if ((rc = ensure_film_table(E, E.T, E.own_stream))) return rc;
  auto& S = E.ss;
  cudaStream_t s = E.own_stream;
  const long long n = (long long)sd.batch * sd.height * sd.width;
  const bool pair_unet = sd.branch_out && !(sd.mask_x && sd.ood_uses_cond);
  // hand over from the caller's stream to the engine stream
  CK(cudaEventRecord(E.ev_in, cs));
  CK(cudaStreamWaitEvent(s, E.ev_in, 0));
  CK(cudaMemsetAsync(S.counters, 0, 8 * sizeof(unsigned int), s));
  CK(cudaMemcpyAsync(S.cond, cond, n * 4, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(S.xs, noise, n * 4, cudaMemcpyDeviceToDevice, s));          // x_T (ddpm.py:935)
  const int t0 = sd.num_timesteps - 1;
  CK(cudaMemcpyAsync(S.t_dev, &t0, sizeof(int), cudaMemcpyHostToDevice, s));
  bool branched = sd.branch_out != 0;
  Plan* up = branched ? E.plans[pair_unet ? "unet_pair/2" : "unet_pair/1"].get() : nullptr;
  Plan* uf = E.plans["unet_full"].get();
  if (branched) {
    CK(cudaMemcpyAsync(S.xs + n, noise, n * 4, cudaMemcpyDeviceToDevice, s));    // img = [img, img] (ddpm.py:957)
    PrepP pp{}; pp.cond = S.cond; pp.mask = mask; pp.bm = S.bm; pp.cond_out = S.cond_out; pp.cond_in = S.cond_in;
    pp.floor = sd.cond_in_floor; pp.n = n; pp.counters = S.counters;
    E.launches += launch_prep_cond(pp, s);
    if ((rc = run_plan(E, *E.plans[pair_unet ? "cond_pair/2" : "cond_pair/1"], s))) return rc;
  }
  const bool will_fuse = branched && sd.start_intermediate && sd.start_timestep >= 0;
  if (!branched || will_fuse) { if ((rc = run_plan(E, *E.plans["cond_full"], s))) return rc; }

  auto body_branch = [&](cudaStream_t st) -> int {
    int r = run_plan(E, *up, st); if (r) return r;
    StepP p = make_step(E, sd, 0, noise, x0_trace);
    if (!pair_unet) p.o_out = nullptr;
    E.launches += launch_step(p, st);
    return 0;
  };
  auto body_single = [&](cudaStream_t st) -> int {
    int r = run_plan(E, *uf, st); if (r) return r;
    StepP p = make_step(E, sd, 2, noise, x0_trace);
    E.launches += launch_step(p, st);
    return 0;
  };
  // graphs bake the tape / trace pointers: re-capture on every call (cheap next to T replays)
  if (S.g_branch) { cudaGraphExecDestroy(S.g_branch); S.g_branch = nullptr; }
  if (S.g_single) { cudaGraphExecDestroy(S.g_single); S.g_single = nullptr; }
  const bool use_graph = E.opt_use_graph != 0;
  int64_t per_branch = 0, per_single = 0;
  if (use_graph) {
    int64_t l0 = E.launches;
    if (branched) { if ((rc = capture(E, &S.g_branch, body_branch))) return rc; per_branch = E.launches - l0; }
    l0 = E.launches;
    if ((rc = capture(E, &S.g_single, body_single))) return rc;
    per_single = E.launches - l0;
    E.launches -= per_branch + per_single;  // captured, not launched
  }
  unsigned int la_under = 0;
  for (int t = sd.num_timesteps - 1; t >= 0; --t) {
    if (branched) {
      const bool fuse = sd.start_intermediate && t <= sd.start_timestep;  // ddpm.py:779
      if (fuse) {
        if ((rc = run_plan(E, *up, s))) return rc;
        StepP p = make_step(E, sd, 1, noise, x0_trace);
        if (!pair_unet) p.o_out = nullptr;
        E.launches += launch_step(p, s);
        branched = false;  // config['branch_out'] = False (ddpm.py:780)
      } else if (use_graph) {
        CK(cudaGraphLaunch(S.g_branch, s)); E.launches += per_branch;
      } else if ((rc = body_branch(s))) return rc;
    } else {
      if (use_graph) { CK(cudaGraphLaunch(S.g_single, s)); E.launches += per_single; }
      else if ((rc = body_single(s))) return rc;
    }
    if (t == sd.num_timesteps - 1 && sd.num_timesteps > 8 && allow_retry && !E.opt_la_exact && !E.opt_async) {
      // early look at the LinearAttention underflow counter: a long chain is not run to its end on a path that has to be repeated
      CK(cudaMemcpyAsync(&la_under, E.la_flag, sizeof la_under, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      if (la_under) {
        cudaMemset(E.la_flag, 0, sizeof(unsigned int));
        if ((rc = switch_to_exact_linattn(E))) return rc;
        return sample_impl(h, sd, cond, mask, noise, out, x0_trace, cs, false);
      }
    }
  }
  // result (ddpm.py:964-970)
  CK(cudaMemcpyAsync(out, S.xs, n * 4, cudaMemcpyDeviceToDevice, s));
  if (sd.return_pair) CK(cudaMemcpyAsync(out + n, branched ? S.xs + n : S.xs, n * 4, cudaMemcpyDeviceToDevice, s));
  CK(cudaEventRecord(E.ev_out, s));
  CK(cudaStreamWaitEvent(cs, E.ev_out, 0));
  if (E.opt_async) {   // no host synchronisation: the checks wait for ld_sample_finish
    E.pending = true; E.pend_sd = sd; E.pend_fuse = will_fuse;
    return 0;
  }
  unsigned int cnt[4];
  CK(cudaMemcpyAsync(cnt, S.counters, sizeof cnt, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&la_under, E.la_flag, sizeof la_under, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  bool retry = false;
  if ((rc = finish_checks(E, sd, will_fuse, cnt, la_under, allow_retry ? &retry : nullptr))) return rc;
  if (retry) {
    if ((rc = switch_to_exact_linattn(E))) return rc;
    return sample_impl(h, sd, cond, mask, noise, out, x0_trace, cs, false);
  }
  return 0;
}

int ld_sample(ld_handle* h, const ld_sample_desc* sdp, const float* cond, const float* mask, const float* noise, float* out,
              float* x0_trace, void* stream) {
  if (!h || !sdp || !cond || !noise || !out) return fail(LD_ERR_INVALID, "null argument");
  Engine& E = h->E;
  const ld_sample_desc sd = *sdp;
  int rc = need_ready(E); if (rc) return rc;
  if (!E.T) return fail(LD_ERR_STATE, "schedule not set");
  if (E.pending) return fail(LD_ERR_STATE, "a deferred call is pending: call ld_sample_finish first");
  if (sd.num_timesteps < 1 || sd.num_timesteps > E.T) return fail(LD_ERR_INVALID, "num_timesteps out of range");
  if (sd.branch_out && !mask) return fail(LD_ERR_INVALID, "branch mode needs a mask");
  if (sd.branch_out && E.obj_a) return fail(LD_ERR_INVALID, "branch sampling needs objective pred_x0 (ddpm.py:731-761)");
  if ((rc = check_shape(E, sd.height, sd.width))) return rc;
  return sample_impl(h, sd, cond, mask, noise, out, x0_trace, (cudaStream_t)stream, true);
}

// Deferred half of an "async" ld_sample / ld_sample_ddim: waits for the engine stream and reports what the synchronous call would have.
int ld_sample_finish(ld_handle* h) {
  if (!h) return fail(LD_ERR_INVALID, "null handle");
  Engine& E = h->E;
  int rc = need_ready(E); if (rc) return rc;
  if (!E.pending) return 0;
  E.pending = false;
  unsigned int cnt[4], la_under = 0;
  CK(cudaMemcpyAsync(cnt, E.ss.counters, sizeof cnt, cudaMemcpyDeviceToHost, E.own_stream));
  CK(cudaMemcpyAsync(&la_under, E.la_flag, sizeof la_under, cudaMemcpyDeviceToHost, E.own_stream));
  CK(cudaStreamSynchronize(E.own_stream));
  return finish_checks(E, E.pend_sd, E.pend_fuse, cnt, la_under, nullptr);
}

// DDIM variant of the branch sampler (ddpm.py:979-1075): same UNet plans and masks, the schedule of (time, coefficient)
// pairs comes from the host (diffusion.py builds it with the reference's own tensor ops), the step index lives on the device.
static int ddim_impl(ld_handle* h, const ld_sample_desc& sd, const float* cond, const float* mask, const float* noise, float* out,
                     const int32_t* times, const float* coefs, int nsteps, int fuse_step, cudaStream_t cs, bool allow_retry) {
  Engine& E = h->E;
  int rc;
  if ((rc = prepare_sampler(E, sd))) return rc;
  if (E.T && (rc = ensure_film_table(E, E.T, E.own_stream))) return rc;
  auto& S = E.ss;
  cudaStream_t s = E.own_stream;
  const long long n = (long long)sd.batch * sd.height * sd.width;
  const bool pair_unet = sd.branch_out && !(sd.mask_x && sd.ood_uses_cond);
  // device copy of the schedule + step index (owned by the sampler state: an async call outlives this function)
  int* times_d = nullptr; float* coefs_d = nullptr; int* idx_d = nullptr;
  {
    void* blk = nullptr;
    const size_t bytes = (size_t)nsteps * 4 + (size_t)nsteps * 5 * 4 + 16;
    if (S.ddim_blk && S.ddim_bytes >= bytes) blk = S.ddim_blk;
    else {
      if (S.ddim_blk) { CK(cudaStreamSynchronize(s)); cudaFree(S.ddim_blk); S.ddim_blk = nullptr; }
      CK(cudaMalloc(&blk, bytes));
      S.ddim_blk = blk; S.ddim_bytes = bytes;
    }
    idx_d = (int*)blk; times_d = idx_d + 4; coefs_d = (float*)(times_d + nsteps);
  }
  CK(cudaEventRecord(E.ev_in, cs));
  CK(cudaStreamWaitEvent(s, E.ev_in, 0));
  CK(cudaMemcpyAsync(times_d, times, (size_t)nsteps * sizeof(int), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(coefs_d, coefs, (size_t)nsteps * 5 * sizeof(float), cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(idx_d, 0, sizeof(int), s));
  CK(cudaMemcpyAsync(S.t_dev, times, sizeof(int), cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(S.counters, 0, 8 * sizeof(unsigned int), s));
  CK(cudaMemcpyAsync(S.cond, cond, n * 4, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(S.xs, noise, n * 4, cudaMemcpyDeviceToDevice, s));          // img = randn(shape) (ddpm.py:989)
  bool branched = sd.branch_out != 0;
  Plan* up = branched ? E.plans[pair_unet ? "unet_pair/2" : "unet_pair/1"].get() : nullptr;
  Plan* uf = E.plans["unet_full"].get();
  if (branched) {
    CK(cudaMemcpyAsync(S.xs + n, noise, n * 4, cudaMemcpyDeviceToDevice, s));    // img = [img, img] (ddpm.py:1003-1004)
    PrepP pp{}; pp.cond = S.cond; pp.mask = mask; pp.bm = S.bm; pp.cond_out = S.cond_out; pp.cond_in = S.cond_in;
    pp.floor = sd.cond_in_floor; pp.n = n; pp.counters = S.counters;
    E.launches += launch_prep_cond(pp, s);
    if ((rc = run_plan(E, *E.plans[pair_unet ? "cond_pair/2" : "cond_pair/1"], s))) return rc;
  }
  const bool will_fuse = branched && fuse_step >= 0 && fuse_step < nsteps - 1;   // the last step never fuses (ddpm.py:1009-1012)
  if (!branched || will_fuse) { if ((rc = run_plan(E, *E.plans["cond_full"], s))) return rc; }
  auto mk = [&](int kind) {
    DdimP p{};
    p.kind = kind; p.o_out = pair_unet || kind == 2 ? S.o : nullptr; p.o_in = S.o + n; p.x_out = S.xs; p.x_in = S.xs + n;
    p.bm = S.bm; p.cond_out = S.cond_out; p.z = noise; p.z_stride = n; p.idx_ptr = idx_d; p.nsteps = nsteps; p.coefs = coefs_d;
    p.mask_x = sd.mask_x; p.ood_uses_cond = sd.ood_uses_cond; p.lo = sd.min_val; p.hi = sd.max_val; p.n = n; p.counters = S.counters;
    p.ticket = S.counters + 4; p.times = times_d; p.t_ptr = S.t_dev;   // the step kernel also advances (idx, t) (ddpm.py:996-998)
    p.ca = E.obj_a; p.cb = E.obj_b;
    return p;
  };
  auto body = [&](int kind, cudaStream_t st) -> int {
    int r = run_plan(E, kind == 2 ? *uf : *up, st); if (r) return r;
    E.launches += launch_ddim_step(mk(kind), st);
    return 0;
  };
  if (S.g_branch) { cudaGraphExecDestroy(S.g_branch); S.g_branch = nullptr; }
  if (S.g_single) { cudaGraphExecDestroy(S.g_single); S.g_single = nullptr; }
  const bool use_graph = E.opt_use_graph != 0;
  int64_t per_branch = 0, per_single = 0;
  if (use_graph) {
    int64_t l0 = E.launches;
    if (branched) { if ((rc = capture(E, &S.g_branch, [&](cudaStream_t st) { return body(0, st); }))) return rc; per_branch = E.launches - l0; }
    l0 = E.launches;
    if ((rc = capture(E, &S.g_single, [&](cudaStream_t st) { return body(2, st); }))) return rc;
    per_single = E.launches - l0;
    E.launches -= per_branch + per_single;
  }
  for (int i = 0; i < nsteps; ++i) {
    if (branched) {
      if (will_fuse && i >= fuse_step) {
        if ((rc = body(1, s))) return rc;
        branched = false;   // config['branch_out'] = False (ddpm.py:1023)
      } else if (use_graph) {
        CK(cudaGraphLaunch(S.g_branch, s)); E.launches += per_branch;
      } else if ((rc = body(0, s))) return rc;
    } else {
      if (use_graph) { CK(cudaGraphLaunch(S.g_single, s)); E.launches += per_single; }
      else if ((rc = body(2, s))) return rc;
    }
  }
  CK(cudaMemcpyAsync(out, S.xs, n * 4, cudaMemcpyDeviceToDevice, s));
  if (sd.return_pair) CK(cudaMemcpyAsync(out + n, branched ? S.xs + n : S.xs, n * 4, cudaMemcpyDeviceToDevice, s));
  CK(cudaEventRecord(E.ev_out, s));
  CK(cudaStreamWaitEvent(cs, E.ev_out, 0));
  if (E.opt_async) {
    E.pending = true; E.pend_sd = sd; E.pend_fuse = will_fuse;
    return 0;
  }
  unsigned int cnt[4], la_under = 0;
  CK(cudaMemcpyAsync(cnt, S.counters, sizeof cnt, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&la_under, E.la_flag, sizeof la_under, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  bool retry = false;
  if ((rc = finish_checks(E, sd, will_fuse, cnt, la_under, allow_retry ? &retry : nullptr))) return rc;
  if (retry) {
    if ((rc = switch_to_exact_linattn(E))) return rc;
    return ddim_impl(h, sd, cond, mask, noise, out, times, coefs, nsteps, fuse_step, cs, false);
  }
  return 0;
}

int ld_sample_ddim(ld_handle* h, const ld_sample_desc* sdp, const float* cond, const float* mask, const float* noise, float* out,
                   const int32_t* times, const float* coefs, int nsteps, int fuse_step, void* stream) {
  if (!h || !sdp || !cond || !noise || !out || !times || !coefs) return fail(LD_ERR_INVALID, "null argument");
  Engine& E = h->E;
  const ld_sample_desc sd = *sdp;
  int rc = need_ready(E); if (rc) return rc;
  if (E.pending) return fail(LD_ERR_STATE, "a deferred call is pending: call ld_sample_finish first");
  if (nsteps < 1) return fail(LD_ERR_INVALID, "nsteps out of range");
  if (!E.T) return fail(LD_ERR_STATE, "schedule not set");
  for (int i = 0; i < nsteps; ++i)   // times index the per-timestep FiLM table and the schedule
    if (times[i] < 0 || times[i] >= E.T) return fail(LD_ERR_INVALID, "times[%d] = %d is outside [0, %d)", i, times[i], E.T);
  if (sd.branch_out && !mask) return fail(LD_ERR_INVALID, "branch mode needs a mask");
  if (sd.branch_out && E.obj_a) return fail(LD_ERR_INVALID, "branch sampling needs objective pred_x0 (ddpm.py:731-761)");
  if ((rc = check_shape(E, sd.height, sd.width))) return rc;
  return ddim_impl(h, sd, cond, mask, noise, out, times, coefs, nsteps, fuse_step, (cudaStream_t)stream, true);
}

int ld_posterior_step(ld_handle* h, int kind, int t, float* x_out, float* x_in, float* x0_out, float* x0_in, const float* cond,
                      const float* mask, const float* z, const ld_sample_desc* sd, int64_t n, void* stream) {
  if (!h || !sd || !x_out || !x0_out) return fail(LD_ERR_INVALID, "null argument");
  Engine& E = h->E;
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  if (!E.T || t < 0 || t >= E.T) return fail(LD_ERR_STATE, "schedule not set or t out of range");
  if (n <= 0 || n % 4) return fail(LD_ERR_INVALID, "n must be a positive multiple of 4");
  CK(cudaSetDevice(E.device));
  if (kind != 2 && (!mask || !cond || !x_in || !x0_in)) return fail(LD_ERR_INVALID, "branched step needs mask, cond and both branches");
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc;
  float *bm = nullptr, *co = nullptr, *ci = nullptr, *oo = nullptr, *oi = nullptr; unsigned int* cnt = nullptr; int* td = nullptr;
  CK(sc.get(&bm, n * 4)); CK(sc.get(&co, n * 4)); CK(sc.get(&ci, n * 4));
  CK(sc.get(&oo, n * 4)); CK(sc.get(&oi, n * 4));
  CK(sc.get(&cnt, 32)); CK(sc.get(&td, 4));
  CK(cudaMemsetAsync(cnt, 0, 32, s));
  CK(cudaMemcpyAsync(td, &t, 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(oo, x0_out, n * 4, cudaMemcpyDeviceToDevice, s));
  if (x0_in) CK(cudaMemcpyAsync(oi, x0_in, n * 4, cudaMemcpyDeviceToDevice, s));
  if (kind != 2) {
    PrepP pp{}; pp.cond = cond; pp.mask = mask; pp.bm = bm; pp.cond_out = co; pp.cond_in = ci; pp.floor = sd->cond_in_floor;
    pp.n = n; pp.counters = cnt;
    E.launches += launch_prep_cond(pp, s);
  }
  StepP p{};
  p.kind = kind; p.o_out = oo; p.o_in = oi; p.x_out = x_out; p.x_in = x_in; p.x0_out = x0_out; p.x0_in = x0_in;
  p.bm = bm; p.cond_out = co; p.z = z; p.t_ptr = td; p.coef1 = E.coef1; p.coef2 = E.coef2; p.sigma = E.sigma;
  p.mask_x = sd->mask_x; p.ood_uses_cond = sd->ood_uses_cond; p.lo = sd->min_val; p.hi = sd->max_val;
  p.n = n; p.z_stride = 0; p.tloop = t; p.counters = cnt; p.ticket = cnt + 4;
  E.launches += launch_step(p, s);
  CK(cudaStreamSynchronize(s));
  return 0;
}

// ---- test hooks ---------------------------------------------------------------------------------
static Plan* last_staged_unet(Engine& E) {
  if (E.staged.empty()) return nullptr;
  return E.staged.begin()->second.unet.get();
}
int ld_debug_num_taps(ld_handle* h) {
  Plan* P = h ? last_staged_unet(h->E) : nullptr;
  return P ? (int)P->tags.size() : 0;
}
int ld_debug_tap_info(ld_handle* h, int i, const char** name, int32_t dims[4]) {
  Plan* P = h ? last_staged_unet(h->E) : nullptr;
  if (!P || i < 0 || i >= (int)P->tags.size()) return fail(LD_ERR_INVALID, "tap index out of range");
  *name = P->tags[i].first.c_str();
  const Ten& t = P->tags[i].second;
  dims[0] = t.N; dims[1] = t.C; dims[2] = t.H; dims[3] = t.W;
  return 0;
}
int ld_debug_tap_fetch(ld_handle* h, int i, float* out_nchw, void* stream) {
  Plan* P = h ? last_staged_unet(h->E) : nullptr;
  if (!P || i < 0 || i >= (int)P->tags.size()) return fail(LD_ERR_INVALID, "tap index out of range");
  const Ten& t = P->tags[i].second;
  launch_nhwc_to_nchw_f32(t.p, out_nchw, t.N, t.H * t.W, t.C, h->E.bf, (cudaStream_t)stream);
  return 0;
}

int ld_debug_conv(int kernel, const float* x0, int C0, const float* x1, int C1, int N, int Hin, int Win, int up, int H, int W,
                  const float* w_host, const float* bias_host, int Cout, int ks, const float* res, float* out, void* stream) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  cudaStream_t s = (cudaStream_t)stream;
  const bool bf = kernel != 0;
  const int Cin = C0 + C1, taps = ks * ks;
  const size_t esz = bf ? 2 : 4;
  std::vector<float> pk((size_t)taps * Cin * Cout);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t) pk[((size_t)t * Cin + c) * Cout + o] = w_host[((size_t)o * Cin + c) * taps + t];
  float *dw = nullptr, *db = nullptr; void *a0 = nullptr, *a1 = nullptr, *ar = nullptr, *ao = nullptr;
  const size_t nin = (size_t)N * Hin * Win, nout = (size_t)N * H * W;
  CK(cudaMalloc(&dw, pk.size() * 4)); CK(cudaMemcpy(dw, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
  if (bias_host) { CK(cudaMalloc(&db, Cout * 4)); CK(cudaMemcpy(db, bias_host, Cout * 4, cudaMemcpyHostToDevice)); }
  CK(cudaMalloc(&a0, nin * C0 * esz)); launch_convert(x0, false, a0, bf, (long long)nin * C0, s);
  if (x1) { CK(cudaMalloc(&a1, nin * C1 * esz)); launch_convert(x1, false, a1, bf, (long long)nin * C1, s); }
  if (res) { CK(cudaMalloc(&ar, nout * Cout * esz)); launch_convert(res, false, ar, bf, (long long)nout * Cout, s); }
  CK(cudaMalloc(&ao, nout * Cout * esz));
  int rc = 0;
  if (kernel == 3) {   // nearest x2 + 3x3 folded into one low-resolution convolution with pixel-shuffle output
    ConvTcW tw;
    ConvTcArgs ta; ta.src0 = a0; ta.C0 = C0; ta.N = N; ta.H = Hin; ta.W = Win; ta.Hin = Hin; ta.Win = Win; ta.ps = Cout; ta.dst = ao;
    if (!up || ks != 3 || x1 || res || H != 2 * Hin || W != 2 * Win) rc = fail(LD_ERR_INVALID, "kernel 3 is the up-sampling 3x3 convolution");
    else if (conv_tc_pack_up2(pk.data(), bias_host, Cin, Cout, &tw) || !tw.ready) rc = fail(LD_ERR_INVALID, "conv_tc_pack_up2: unsupported shape");
    else if (conv_tc_launch(tw, ta, s) < 0) rc = fail(LD_ERR_INVALID, "conv_tc_launch: unsupported arguments");
    cudaStreamSynchronize(s);
    cudaFree(tw.w); cudaFree(tw.w32); cudaFree(tw.bias);
  } else if (kernel == 2) {
    ConvTcW tw;
    if (conv_tc_pack(pk.data(), bias_host, Cin, Cout, ks, 1, ks / 2, &tw) || !tw.ready) rc = fail(LD_ERR_INVALID, "conv_tc_pack: unsupported shape");
    else {
      ConvTcArgs ta; ta.src0 = a0; ta.C0 = C0; ta.src1 = a1; ta.C1 = C1; ta.N = N; ta.H = H; ta.W = W; ta.Hin = Hin; ta.Win = Win;
      ta.up = up; ta.dst = ao; ta.res = ar;
      if (conv_tc_launch(tw, ta, s) < 0) rc = fail(LD_ERR_INVALID, "conv_tc_launch: unsupported arguments");
    }
    cudaStreamSynchronize(s);
    cudaFree(tw.w); cudaFree(tw.w32); cudaFree(tw.bias);
  } else {
    ConvP p{};
    p.src0 = a0; p.C0 = C0; p.src1 = a1; p.C1 = C1; p.N = N; p.H = H; p.W = W; p.Hin = Hin; p.Win = Win;
    p.ks = ks; p.stride = 1; p.pad = ks / 2; p.up = up; p.w = dw; p.bias = db; p.Cout = Cout; p.dst = ao; p.res = ar;
    p.M = (long long)nout;
    launch_conv_simt(p, bf, s);
  }
  launch_convert(ao, bf, out, false, (long long)nout * Cout, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(dw); cudaFree(db); cudaFree(a0); cudaFree(a1); cudaFree(ar); cudaFree(ao);
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "debug conv failed: %s", cudaGetErrorString(e));
  return rc;
}

// 3x3 tcgen05 convolution with the fused GroupNorm prologue / statistics epilogue (test hook).
int ld_debug_conv_fused(const float* x0, int C0, int N, int H, int W, const float* w_host, const float* bias_host, int Cout,
                        const double* pro_stats, const float* pro_gamma, const float* pro_beta, const float* pro_film,
                        int pro_film_stride, int pro_G, int pro_act, double* stats_out, int stats_G, float* out, void* stream) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  cudaStream_t s = (cudaStream_t)stream;
  const int taps = 9;
  std::vector<float> pk((size_t)taps * C0 * Cout);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < C0; ++c)
      for (int t = 0; t < taps; ++t) pk[((size_t)t * C0 + c) * Cout + o] = w_host[((size_t)o * C0 + c) * taps + t];
  const size_t npx = (size_t)N * H * W;
  void *a0 = nullptr, *ao = nullptr; float* abd = nullptr;
  CK(cudaMalloc(&a0, npx * C0 * 2)); launch_convert(x0, false, a0, true, (long long)npx * C0, s);
  CK(cudaMalloc(&ao, npx * Cout * 2));
  if (stats_out) CK(cudaMemsetAsync(stats_out, 0, (size_t)N * stats_G * 2 * sizeof(double), s));
  ConvTcW tw;
  int rc = 0;
  if (conv_tc_pack(pk.data(), bias_host, C0, Cout, 3, 1, 1, &tw) || !tw.ready) rc = fail(LD_ERR_INVALID, "conv_tc_pack: unsupported shape");
  else {
    ConvTcArgs ta; ta.src0 = a0; ta.C0 = C0; ta.N = N; ta.H = H; ta.W = W; ta.Hin = H; ta.Win = W; ta.dst = ao;
    if (pro_stats) {
      CK(cudaMalloc(&abd, (size_t)N * 2 * C0 * 4));
      gn_coef_launch(pro_stats, pro_gamma, pro_beta, pro_film, pro_film_stride, pro_G, C0, N, (long long)H * W, 1e-5f, abd, s);
      ta.pro_ab = abd; ta.pro_act = pro_act;
    }
    ta.stats = stats_out; ta.stats_G = stats_G;
    if (conv_tc_launch(tw, ta, s) < 0) rc = fail(LD_ERR_INVALID, "conv_tc_launch: unsupported arguments");
  }
  launch_convert(ao, true, out, false, (long long)npx * Cout, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(tw.w); cudaFree(tw.w32); cudaFree(tw.bias); cudaFree(a0); cudaFree(ao); cudaFree(abd);
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "debug fused conv failed: %s", cudaGetErrorString(e));
  return rc;
}

// 3x3 tcgen05 convolution of the virtual concat [x0 | x1] with GroupNorm statistics, fused with the 1x1 convolution of
// the same input as a second output: ResnetBlock block1.proj + res_conv (ddpm.py:207,212) (test hook).
int ld_debug_conv_dual(const float* x0, int C0, const float* x1, int C1, int N, int H, int W, const float* w3_host,
                       const float* b3_host, const float* w1_host, const float* b1_host, int Cout, double* stats_out, int stats_G,
                       float* out, float* out2, void* stream) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  cudaStream_t s = (cudaStream_t)stream;
  const int Cin = C0 + C1;
  std::vector<float> p3((size_t)9 * Cin * Cout), p1((size_t)Cin * Cout);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c) {
      for (int t = 0; t < 9; ++t) p3[((size_t)t * Cin + c) * Cout + o] = w3_host[((size_t)o * Cin + c) * 9 + t];
      p1[(size_t)c * Cout + o] = w1_host[(size_t)o * Cin + c];
    }
  const size_t npx = (size_t)N * H * W;
  void *a0 = nullptr, *a1 = nullptr, *ao = nullptr, *ao2 = nullptr;
  CK(cudaMalloc(&a0, npx * C0 * 2)); launch_convert(x0, false, a0, true, (long long)npx * C0, s);
  if (C1) { CK(cudaMalloc(&a1, npx * C1 * 2)); launch_convert(x1, false, a1, true, (long long)npx * C1, s); }
  CK(cudaMalloc(&ao, npx * Cout * 2)); CK(cudaMalloc(&ao2, npx * Cout * 2));
  CK(cudaMemsetAsync(stats_out, 0, (size_t)N * stats_G * 2 * sizeof(double), s));
  ConvTcW tw;
  int rc = 0;
  if (conv_tc_pack(p3.data(), b3_host, Cin, Cout, 3, 1, 1, &tw, p1.data(), b1_host) || !tw.ready)
    rc = fail(LD_ERR_INVALID, "conv_tc_pack: unsupported shape");
  else {
    ConvTcArgs ta; ta.src0 = a0; ta.C0 = C0; ta.src1 = a1; ta.C1 = C1; ta.N = N; ta.H = H; ta.W = W; ta.Hin = H; ta.Win = W;
    ta.dst = ao; ta.dst2 = ao2; ta.stats = stats_out; ta.stats_G = stats_G;
    if (conv_tc_launch(tw, ta, s) < 0) rc = fail(LD_ERR_INVALID, "conv_tc_launch: unsupported arguments");
  }
  launch_convert(ao, true, out, false, (long long)npx * Cout, s);
  launch_convert(ao2, true, out2, false, (long long)npx * Cout, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(tw.w); cudaFree(tw.w32); cudaFree(tw.bias); cudaFree(tw.bias2); cudaFree(a0); cudaFree(a1); cudaFree(ao); cudaFree(ao2);
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "debug dual conv failed: %s", cudaGetErrorString(e));
  return rc;
}

// Fused tcgen05 LinearAttention block, attn(x) + x (test hook).  x/out: fp32 NHWC device; weights: host, torch layout.
int ld_debug_linattn(const float* x, int C, int N, int HW, const float* wqkv, const float* g, const float* wout, const float* bout,
                     const float* g2, float* out, void* stream) {
  return ld_debug_linattn_h(x, C, N, HW, 4, wqkv, g, wout, bout, g2, out, stream);
}
int ld_debug_linattn_h(const float* x, int C, int N, int HW, int heads, const float* wqkv, const float* g, const float* wout, const float* bout,
                       const float* g2, float* out, void* stream) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  cudaStream_t s = (cudaStream_t)stream;
  LinAttnTcW w;
  if (linattn_tc_pack(wqkv, g, wout, bout, g2, C, heads, &w) || !w.ready) return fail(LD_ERR_INVALID, "linattn_tc_pack: unsupported shape");
  const size_t n = (size_t)N * HW * C, hid = (size_t)heads * 32;
  void *xb = nullptr, *ob = nullptr, *mn = nullptr; float *ctx = nullptr, *ks = nullptr; unsigned int* flag = nullptr;
  CK(cudaMalloc(&xb, n * 2)); CK(cudaMalloc(&ob, n * 2)); CK(cudaMalloc(&mn, (size_t)N * hid * C * 2));
  CK(cudaMalloc(&ctx, (size_t)N * hid * C * 4)); CK(cudaMalloc(&ks, (size_t)N * hid * 4)); CK(cudaMalloc(&flag, 4));
  CK(cudaMemsetAsync(ctx, 0, (size_t)N * hid * C * 4, s)); CK(cudaMemsetAsync(ks, 0, (size_t)N * hid * 4, s)); CK(cudaMemsetAsync(flag, 0, 4, s));
  launch_convert(x, false, xb, true, (long long)n, s);
  LinAttnTcArgs a; a.x = xb; a.out = ob; a.N = N; a.HW = HW; a.Z = ctx; a.ksum = ks; a.Mn = mn; a.flag = flag;
  int rc = linattn_tc_launch(w, a, s) < 0 ? fail(LD_ERR_INVALID, "linattn_tc_launch failed") : 0;
  launch_convert(ob, true, out, false, (long long)n, s);
  unsigned int under = 0;
  cudaMemcpyAsync(&under, flag, 4, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(xb); cudaFree(ob); cudaFree(mn); cudaFree(ctx); cudaFree(ks); cudaFree(flag);
  linattn_tc_free(&w);
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "debug linattn failed: %s", cudaGetErrorString(e));
  if (!rc && under) return fail(LD_ERR_STATE, "soft-max shift underflowed (%u rows)", under);
  return rc;
}

// tcgen05 flash attention (test hook).  qkv: fp32 [N][n][3*heads*32] device; out: fp32 [N][n][heads*32].
int ld_debug_attention(const float* qkv, int N, int n, int heads, float* out, void* stream) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  cudaStream_t s = (cudaStream_t)stream;
  if (attn_tc_configure()) return fail(LD_ERR_CUDA, "attn_tc_configure failed");
  const size_t nq = (size_t)N * n * 3 * heads * 32, no = (size_t)N * n * heads * 32;
  void *qb = nullptr, *ob = nullptr, *sc = nullptr;
  CK(cudaMalloc(&qb, nq * 2)); CK(cudaMalloc(&ob, no * 2)); CK(cudaMalloc(&sc, attn_tc_scratch_bytes(N, n, heads)));
  launch_convert(qkv, false, qb, true, (long long)nq, s);
  int rc = attn_tc_launch(qb, ob, sc, N, n, heads, s) < 0 ? fail(LD_ERR_INVALID, "attn_tc_launch failed") : 0;
  launch_convert(ob, true, out, false, (long long)no, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(qb); cudaFree(ob); cudaFree(sc);
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "debug attention failed: %s", cudaGetErrorString(e));
  return rc;
}

// tcgen05 7x7 single-channel convolution (init_conv) test hook.  x: fp32 [N][H][W] device; w_host [Cout][1][7][7]; out fp32 NHWC.
int ld_debug_conv7(const float* x, int N, int H, int W, const float* w_host, const float* bias_host, int Cout, float* out, void* stream) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<float> wt((size_t)49 * Cout);
  for (int o = 0; o < Cout; ++o)
    for (int t = 0; t < 49; ++t) wt[(size_t)t * Cout + o] = w_host[(size_t)o * 49 + t];
  Conv7TcW w;
  if (conv7_tc_pack(wt.data(), bias_host, Cout, &w) || !w.ready) return fail(LD_ERR_INVALID, "conv7_tc_pack: unsupported width");
  const size_t n = (size_t)N * H * W * Cout;
  void* ob = nullptr;
  CK(cudaMalloc(&ob, n * 2));
  int rc = conv7_tc_launch(w, x, ob, N, H, W, s) < 0 ? fail(LD_ERR_INVALID, "conv7_tc_launch failed") : 0;
  launch_convert(ob, true, out, false, (long long)n, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(ob);
  conv7_tc_free(&w);
  if (e != cudaSuccess) return fail(LD_ERR_CUDA, "debug conv7 failed: %s", cudaGetErrorString(e));
  return rc;
}

// Time one convolution kernel in isolation with CUDA events on its launch stream (bench.py roofline leg).
int ld_debug_conv_time(int kernel, int C0, int C1, int N, int H, int W, int up, int Cout, int ks, int iters, float* ms_out,
                       void* stream) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  cudaStream_t s = (cudaStream_t)stream;
  const bool bf = kernel != 0;
  const int Cin = C0 + C1, taps = ks * ks, Hin = up ? H / 2 : H, Win = up ? W / 2 : W;
  const size_t esz = bf ? 2 : 4, nin = (size_t)N * Hin * Win, nout = (size_t)N * H * W;
  std::vector<float> pk((size_t)taps * Cin * Cout), bias(Cout, 0.1f);
  for (size_t i = 0; i < pk.size(); ++i) pk[i] = (float)((int)(i * 2654435761u % 2001) - 1000) * 1e-4f;
  float *dw = nullptr, *db = nullptr, *tmp = nullptr; void *a0 = nullptr, *a1 = nullptr, *ao = nullptr;
  CK(cudaMalloc(&dw, pk.size() * 4)); CK(cudaMemcpy(dw, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&db, Cout * 4)); CK(cudaMemcpy(db, bias.data(), Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&a0, nin * C0 * esz)); CK(cudaMemsetAsync(a0, 0x3c, nin * C0 * esz, s));
  if (C1) { CK(cudaMalloc(&a1, nin * C1 * esz)); CK(cudaMemsetAsync(a1, 0x3c, nin * C1 * esz, s)); }
  CK(cudaMalloc(&ao, nout * Cout * esz));
  (void)tmp;
  ConvTcW tw;
  ConvTcArgs ta; ta.src0 = a0; ta.C0 = C0; ta.src1 = a1; ta.C1 = C1; ta.N = N; ta.H = H; ta.W = W; ta.Hin = Hin; ta.Win = Win;
  ta.up = up; ta.dst = ao; ta.res = nullptr;
  ConvP p{};
  p.src0 = a0; p.C0 = C0; p.src1 = a1; p.C1 = C1; p.N = N; p.H = H; p.W = W; p.Hin = Hin; p.Win = Win;
  p.ks = ks; p.stride = 1; p.pad = ks / 2; p.up = up; p.w = dw; p.bias = db; p.Cout = Cout; p.dst = ao; p.res = nullptr;
  p.M = (long long)nout;
  int rc = 0;
  if (kernel == 3) {
    ta = ConvTcArgs(); ta.src0 = a0; ta.C0 = C0; ta.N = N; ta.H = Hin; ta.W = Win; ta.Hin = Hin; ta.Win = Win; ta.ps = Cout; ta.dst = ao;
    if (!up || C1 || conv_tc_pack_up2(pk.data(), bias.data(), Cin, Cout, &tw) || !conv_tc_supports(tw, ta)) rc = fail(LD_ERR_INVALID, "conv_tc up2: unsupported shape");
  } else
  if (kernel == 2 && (conv_tc_pack(pk.data(), bias.data(), Cin, Cout, ks, 1, ks / 2, &tw) || !conv_tc_supports(tw, ta)))
    rc = fail(LD_ERR_INVALID, "conv_tc: unsupported shape");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  if (!rc) {
    for (int i = 0; i < 3; ++i) { if (kernel >= 2) conv_tc_launch(tw, ta, s); else launch_conv_simt(p, bf, s); }
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters; ++i) { if (kernel >= 2) conv_tc_launch(tw, ta, s); else launch_conv_simt(p, bf, s); }
    cudaEventRecord(e1, s);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / (float)iters;
    if (e != cudaSuccess) rc = fail(LD_ERR_CUDA, "conv timing failed: %s", cudaGetErrorString(e));
  }
  if (conv_tc_trace()) {   // development aid: dump the hand-shake timeline of CTA 0 (last launch)
    std::vector<long long> tr(4 * 64 * 4);
    cudaMemcpy(tr.data(), conv_tc_trace(), tr.size() * 8, cudaMemcpyDeviceToHost);
    long long t0 = 0;
    for (auto v : tr) if (v && (!t0 || v < t0)) t0 = v;
    const char* names[4] = {"prod0", "prod1", "mma", "epi"};
    for (int t = 0; t < 24; ++t)
      for (int r = 0; r < 4; ++r) {
        const long long* e = &tr[(r * 64 + t) * 4];
        if (e[0]) printf("tile %2d %-5s  %8lld %8lld %8lld %8lld\n", t, names[r], e[0] - t0, e[1] ? e[1] - t0 : -1, e[2] ? e[2] - t0 : -1, e[3] ? e[3] - t0 : -1);
      }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(tw.w); cudaFree(tw.w32); cudaFree(tw.bias);
  cudaFree(dw); cudaFree(db); cudaFree(a0); cudaFree(a1); cudaFree(ao);
  return rc;
}

// Time the dominant convolution family (3x3 tcgen05, bf16) in the variants the sampler actually launches (bench.py roofline leg):
// 0 plain, 1 + GroupNorm statistics of the output, 2 + normalise-on-load prologue (GroupNorm affine + FiLM + SiLU of the source)
// and statistics, 3 dual: 3x3 + 1x1 res_conv of the same virtual concat [C0 | C1], two outputs, statistics.
int ld_debug_conv_variant_time(int variant, int C0, int C1, int N, int H, int W, int Cout, int iters, float* ms_out, void* stream) {
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  if (variant < 0 || variant > 3 || !ms_out || iters < 1) return fail(LD_ERR_INVALID, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int Cin = C0 + C1, G = 8;
  const size_t npx = (size_t)N * H * W;
  std::vector<float> pk((size_t)9 * Cin * Cout), p1((size_t)Cin * Cout), bias(Cout, 0.1f), ab((size_t)N * 2 * C0);
  for (size_t i = 0; i < pk.size(); ++i) pk[i] = (float)((int)(i * 2654435761u % 2001) - 1000) * 1e-4f;
  for (size_t i = 0; i < p1.size(); ++i) p1[i] = (float)((int)(i * 40503u % 2001) - 1000) * 1e-3f;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C0; ++c) { ab[((size_t)n * 2) * C0 + c] = 1.0f + 0.01f * (float)(c % 7); ab[((size_t)n * 2 + 1) * C0 + c] = 0.05f * (float)(c % 5) - 0.1f; }
  void *a0 = nullptr, *a1 = nullptr, *ao = nullptr, *ao2 = nullptr; float* abd = nullptr; double* st = nullptr;
  ConvTcW tw;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
    cudaFree(a0); cudaFree(a1); cudaFree(ao); cudaFree(ao2); cudaFree(abd); cudaFree(st);
    cudaFree(tw.w); cudaFree(tw.w32); cudaFree(tw.bias); cudaFree(tw.bias2);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(LD_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
  CKC(cudaMalloc(&a0, npx * C0 * 2)); CKC(cudaMemsetAsync(a0, 0x3c, npx * C0 * 2, s));
  if (C1) { CKC(cudaMalloc(&a1, npx * C1 * 2)); CKC(cudaMemsetAsync(a1, 0x3c, npx * C1 * 2, s)); }
  CKC(cudaMalloc(&ao, npx * Cout * 2));
  if (variant == 3) CKC(cudaMalloc(&ao2, npx * Cout * 2));
  CKC(cudaMalloc(&st, (size_t)N * G * 2 * sizeof(double))); CKC(cudaMemsetAsync(st, 0, (size_t)N * G * 2 * sizeof(double), s));
  CKC(cudaMalloc(&abd, ab.size() * 4)); CKC(cudaMemcpyAsync(abd, ab.data(), ab.size() * 4, cudaMemcpyHostToDevice, s));
  ConvTcArgs ta; ta.src0 = a0; ta.C0 = C0; ta.src1 = a1; ta.C1 = C1; ta.N = N; ta.H = H; ta.W = W; ta.Hin = H; ta.Win = W; ta.dst = ao;
  if (variant >= 1) { ta.stats = st; ta.stats_G = G; }
  if (variant == 2) { ta.pro_ab = abd; ta.pro_act = 1; }
  if (variant == 3) ta.dst2 = ao2;
  if (conv_tc_pack(pk.data(), bias.data(), Cin, Cout, 3, 1, 1, &tw, variant == 3 ? p1.data() : nullptr, variant == 3 ? bias.data() : nullptr) ||
      !conv_tc_supports(tw, ta)) { cleanup(); return fail(LD_ERR_INVALID, "conv_tc: unsupported shape for variant %d", variant); }
  CKC(cudaEventCreate(&e0)); CKC(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) conv_tc_launch(tw, ta, s);
  CKC(cudaEventRecord(e0, s));
  for (int i = 0; i < iters; ++i) conv_tc_launch(tw, ta, s);
  CKC(cudaEventRecord(e1, s));
  CKC(cudaEventSynchronize(e1));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  *ms_out = ms / (float)iters;
#undef CKC
  cleanup();
  return 0;
}

// ---- stages in front of the sampler (ld_producers.cu) ------------------------------------------------------------------------
int ld_prep_mnist(const float* raw, float* hr, float* cond, int N, int S, void* stream) {
  if (!raw || !hr || !cond || N < 1 || S < 2) return fail(LD_ERR_INVALID, "bad argument");
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  launch_mnist_cond(raw, hr, cond, N, S, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(LD_ERR_CUDA, "ld_prep_mnist launch failed");
}
int ld_prep_mri(const float* raw, float* out, void* scratch, int N, int Hs, int Ws, int crop, float mean, float std, int translate_zero,
                void* stream) {
  if (!raw || !out || N < 1 || crop < 1 || crop > Hs || crop > Ws || !(std != 0.f)) return fail(LD_ERR_INVALID, "bad argument");
  if (translate_zero && !scratch) return fail(LD_ERR_INVALID, "translate_zero needs N * 4 bytes of scratch");
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  launch_mri_norm(raw, out, (unsigned int*)scratch, N, Hs, Ws, crop, mean, std, translate_zero, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(LD_ERR_CUDA, "ld_prep_mri launch failed");
}
int64_t ld_mask_scratch_bytes(int B, int S) { return (int64_t)mask_scratch_bytes(B, S); }
int ld_mask_from_anomaly(const float* amap, int B, int h, int w, int S, int rule, int manual_cols, float* mask_pred, float* binary_mask,
                         void* scratch, void* stream) {
  if (!amap || !mask_pred || !scratch || B < 1 || h < 1 || w < 1 || S < 1 || rule < 0 || rule > LD_MASK_MVTEC_GRID)
    return fail(LD_ERR_INVALID, "bad argument");
  if ((long long)B * S * S < 2) return fail(LD_ERR_INVALID, "the standard deviation needs at least two elements");
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  launch_mask_from_anomaly(amap, B, h, w, S, rule, manual_cols, mask_pred, binary_mask, scratch, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(LD_ERR_CUDA, "ld_mask_from_anomaly launch failed");
}

int64_t ld_knn_scratch_bytes(int M, int Nb, int D) { return (int64_t)knn_scratch_bytes(M, Nb, D); }
int ld_knn_min(const float* x, const float* bank, int M, int Nb, int D, float* score, int64_t* loc, void* scratch, void* stream) {
  if (!x || !bank || !score || !loc || !scratch || M < 1 || Nb < 1 || D < 1) return fail(LD_ERR_INVALID, "bad argument");
  if (ld_device_count() == 0) return fail(LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)");
  if (knn_tc_launch(x, bank, M, Nb, D, score, (long long*)loc, scratch, (cudaStream_t)stream) < 0) return fail(LD_ERR_CUDA, "knn_tc_launch failed");
  return cudaGetLastError() == cudaSuccess ? 0 : fail(LD_ERR_CUDA, "ld_knn_min launch failed");
}

int64_t ld_launch_count(const ld_handle* h) { return h ? h->E.launches : 0; }
int64_t ld_workspace_bytes(const ld_handle* h) {
  if (!h) return 0;
  int64_t b = 0;
  for (auto& kv : h->E.plans) b += (int64_t)kv.second->total_bytes;
  return b;
}
int ld_set_option(ld_handle* h, const char* name, int64_t value) {
  if (!h || !name) return fail(LD_ERR_INVALID, "null argument");
  Engine& E = h->E;
  if (!strcmp(name, "use_graph")) E.opt_use_graph = value;
  else if (!strcmp(name, "async")) E.opt_async = value;
  else if (!strcmp(name, "debug_keep")) E.opt_debug_keep = value;
  else if (!strcmp(name, "la_exact")) E.opt_la_exact = value;
  else if (!strcmp(name, "attn_simt")) E.opt_attn_simt = value;
  else if (!strcmp(name, "pdl")) pdl_flag() = (int)value;      // process-wide: programmatic dependent launch, 1 all kernels, 2 only after tiny kernels (ld_launch.cuh)
  else if (!strcmp(name, "up2")) E.opt_up2 = value;
  else if (!strcmp(name, "use_tc")) { if (E.finalized) return fail(LD_ERR_STATE, "use_tc must be set before finalize"); E.use_tc = value != 0 && E.bf; }
  else return fail(LD_ERR_INVALID, "unknown option '%s'", name);
  // options are read when a plan is built: drop every cached plan so that the new value takes effect on the next call
  if (E.finalized && ld_device_count() > 0 && cudaSetDevice(E.device) == cudaSuccess) drop_plans(E);
  return 0;
}

int ld_get_option(const ld_handle* h, const char* name, int64_t* value) {
  if (!h || !name || !value) return fail(LD_ERR_INVALID, "null argument");
  const Engine& E = h->E;
  if (!strcmp(name, "use_graph")) *value = E.opt_use_graph;
  else if (!strcmp(name, "async")) *value = E.opt_async;
  else if (!strcmp(name, "debug_keep")) *value = E.opt_debug_keep;
  else if (!strcmp(name, "la_exact")) *value = E.opt_la_exact;
  else if (!strcmp(name, "attn_simt")) *value = E.opt_attn_simt;
  else if (!strcmp(name, "pdl")) *value = pdl_flag();
  else if (!strcmp(name, "up2")) *value = E.opt_up2;
  else if (!strcmp(name, "use_tc")) *value = E.use_tc ? 1 : 0;
  else return fail(LD_ERR_INVALID, "unknown option '%s'", name);
  return 0;
}

}  // extern "C"
