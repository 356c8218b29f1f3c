// Wan2.1 I2V DiT forward (WanTransformer3DModel.forward, reference call site wan:910-917) as a native
// orchestrator over the sm_100a kernels: patch gather -> tcgen05 GEMMs with fused epilogues -> tcgen05 flash
// attention -> LayerNorm / RMSNorm+RoPE -> unpatchify.  The 2 or 3 CFG passes (wan:882-894) are batched along the
// token dimension; the x3 replicated / concatenated / cast model input of the reference is never materialised.
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "dit_kernels.cuh"

struct WRef {
  const void* ptr = nullptr;
  int64_t numel = 0;
  int dtype = 0;
};

struct AttnW {
  const void *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b, *norm_q, *norm_k;
  const void *add_k_w = nullptr, *add_k_b = nullptr, *add_v_w = nullptr, *add_v_b = nullptr, *norm_added_k = nullptr;
};
struct BlockW {
  const float* table;
  AttnW attn1, attn2;
  const float *norm2_w, *norm2_b;
  const void *ffn1_w, *ffn1_b, *ffn2_w, *ffn2_b;
};

struct alg_wan_engine {
  alg_wan_config_t cfg;
  std::map<std::string, WRef> weights;
  bool resolved = false;
  std::vector<BlockW> blocks;
  const void *patch_w, *patch_b;
  const float *te1_w, *te1_b, *te2_w, *te2_b;
  const void *tp_w, *tp_b, *tx1_w, *tx1_b, *tx2_w, *tx2_b;
  const float *in1_w, *in1_b, *in2_w, *in2_b;
  const void *if1_w, *if1_b, *if2_w, *if2_b;
  const float* head_table;
  const void *proj_w, *proj_b;
  double *rope_t = nullptr, *rope_h = nullptr, *rope_w = nullptr;
  int n_t = 0, n_h = 0, n_w = 0;
  void* debug_buf = nullptr;
  size_t debug_bytes = 0;
  // optional per-class device timing (CUDA events on the launching stream)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  // Step-invariant cross-attention context (text / image embedders and the K, V^T projections of all layers): memoised per
  // (prompt pointers, image pointer, pass layout) when the caller enables it -- the conditioning does not change between the
  // denoise steps of one video (wan:844-944), so steps after the first of each pass layout skip ~10 TFLOP of projections.
  struct CtxSlot {
    const void* text[3] = {nullptr, nullptr, nullptr};
    const void* image = nullptr;
    int n_pass = 0, n_img = 0;
    bool valid = false;
    char* buf = nullptr;  // [layers] x (k_text | vt_text | k_img | vt_img)
    size_t bytes = 0;
    uint64_t stamp = 0;
  };
  bool ctx_caching = false;
  CtxSlot ctx_slots[2];  // the three-pass and the two-pass layouts of one video
  uint64_t ctx_clock = 0;
  int dim() const { return cfg.num_heads * cfg.head_dim; }
};

namespace {
using namespace alg;
typedef __nv_bfloat16 bf16;

enum { PROF_SELF_ATTN = 0, PROF_CROSS_ATTN = 1, PROF_GEMM = 2, PROF_ELEMENTWISE = 3, PROF_NUM = 4 };
thread_local alg_wan_engine* tl_engine = nullptr;  // engine whose forward is running on this thread (profiling)

cudaEvent_t take_event(alg_wan_engine* e) {
  if (e->ev_used == e->ev_pool.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    e->ev_pool.push_back(ev);
  }
  return e->ev_pool[e->ev_used++];
}
struct ProfScope {  // times everything enqueued on `st` during its lifetime as one span of class `cls`
  alg_wan_engine* e;
  cudaStream_t st;
  alg_wan_engine::Span sp;
  ProfScope(alg_wan_engine* e_, cudaStream_t st_, int cls) : e(e_), st(st_) {
    if (!e || !e->profiling) return;
    sp.cls = cls;
    sp.a = take_event(e);
    sp.b = take_event(e);
    cudaEventRecord(sp.a, st);
  }
  ~ProfScope() {
    if (!e || !e->profiling) return;
    cudaEventRecord(sp.b, st);
    e->spans.push_back(sp);
  }
};

struct Bump {
  char* base;
  size_t off = 0, cap;
  Bump(void* b, size_t c) : base(reinterpret_cast<char*>(b)), cap(c) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

int gemm(cudaStream_t st, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int64_t M,
         int64_t N, int64_t K, const void* bias, int epi = ALG_EPI_NONE, const void* R = nullptr,
         const float* gate = nullptr, int bias_per_row = 0) {
  ProfScope ps(tl_engine, st, PROF_GEMM);
  alg_gemm_t g{};
  g.A = A; g.B = B; g.D = D; g.bias = bias; g.R = R; g.gate = gate;
  g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldd = ldd;
  g.rows_per_batch = M; g.gate_ld = 0; g.epilogue = epi; g.bias_per_row = bias_per_row; g.out_f32 = 0;
  return alg_gemm_bf16(&g, st);
}

int attention(cudaStream_t st, const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int batch, int heads, int hd,
              int64_t n_q, int64_t n_kv, int64_t kv_pad, int accumulate) {
  ProfScope ps(tl_engine, st, n_q == n_kv ? PROF_SELF_ATTN : PROF_CROSS_ATTN);
  alg_attention_t a{};
  const int64_t d = (int64_t)heads * hd;
  a.Q = Q; a.K = K; a.Vt = Vt; a.O = O;
  a.batch = batch; a.heads = heads; a.head_dim = hd; a.n_q = n_q; a.n_kv = n_kv;
  a.q_bs = n_q * d; a.q_rs = d; a.k_bs = n_kv * d; a.k_rs = d; a.v_bs = d * kv_pad; a.v_rs = kv_pad;
  a.o_bs = n_q * d; a.o_rs = d;
  a.scale = 1.0f / sqrtf((float)hd);
  a.accumulate = accumulate;
  return alg_attention_bf16(&a, st);
}

// FP32LayerNorm (+affine | +AdaLN modulate) of the Wan block: the shared LayerNorm kernel in its fp32-chain mode
int layer_norm(cudaStream_t st, const bf16* x, bf16* out, int64_t rows, int d, float eps, const float* w, const float* b,
               const float* scale, const float* shift) {
  alg_layer_norm_t p{};
  p.x = x; p.out = out; p.rows = rows; p.d = d; p.eps = eps;
  p.weight = w; p.bias = b; p.affine_dtype = ALG_F32;
  p.scale = scale; p.shift = shift; p.mod_dtype = ALG_F32;
  p.rows_per_batch = rows > 0 ? rows : 1;
  return alg_layer_norm(&p, st);
}

inline int64_t pad8(int64_t n) { return (n + 7) & ~int64_t(7); }

#define ALG_TRY(expr)          \
  do {                         \
    if (int _rc = (expr)) return _rc; \
  } while (0)
#define ALG_TRY_EW(expr)                            \
  do {                                              \
    ProfScope _ps(tl_engine, st, PROF_ELEMENTWISE); \
    if (int _rc = (expr)) return _rc;               \
  } while (0)

int resolve(alg_wan_engine* e) {
  if (e->resolved) return 0;
  const alg_wan_config_t& c = e->cfg;
  const int64_t d = e->dim();
  std::string missing;
  auto get = [&](const std::string& name, int64_t numel, int dtype) -> const void* {
    auto it = e->weights.find(name);
    if (it == e->weights.end()) {
      if (missing.empty()) missing = name;
      return nullptr;
    }
    if (it->second.numel != numel || it->second.dtype != dtype) {
      if (missing.empty())
        missing = name + " (expected " + std::to_string(numel) + " elements of dtype " + std::to_string(dtype) + ", got " +
                  std::to_string(it->second.numel) + " of dtype " + std::to_string(it->second.dtype) + ")";
      return nullptr;
    }
    return it->second.ptr;
  };
  auto f32 = [&](const std::string& n, int64_t numel) { return reinterpret_cast<const float*>(get(n, numel, ALG_F32)); };
  const int64_t pk = (int64_t)c.in_channels * c.patch_t * c.patch_h * c.patch_w;
  e->patch_w = get("patch_embedding.weight", d * pk, ALG_BF16);
  e->patch_b = get("patch_embedding.bias", d, ALG_BF16);
  const std::string ce = "condition_embedder.";
  e->te1_w = f32(ce + "time_embedder.linear_1.weight", d * c.freq_dim);
  e->te1_b = f32(ce + "time_embedder.linear_1.bias", d);
  e->te2_w = f32(ce + "time_embedder.linear_2.weight", d * d);
  e->te2_b = f32(ce + "time_embedder.linear_2.bias", d);
  e->tp_w = get(ce + "time_proj.weight", 6 * d * d, ALG_BF16);
  e->tp_b = get(ce + "time_proj.bias", 6 * d, ALG_BF16);
  e->tx1_w = get(ce + "text_embedder.linear_1.weight", d * c.text_dim, ALG_BF16);
  e->tx1_b = get(ce + "text_embedder.linear_1.bias", d, ALG_BF16);
  e->tx2_w = get(ce + "text_embedder.linear_2.weight", d * d, ALG_BF16);
  e->tx2_b = get(ce + "text_embedder.linear_2.bias", d, ALG_BF16);
  if (c.image_dim > 0) {
    const std::string ie = ce + "image_embedder.";
    e->in1_w = f32(ie + "norm1.weight", c.image_dim);
    e->in1_b = f32(ie + "norm1.bias", c.image_dim);
    e->if1_w = get(ie + "ff.net.0.proj.weight", (int64_t)c.image_dim * c.image_dim, ALG_BF16);
    e->if1_b = get(ie + "ff.net.0.proj.bias", c.image_dim, ALG_BF16);
    e->if2_w = get(ie + "ff.net.2.weight", d * c.image_dim, ALG_BF16);
    e->if2_b = get(ie + "ff.net.2.bias", d, ALG_BF16);
    e->in2_w = f32(ie + "norm2.weight", d);
    e->in2_b = f32(ie + "norm2.bias", d);
  }
  e->blocks.resize(c.num_layers);
  for (int i = 0; i < c.num_layers; ++i) {
    const std::string p = "blocks." + std::to_string(i) + ".";
    BlockW& b = e->blocks[i];
    b.table = f32(p + "scale_shift_table", 6 * d);
    for (int a = 0; a < 2; ++a) {
      const std::string ap = p + (a == 0 ? "attn1." : "attn2.");
      AttnW& w = a == 0 ? b.attn1 : b.attn2;
      w.q_w = get(ap + "to_q.weight", d * d, ALG_BF16);
      w.q_b = get(ap + "to_q.bias", d, ALG_BF16);
      w.k_w = get(ap + "to_k.weight", d * d, ALG_BF16);
      w.k_b = get(ap + "to_k.bias", d, ALG_BF16);
      w.v_w = get(ap + "to_v.weight", d * d, ALG_BF16);
      w.v_b = get(ap + "to_v.bias", d, ALG_BF16);
      w.o_w = get(ap + "to_out.0.weight", d * d, ALG_BF16);
      w.o_b = get(ap + "to_out.0.bias", d, ALG_BF16);
      w.norm_q = get(ap + "norm_q.weight", d, ALG_BF16);
      w.norm_k = get(ap + "norm_k.weight", d, ALG_BF16);
      if (a == 1 && c.image_dim > 0) {
        w.add_k_w = get(ap + "add_k_proj.weight", d * d, ALG_BF16);
        w.add_k_b = get(ap + "add_k_proj.bias", d, ALG_BF16);
        w.add_v_w = get(ap + "add_v_proj.weight", d * d, ALG_BF16);
        w.add_v_b = get(ap + "add_v_proj.bias", d, ALG_BF16);
        w.norm_added_k = get(ap + "norm_added_k.weight", d, ALG_BF16);
      }
    }
    b.norm2_w = f32(p + "norm2.weight", d);
    b.norm2_b = f32(p + "norm2.bias", d);
    b.ffn1_w = get(p + "ffn.net.0.proj.weight", (int64_t)c.ffn_dim * d, ALG_BF16);
    b.ffn1_b = get(p + "ffn.net.0.proj.bias", c.ffn_dim, ALG_BF16);
    b.ffn2_w = get(p + "ffn.net.2.weight", d * c.ffn_dim, ALG_BF16);
    b.ffn2_b = get(p + "ffn.net.2.bias", d, ALG_BF16);
  }
  e->head_table = f32("scale_shift_table", 2 * d);
  const int64_t po = (int64_t)c.out_channels * c.patch_t * c.patch_h * c.patch_w;
  e->proj_w = get("proj_out.weight", po * d, ALG_BF16);
  e->proj_b = get("proj_out.bias", po, ALG_BF16);
  ALG_REQUIRE(missing.empty(), "wan engine: parameter missing or mis-shaped: " + missing);
  e->resolved = true;
  return 0;
}

struct Sizes {
  int64_t N, M, d, Npad, img_pad, txt_pad;
};

size_t plan(const alg_wan_engine* e, int n_pass, int T, int H, int W, int n_img, Bump* b, void** out_ptrs) {
  // single source of truth for the workspace layout: called with a null base to size it, then with the real base
  const alg_wan_config_t& c = e->cfg;
  const int64_t d = e->dim();
  const int64_t N = (int64_t)T * (H / c.patch_h) * (W / c.patch_w), M = N * n_pass;
  const int64_t pk = (int64_t)c.in_channels * c.patch_t * c.patch_h * c.patch_w;
  int i = 0;
  out_ptrs[i++] = b->take<bf16>(M * pad8(pk));               // 0 Apatch
  out_ptrs[i++] = b->take<bf16>(M * d);                      // 1 x
  out_ptrs[i++] = b->take<bf16>(M * d);                      // 2 h
  out_ptrs[i++] = b->take<bf16>(M * d);                      // 3 q
  out_ptrs[i++] = b->take<bf16>(M * d);                      // 4 k
  out_ptrs[i++] = b->take<bf16>((int64_t)n_pass * d * pad8(N));  // 5 vt
  out_ptrs[i++] = b->take<bf16>(M * d);                      // 6 attn out
  out_ptrs[i++] = b->take<bf16>(M * (int64_t)c.ffn_dim);     // 7 ffn hidden
  out_ptrs[i++] = b->take<float>(c.freq_dim);                // 8 sinusoid
  out_ptrs[i++] = b->take<float>(d);                         // 9 te hidden
  out_ptrs[i++] = b->take<float>(d);                         // 10 te out
  out_ptrs[i++] = b->take<bf16>(d);                          // 11 temb
  out_ptrs[i++] = b->take<bf16>(d);                          // 12 silu(temb)
  out_ptrs[i++] = b->take<bf16>(6 * d);                      // 13 tproj
  out_ptrs[i++] = b->take<float>(6 * d);                     // 14 mod
  out_ptrs[i++] = b->take<bf16>((int64_t)n_pass * c.text_len * d);  // 15 text hidden
  out_ptrs[i++] = b->take<bf16>((int64_t)n_pass * c.text_len * d);  // 16 ctx_text
  out_ptrs[i++] = b->take<bf16>((int64_t)n_pass * c.text_len * d);  // 17 k_text
  out_ptrs[i++] = b->take<bf16>((int64_t)n_pass * d * pad8(c.text_len));  // 18 vt_text
  const int64_t ni = std::max(n_img, 1);
  out_ptrs[i++] = b->take<bf16>(ni * std::max<int64_t>(c.image_dim, 8));  // 19 img normed
  out_ptrs[i++] = b->take<bf16>(ni * std::max<int64_t>(c.image_dim, 8));  // 20 img hidden
  out_ptrs[i++] = b->take<bf16>(ni * d);                                  // 21 img proj
  out_ptrs[i++] = b->take<bf16>((int64_t)n_pass * ni * d);                // 22 ctx_img (replicated per pass)
  out_ptrs[i++] = b->take<bf16>((int64_t)n_pass * ni * d);                // 23 k_img
  out_ptrs[i++] = b->take<bf16>((int64_t)n_pass * d * pad8(ni));          // 24 vt_img
  out_ptrs[i++] = b->take<bf16>(M * pad8((int64_t)c.out_channels * c.patch_t * c.patch_h * c.patch_w));  // 25 proj
  return b->off;
}

}  // namespace

extern "C" int alg_wan_create(const alg_wan_config_t* cfg, alg_wan_engine_t** out) {
  using namespace alg;
  ALG_REQUIRE(cfg && out, "wan_create: null pointer");
  ALG_REQUIRE(cfg->head_dim == 128 || cfg->head_dim == 64, "wan_create: head_dim must be 64 or 128");
  ALG_REQUIRE(cfg->patch_t == 1 && cfg->patch_h == 2 && cfg->patch_w == 2, "wan_create: patch size must be (1, 2, 2)");
  ALG_REQUIRE((cfg->num_heads * cfg->head_dim) % 64 == 0 && cfg->text_dim % 8 == 0 && cfg->ffn_dim % 8 == 0 &&
                  cfg->image_dim % 8 == 0 && cfg->freq_dim % 2 == 0,
              "wan_create: widths must be multiples of 8");
  alg_wan_engine* e = new alg_wan_engine();
  e->cfg = *cfg;
  // WanRotaryPosEmbed: per-axis complex128 tables, theta = 10000
  const int hd = cfg->head_dim;
  const int h_dim = 2 * (hd / 6), w_dim = h_dim, t_dim = hd - h_dim - w_dim;
  const int dims[3] = {t_dim, h_dim, w_dim};
  double** dst[3] = {&e->rope_t, &e->rope_h, &e->rope_w};
  int* cnt[3] = {&e->n_t, &e->n_h, &e->n_w};
  for (int a = 0; a < 3; ++a) {
    const int half = dims[a] / 2;
    *cnt[a] = half;
    std::vector<double> tab((size_t)cfg->rope_max_seq_len * half * 2);
    for (int pos = 0; pos < cfg->rope_max_seq_len; ++pos)
      for (int k = 0; k < half; ++k) {
        const double freq = 1.0 / std::pow(10000.0, (double)(2 * k) / (double)dims[a]);
        const double ang = (double)pos * freq;
        tab[((size_t)pos * half + k) * 2] = std::cos(ang);
        tab[((size_t)pos * half + k) * 2 + 1] = std::sin(ang);
      }
    ALG_CUDA_OK(cudaMalloc(dst[a], tab.size() * sizeof(double)));
    ALG_CUDA_OK(cudaMemcpy(*dst[a], tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  *out = e;
  return 0;
}

extern "C" void alg_wan_destroy(alg_wan_engine_t* e) {
  if (!e) return;
  for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
  cudaFree(e->rope_t);
  cudaFree(e->rope_h);
  cudaFree(e->rope_w);
  for (auto& s : e->ctx_slots) cudaFree(s.buf);
  delete e;
}

extern "C" int alg_wan_set_weight(alg_wan_engine_t* e, const char* name, const void* ptr, int64_t numel, int dtype) {
  using namespace alg;
  ALG_REQUIRE(e && name && ptr, "wan_set_weight: null pointer");
  ALG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, std::string("wan_set_weight: ") + name + " is not 16-byte aligned");
  WRef w;
  w.ptr = ptr;
  w.numel = numel;
  w.dtype = dtype;
  e->weights[name] = w;
  e->resolved = false;
  return 0;
}

extern "C" int alg_wan_weights_complete(alg_wan_engine_t* e) {
  using namespace alg;
  ALG_REQUIRE(e, "wan_weights_complete: null engine");
  return resolve(e);
}

extern "C" int alg_wan_set_debug_buffer(alg_wan_engine_t* e, void* buf, size_t bytes) {
  using namespace alg;
  ALG_REQUIRE(e, "wan_set_debug_buffer: null engine");
  e->debug_buf = buf;
  e->debug_bytes = bytes;
  return 0;
}

extern "C" int alg_wan_workspace_bytes(alg_wan_engine_t* e, int n_pass, int T, int H, int W, int n_img_tokens,
                                       size_t* bytes) {
  using namespace alg;
  ALG_REQUIRE(e && bytes, "wan_workspace_bytes: null pointer");
  Bump b(nullptr, ~size_t(0));
  void* ptrs[32];
  *bytes = plan(e, n_pass, T, H, W, n_img_tokens, &b, ptrs) + 256;
  return 0;
}

extern "C" int alg_wan_forward(alg_wan_engine_t* e, const float* const* latents, const float* const* cond,
                               const void* const* text, const void* image, int n_img, int n_pass, int T, int H, int W,
                               int64_t timestep, void* noise_out, void* workspace, size_t workspace_bytes,
                               void* stream) {
  using namespace alg;
  ALG_REQUIRE(e && latents && cond && text && noise_out && workspace, "wan_forward: null pointer");
  ALG_REQUIRE(n_pass >= 1 && n_pass <= 3, "wan_forward: n_pass must be 1, 2 or 3");
  const alg_wan_config_t& c = e->cfg;
  ALG_REQUIRE(H % c.patch_h == 0 && W % c.patch_w == 0 && T >= 1, "wan_forward: latent grid not divisible by the patch size");
  ALG_REQUIRE(T <= c.rope_max_seq_len && H / 2 <= c.rope_max_seq_len && W / 2 <= c.rope_max_seq_len,
              "wan_forward: grid exceeds rope_max_seq_len");
  ALG_REQUIRE((n_img > 0) == (c.image_dim > 0 && image != nullptr), "wan_forward: image tokens / image_dim mismatch");
  ALG_TRY(resolve(e));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  struct TlGuard {
    TlGuard(alg_wan_engine* x) { tl_engine = x; }
    ~TlGuard() { tl_engine = nullptr; }
  } tl_guard(e);
  const int64_t d = e->dim();
  const int heads = c.num_heads, hd = c.head_dim;
  const int lat_ch = c.out_channels, cond_ch = c.in_channels - c.out_channels;
  const int pph = H / c.patch_h, ppw = W / c.patch_w;
  const int64_t N = (int64_t)T * pph * ppw, M = N * n_pass;
  const int64_t pk = (int64_t)c.in_channels * 4, pkp = pad8(pk);
  const int64_t Npad = pad8(N), txt = c.text_len, txt_pad = pad8(txt), ni = n_img, ni_pad = pad8(std::max(n_img, 1));
  const int64_t po = (int64_t)c.out_channels * 4, pop = pad8(po);

  Bump bump(workspace, workspace_bytes);
  void* P[32];
  const size_t need = plan(e, n_pass, T, H, W, n_img, &bump, P);
  ALG_REQUIRE(need <= workspace_bytes, "wan_forward: workspace too small (" + std::to_string(workspace_bytes) + " < " +
                                           std::to_string(need) + ")");
  ALG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "wan_forward: workspace must be 256-byte aligned");
  bf16 *Apatch = (bf16*)P[0], *x = (bf16*)P[1], *h = (bf16*)P[2], *q = (bf16*)P[3], *k = (bf16*)P[4], *vt = (bf16*)P[5],
       *ao = (bf16*)P[6], *ffn = (bf16*)P[7];
  float *sinus = (float*)P[8], *te_h = (float*)P[9], *te_o = (float*)P[10];
  bf16 *temb = (bf16*)P[11], *silu_temb = (bf16*)P[12], *tproj = (bf16*)P[13];
  float* mod = (float*)P[14];
  bf16 *txt_h = (bf16*)P[15], *ctx_text = (bf16*)P[16], *k_text = (bf16*)P[17], *vt_text = (bf16*)P[18];
  bf16 *img_n = (bf16*)P[19], *img_h = (bf16*)P[20], *img_p = (bf16*)P[21], *ctx_img = (bf16*)P[22],
       *k_img = (bf16*)P[23], *vt_img = (bf16*)P[24], *proj = (bf16*)P[25];

  size_t dbg_off = 0;
  auto debug_dump = [&](const void* src, size_t bytes) {
    if (e->debug_buf && dbg_off + bytes <= e->debug_bytes) {
      cudaMemcpyAsync(reinterpret_cast<char*>(e->debug_buf) + dbg_off, src, bytes, cudaMemcpyDeviceToDevice, st);
      dbg_off += bytes;
    }
  };

  // ---- 1. patch embedding: gather (concat + cast fused) + GEMM (Conv3d k = s = (1,2,2)) -------------------
  ALG_REQUIRE(pkp == pk && pop == po, "wan_forward: in_channels * 4 and out_channels * 4 must be multiples of 8");
  dit::CondPtrs cp{};
  for (int p = 0; p < n_pass; ++p) {
    ALG_REQUIRE(latents[p] && cond[p] && text[p], "wan_forward: null per-pass input");
    cp.lat[p] = latents[p];
    cp.p[p] = cond[p];
  }
  ALG_TRY_EW(dit::patch_gather(cp, n_pass, lat_ch, cond_ch, T, H, W, Apatch, st));
  ALG_TRY(gemm(st, Apatch, pk, e->patch_w, pk, x, d, M, d, pk, e->patch_b));
  debug_dump(x, (size_t)M * d * 2);

  // ---- context cache lookup (see alg_wan_engine::CtxSlot) ---------------------------------------------------
  const size_t kt_elems = (size_t)n_pass * txt * d, vtt_elems = (size_t)n_pass * d * txt_pad;
  const size_t ki_elems = n_img > 0 ? (size_t)n_pass * ni * d : 0, vti_elems = n_img > 0 ? (size_t)n_pass * d * ni_pad : 0;
  auto al = [](size_t n) { return (n + 127) & ~size_t(127); };  // 256-byte aligned sections (bf16 elements)
  const size_t layer_elems = al(kt_elems) + al(vtt_elems) + al(ki_elems) + al(vti_elems);
  alg_wan_engine::CtxSlot* slot = nullptr;
  bool ctx_hit = false;
  if (e->ctx_caching) {
    for (auto& s : e->ctx_slots) {
      bool same = s.valid && s.n_pass == n_pass && s.n_img == n_img && s.image == image;
      for (int p = 0; same && p < n_pass; ++p) same = s.text[p] == text[p];
      if (same) {
        slot = &s;
        ctx_hit = true;
      }
    }
    if (!slot) {  // miss: take the least recently used slot and fill it during this forward
      slot = e->ctx_slots[0].stamp <= e->ctx_slots[1].stamp ? &e->ctx_slots[0] : &e->ctx_slots[1];
      const size_t want = layer_elems * c.num_layers * sizeof(bf16);
      if (slot->bytes < want) {
        if (slot->buf) {
          ALG_CUDA_OK(cudaStreamSynchronize(st));
          ALG_CUDA_OK(cudaFree(slot->buf));
          slot->buf = nullptr;
          slot->bytes = 0;
        }
        ALG_CUDA_OK(cudaMalloc(&slot->buf, want));
        slot->bytes = want;
      }
      slot->valid = false;  // becomes valid once every layer has been written (below)
      slot->n_pass = n_pass;
      slot->n_img = n_img;
      slot->image = image;
      for (int p = 0; p < 3; ++p) slot->text[p] = p < n_pass ? text[p] : nullptr;
    }
    slot->stamp = ++e->ctx_clock;
  }

  // ---- 2. condition embedder --------------------------------------------------------------------------
  ALG_TRY_EW(dit::timestep_sinusoid((float)timestep, c.freq_dim, sinus, st));
  ALG_TRY_EW(dit::gemv_f32(e->te1_w, e->te1_b, sinus, te_h, (int)d, c.freq_dim, 1, st));
  ALG_TRY_EW(dit::gemv_f32(e->te2_w, e->te2_b, te_h, te_o, (int)d, (int)d, 0, st));
  ALG_TRY_EW(dit::temb_finish(te_o, temb, silu_temb, (int)d, st));
  ALG_TRY(gemm(st, silu_temb, d, e->tp_w, d, tproj, 6 * d, 1, 6 * d, d, e->tp_b));
  for (int p = 0; p < n_pass && !ctx_hit; ++p) {
    int same = -1;
    for (int r = 0; r < p; ++r)
      if (text[r] == text[p]) same = r;
    bf16* dst = ctx_text + (int64_t)p * txt * d;
    if (same >= 0) {  // [neg, neg, pos]: embed each distinct prompt once
      ALG_TRY_EW(dit::copy_rows(ctx_text + (int64_t)same * txt * d, d, dst, d, txt, (int)d, st));
    } else {
      ALG_TRY(gemm(st, text[p], c.text_dim, e->tx1_w, c.text_dim, txt_h, d, txt, d, c.text_dim, e->tx1_b, ALG_EPI_GELU_TANH));
      ALG_TRY(gemm(st, txt_h, d, e->tx2_w, d, dst, d, txt, d, d, e->tx2_b));
    }
  }
  if (n_img > 0 && !ctx_hit) {
    ALG_TRY_EW(layer_norm(st, (const bf16*)image, img_n, ni, c.image_dim, 1e-5f, e->in1_w, e->in1_b, nullptr, nullptr));
    ALG_TRY(gemm(st, img_n, c.image_dim, e->if1_w, c.image_dim, img_h, c.image_dim, ni, c.image_dim, c.image_dim,
                 e->if1_b, ALG_EPI_GELU_ERF));
    ALG_TRY(gemm(st, img_h, c.image_dim, e->if2_w, c.image_dim, img_p, d, ni, d, c.image_dim, e->if2_b));
    ALG_TRY_EW(layer_norm(st, img_p, ctx_img, ni, (int)d, 1e-5f, e->in2_w, e->in2_b, nullptr, nullptr));
    for (int p = 1; p < n_pass; ++p)
      ALG_TRY_EW(dit::copy_rows(ctx_img, d, ctx_img + (int64_t)p * ni * d, d, ni, (int)d, st));
  }
  debug_dump(temb, (size_t)d * 2);
  debug_dump(tproj, (size_t)6 * d * 2);

  dit::RopeTables rope{e->rope_t, e->rope_h, e->rope_w, e->n_t, e->n_h, e->n_w, T, pph, ppw};

  // ---- 3. transformer blocks --------------------------------------------------------------------------
  for (int l = 0; l < c.num_layers; ++l) {
    const BlockW& b = e->blocks[l];
    if (slot) {  // this layer's context K / V^T live in the cache slot instead of the workspace
      bf16* base = reinterpret_cast<bf16*>(slot->buf) + (size_t)l * layer_elems;
      k_text = base;
      vt_text = k_text + al(kt_elems);
      k_img = vt_text + al(vtt_elems);
      vt_img = k_img + al(ki_elems);
    }
    ALG_TRY_EW(dit::add_table(b.table, tproj, mod, 6, (int)d, 0, st));
    const float *shift = mod, *scale = mod + d, *gate = mod + 2 * d, *c_shift = mod + 3 * d, *c_scale = mod + 4 * d,
                *c_gate = mod + 5 * d;
    // self-attention
    ALG_TRY_EW(layer_norm(st, x, h, M, (int)d, c.eps, nullptr, nullptr, scale, shift));
    ALG_TRY(gemm(st, h, d, b.attn1.q_w, d, q, d, M, d, d, b.attn1.q_b));
    ALG_TRY(gemm(st, h, d, b.attn1.k_w, d, k, d, M, d, d, b.attn1.k_b));
    for (int p = 0; p < n_pass; ++p)  // V^T = W_v h^T + b_v: swapped operands, bias per row
      ALG_TRY(gemm(st, b.attn1.v_w, d, h + (int64_t)p * N * d, d, vt + (int64_t)p * d * Npad, Npad, d, N, d, b.attn1.v_b,
                   ALG_EPI_NONE, nullptr, nullptr, 1));
    ALG_TRY_EW(dit::rms_norm_rope(q, M, (int)d, hd, c.eps, (const bf16*)b.attn1.norm_q, &rope, st));
    ALG_TRY_EW(dit::rms_norm_rope(k, M, (int)d, hd, c.eps, (const bf16*)b.attn1.norm_k, &rope, st));
    ALG_TRY(attention(st, q, k, vt, ao, n_pass, heads, hd, N, N, Npad, 0));
    ALG_TRY(gemm(st, ao, d, b.attn1.o_w, d, x, d, M, d, d, b.attn1.o_b, ALG_EPI_GATE_RESIDUAL, x, gate));
    // cross-attention (text keys = last text_len context tokens, image keys = the rest)
    ALG_TRY_EW(layer_norm(st, x, h, M, (int)d, c.eps, b.norm2_w, b.norm2_b, nullptr, nullptr));
    ALG_TRY(gemm(st, h, d, b.attn2.q_w, d, q, d, M, d, d, b.attn2.q_b));
    ALG_TRY_EW(dit::rms_norm_rope(q, M, (int)d, hd, c.eps, (const bf16*)b.attn2.norm_q, nullptr, st));
    if (!ctx_hit) {
      ALG_TRY(gemm(st, ctx_text, d, b.attn2.k_w, d, k_text, d, n_pass * txt, d, d, b.attn2.k_b));
      ALG_TRY_EW(dit::rms_norm_rope(k_text, n_pass * txt, (int)d, hd, c.eps, (const bf16*)b.attn2.norm_k, nullptr, st));
      for (int p = 0; p < n_pass; ++p)
        ALG_TRY(gemm(st, b.attn2.v_w, d, ctx_text + (int64_t)p * txt * d, d, vt_text + (int64_t)p * d * txt_pad, txt_pad, d,
                     txt, d, b.attn2.v_b, ALG_EPI_NONE, nullptr, nullptr, 1));
    }
    if (n_img > 0) {
      if (!ctx_hit) {
        ALG_TRY(gemm(st, ctx_img, d, b.attn2.add_k_w, d, k_img, d, n_pass * ni, d, d, b.attn2.add_k_b));
        ALG_TRY_EW(dit::rms_norm_rope(k_img, n_pass * ni, (int)d, hd, c.eps, (const bf16*)b.attn2.norm_added_k, nullptr, st));
        for (int p = 0; p < n_pass; ++p)
          ALG_TRY(gemm(st, b.attn2.add_v_w, d, ctx_img + (int64_t)p * ni * d, d, vt_img + (int64_t)p * d * ni_pad, ni_pad, d,
                       ni, d, b.attn2.add_v_b, ALG_EPI_NONE, nullptr, nullptr, 1));
      }
      ALG_TRY(attention(st, q, k_img, vt_img, ao, n_pass, heads, hd, N, ni, ni_pad, 0));
    }
    ALG_TRY(attention(st, q, k_text, vt_text, ao, n_pass, heads, hd, N, txt, txt_pad, n_img > 0 ? 1 : 0));
    ALG_TRY(gemm(st, ao, d, b.attn2.o_w, d, x, d, M, d, d, b.attn2.o_b, ALG_EPI_RESIDUAL, x));
    // feed-forward
    ALG_TRY_EW(layer_norm(st, x, h, M, (int)d, c.eps, nullptr, nullptr, c_scale, c_shift));
    ALG_TRY(gemm(st, h, d, b.ffn1_w, d, ffn, c.ffn_dim, M, c.ffn_dim, d, b.ffn1_b, ALG_EPI_GELU_TANH));
    ALG_TRY(gemm(st, ffn, c.ffn_dim, b.ffn2_w, c.ffn_dim, x, d, M, d, c.ffn_dim, b.ffn2_b, ALG_EPI_GATE_RESIDUAL, x, c_gate));
    debug_dump(x, (size_t)M * d * 2);
  }

  if (slot && !ctx_hit) slot->valid = true;  // every layer's context K / V^T is in the slot now

  // ---- 4. output norm, projection, unpatchify ------------------------------------------------------------
  ALG_TRY_EW(dit::add_table(e->head_table, temb, mod, 2, (int)d, 1, st));
  ALG_TRY_EW(layer_norm(st, x, h, M, (int)d, c.eps, nullptr, nullptr, mod + d, mod));
  ALG_TRY(gemm(st, h, d, e->proj_w, d, proj, pop, M, po, d, e->proj_b));
  ALG_TRY_EW(dit::unpatchify(proj, (bf16*)noise_out, n_pass, c.out_channels, T, H, W, st));
  return 0;
}

extern "C" int alg_wan_context_cache(alg_wan_engine_t* e, int enable) {
  using namespace alg;
  ALG_REQUIRE(e, "wan_context_cache: null engine");
  e->ctx_caching = enable != 0;
  for (auto& s : e->ctx_slots) s.valid = false;  // enabling, disabling and re-enabling all drop what was memoised
  return 0;
}

extern "C" int alg_wan_profile(alg_wan_engine_t* e, int enable) {
  using namespace alg;
  ALG_REQUIRE(e, "wan_profile: null engine");
  e->profiling = enable != 0;
  e->spans.clear();
  e->ev_used = 0;
  return 0;
}

extern "C" int alg_wan_profile_read(alg_wan_engine_t* e, float* ms_per_class, int32_t* launches_per_class, int n_classes) {
  using namespace alg;
  ALG_REQUIRE(e && ms_per_class && launches_per_class && n_classes >= PROF_NUM, "wan_profile_read: bad arguments");
  for (int i = 0; i < n_classes; ++i) {
    ms_per_class[i] = 0.f;
    launches_per_class[i] = 0;
  }
  for (const auto& sp : e->spans) {
    ALG_CUDA_OK(cudaEventSynchronize(sp.b));
    float ms = 0.f;
    ALG_CUDA_OK(cudaEventElapsedTime(&ms, sp.a, sp.b));
    ms_per_class[sp.cls] += ms;
    launches_per_class[sp.cls] += 1;
  }
  e->spans.clear();
  e->ev_used = 0;
  return 0;
}
