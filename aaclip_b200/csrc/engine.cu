// Context, weight packing and the forward schedules behind the C ABI (include/aaclip_b200.h).
//
//   aaclip_visual_forward  == AdaptedCLIP.forward            (model/adapter.py:67-112)
//   aaclip_forward_fused   == AdaptedCLIP.forward + 4x calculate_similarity_map + level sum + image score
//                             (test.py:80-93), seg tokens never materialised
//   aaclip_text_forward    == AdaptedCLIP.encode_text(adapt_text=True)   (model/adapter.py:114-145)
//
// Data layout in HBM (per context, sized for cfg.max_batch images; larger batches are processed in chunks):
//   x     fp32 [B*L, width]      residual stream, token-major (row = b*L + l, l = 0 is the class token)
//   xn    bf16 [B*L, width]      LayerNorm output / bf16 copy of x: the A operand of the next GEMM
//   qkv   bf16 [B*L, 3*width]    fused in_proj output, read by the attention kernel through TMA
//   att   bf16 [B*L, width]      attention output (heads concatenated)
//   h     bf16 [B*L, mlp_width]  GELU(c_fc) output
//   a     fp32 [B*L, width]      adapter branch before the norm-matching mix
//   tap   bf16 [B*P, width]      ln_post(x[:, 1:]) of the current level
//   s     fp32 [B*P, 2*E]        seg (and, on the last level, det) projection before normalisation
//   weights: GEMM operands bf16 [out, in] (nn.Linear layout, K contiguous); LN params, biases, embeddings fp32.
#include <limits.h>
#include <stdarg.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace {

typedef __nv_bfloat16 bf16;

__global__ void cast_pad_kernel(const float* __restrict__ src, int rows, int cols, bf16* __restrict__ dst, int ld) {
  const long long total = (long long)rows * ld;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = int(t % ld);
    const long long r = t / ld;
    dst[t] = __float2bfloat16(c < cols ? src[r * cols + c] : 0.f);
  }
}

// x[n*ctx + t, :] = token_embedding[tokens[n, t]] + positional_embedding[t]     (model/adapter.py:118-122)
__global__ void text_embed_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ emb,
                                  const float* __restrict__ pos, int ctx, int width, int vocab, float* __restrict__ x) {
  const int row = blockIdx.x;
  const int t = row % ctx;
  int tok = tokens[row];
  tok = min(max(tok, 0), vocab - 1);
  const float4* e = reinterpret_cast<const float4*>(emb + (size_t)tok * width);
  const float4* p = reinterpret_cast<const float4*>(pos + (size_t)t * width);
  float4* o = reinterpret_cast<float4*>(x + (size_t)row * width);
  for (int c = threadIdx.x; c < width / 4; c += blockDim.x) {
    const float4 a = e[c], b = p[c];
    o[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// gathered[n, :] = x[n*ctx + argmax_t tokens[n, t], :]   (EOT token = highest id, first occurrence; adapter.py:140)
__global__ void eot_gather_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ x, int ctx, int width,
                                  float* __restrict__ gathered) {
  const int n = blockIdx.x;
  __shared__ int best_t;
  if (threadIdx.x < 32) {
    int bv = INT_MIN, bt = 0;
    for (int t = threadIdx.x; t < ctx; t += 32) {
      const int v = tokens[n * ctx + t];
      if (v > bv) { bv = v; bt = t; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const int ov = __shfl_xor_sync(0xffffffffu, bv, o), ot = __shfl_xor_sync(0xffffffffu, bt, o);
      if (ov > bv || (ov == bv && ot < bt)) { bv = ov; bt = ot; }
    }
    if (threadIdx.x == 0) best_t = bt;
  }
  __syncthreads();
  const float* src = x + ((size_t)n * ctx + best_t) * width;
  for (int c = threadIdx.x; c < width; c += blockDim.x) gathered[(size_t)n * width + c] = src[c];
}

struct LayerW {
  // folded-LayerNorm schedule (visual tower): fp32 masters of the two Linears that consume a LayerNorm, their
  // gamma-folded bf16 form, the column sums of the folded weights and the beta-folded biases
  float *qkv_w32 = nullptr, *fc_w32 = nullptr, *qkv_cs = nullptr, *qkv_bf = nullptr, *fc_cs = nullptr, *fc_bf = nullptr;
  bf16 *qkv_wf = nullptr, *fc_wf = nullptr;
  bf16 *qkv_w = nullptr, *out_w = nullptr, *fc_w = nullptr, *proj_w = nullptr;
  float *ln1_g = nullptr, *ln1_b = nullptr, *qkv_b = nullptr, *out_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr,
        *fc_b = nullptr, *proj_b = nullptr;
};

struct Tower {
  int width = 0, heads = 0, layers = 0, mlp = 0;
  std::vector<LayerW> lw;
  std::vector<bf16*> adapters;
};

}  // namespace

// kernel classes for the optional per-launch CUDA-event profile (aaclip_profile_*)
enum ProfClass : int {
  PC_GEMM_QKV = 0, PC_GEMM_OUT, PC_GEMM_FC, PC_GEMM_PROJ, PC_GEMM_ADAPTER, PC_GEMM_SEGDET, PC_GEMM_PATCH,
  PC_ATTENTION, PC_LAYERNORM, PC_ADAPTER_MIX, PC_CAST, PC_L2NORM, PC_DET_MEAN, PC_STEM_MISC, PC_HEAD_MAPS,
  PC_OTHER, PC_COUNT
};
struct ProfRec { int cls; cudaEvent_t a, b; };

struct aaclip_ctx {
  bool prof_on = false;
  double prof_span_ms = 0.0;   // first launch start -> last launch end of the log read last
  std::vector<ProfRec> prof;
  aaclip_cfg cfg;
  int device = 0;
  int G = 0, P = 0, L = 0, Kpad = 0, E = 0;
  int cta_group = 0;   // 0: tile shape chosen per launch (k::launch_gemm); 1 / 2 / 3 force one
  // LayerNorm folded into the consumer GEMMs of the visual tower (see gemm_sm100.cuh); AACLIP_LN_FOLD=0 disables it
  bool ln_fold = true, fold_dirty = true;
  // first block (0-based) of the visual tower whose attention is the batch-coupled v-v form of the surgery
  // extractor (aaclip_dapm_replace; model/transformer.py:406-425); >= layers: none
  int vv_first = INT_MAX;
  float2* part = nullptr;   // [rows][width / 128] (sum, sum of squares) per 128-column slice of the fp32 rows
  // CUDA graphs of the fused forward (aaclip_forward_fused), one per (batch, pointers, mode): the ~100 launches of a
  // chunk become one graph launch.  A key is run eagerly the first time, captured the second, replayed from then on.
  struct FusedKey {
    const float* image; const float* anchors; float* maps; float* scores; float* minmax; int B; int mode;
    bool operator==(const FusedKey& o) const {
      return image == o.image && anchors == o.anchors && maps == o.maps && scores == o.scores && minmax == o.minmax &&
             B == o.B && mode == o.mode;
    }
  };
  struct FusedGraph { FusedKey key; cudaGraphExec_t exec; long long launches; unsigned long long last_use; };
  std::vector<FusedGraph> graphs;
  std::vector<FusedKey> seen_once;
  unsigned long long graph_clock = 0;
  bool use_graphs = true;
  long long bytes = 0;
  long long launches = 0;
  std::vector<void*> allocs;
  // visual
  Tower v;
  bf16* conv_w = nullptr;
  float *cls = nullptr, *pos = nullptr, *ln_pre_g = nullptr, *ln_pre_b = nullptr, *ln_post_g = nullptr,
        *ln_post_b = nullptr;
  std::vector<bf16*> segdet_w;  // per level [E, width]; last level [2E, width] (det_proj appended)
  // text
  Tower t;
  float *tok_emb = nullptr, *t_pos = nullptr, *ln_final_g = nullptr, *ln_final_b = nullptr;
  bf16* t_final = nullptr;
  int t_final_act = gemm::ACT_LEAKY;   // aaclip_set_text_final: LEAKY = text_adapter[-1], NONE = CLIP's text_projection
  // workspaces
  int cap_rows = 0;  // rows of the token-major buffers
  float *x = nullptr, *a = nullptr, *s = nullptr, *dots = nullptr, *det = nullptr, *stage = nullptr, *rownorm = nullptr,
        *partials = nullptr;   // [levels][B*P][E/128] float4 partial sums of the fused seg_proj epilogue
  bf16 *xn = nullptr, *qkv = nullptr, *att = nullptr, *h = nullptr, *col = nullptr, *tap = nullptr;
  // host-buffer pipeline (aaclip_submit_host / aaclip_wait_host): two slots of device staging, copy-in, compute
  // and copy-out streams, so the H2D of batch k+1 and the D2H of batch k-1 overlap the compute of batch k
  struct HostSlot {
    float *img = nullptr, *maps = nullptr, *scores = nullptr, *anchors = nullptr, *minmax = nullptr;
    uint8_t *raw = nullptr, *raw_scratch = nullptr;   // aaclip_submit_host_u8: raw images and the resample scratch
    long long raw_cap = 0, raw_scratch_cap = 0;
    cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr;
    bool busy = false;
    long long ticket = -1;
  } slots[2];
  long long next_ticket = 0;
  long long stage_cap = 0;
  cudaStream_t own_stream = nullptr, in_stream = nullptr, out_stream = nullptr;

  template <typename T>
  int alloc(T** p, long long n) {
    void* q = nullptr;
    const long long nb = std::max<long long>(n, 1) * (long long)sizeof(T);
    AACLIP_CUDA_CHECK(cudaMalloc(&q, nb));
    AACLIP_CUDA_CHECK(cudaMemset(q, 0, nb));
    allocs.push_back(q);
    bytes += nb;
    *p = static_cast<T*>(q);
    return host::OK;
  }
};

namespace {

// Every entry point runs on its context's device and puts the caller's current device back on return (the thread's
// current device belongs to the caller - PyTorch - not to this library).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ENTER_DEVICE(c)                                                                            \
  DeviceGuard _guard((c)->device);                                                                 \
  if (!_guard.ok) return host::fail(host::ERR_CUDA, "cudaSetDevice(%d) failed", (c)->device)

#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

inline void prof_begin(aaclip_ctx* c, int cls, cudaStream_t st) {
  if (!c->prof_on) return;
  ProfRec r; r.cls = cls;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  c->prof.push_back(r);
}
inline void prof_end(aaclip_ctx* c, cudaStream_t st) {
  if (c->prof_on) cudaEventRecord(c->prof.back().b, st);
}
// launch one kernel of class CLS: count it, and bracket it with events when profiling is on
#define RUN(CLS, expr)             \
  do {                             \
    prof_begin(c, CLS, st);        \
    int _rc = (expr);              \
    if (_rc) return _rc;           \
    prof_end(c, st);               \
    c->launches++;                 \
  } while (0)

int alloc_tower(aaclip_ctx* c, Tower& t, int n_adapters) {
  const long long w = t.width, ff = t.mlp;
  t.lw.resize(t.layers);
  for (auto& l : t.lw) {
    TRY(c->alloc(&l.qkv_w, 3 * w * w)); TRY(c->alloc(&l.out_w, w * w));
    TRY(c->alloc(&l.fc_w, ff * w)); TRY(c->alloc(&l.proj_w, w * ff));
    TRY(c->alloc(&l.ln1_g, w)); TRY(c->alloc(&l.ln1_b, w)); TRY(c->alloc(&l.qkv_b, 3 * w));
    TRY(c->alloc(&l.out_b, w)); TRY(c->alloc(&l.ln2_g, w)); TRY(c->alloc(&l.ln2_b, w));
    TRY(c->alloc(&l.fc_b, ff)); TRY(c->alloc(&l.proj_b, w));
  }
  t.adapters.resize(n_adapters);
  for (auto& p : t.adapters) TRY(c->alloc(&p, w * w));
  return host::OK;
}

// One transformer block (+ optional adapter mix) over `rows` token rows of width t.width.
// xn_ready: xn already holds ln_1(x) (written by the previous block's fused adapter mix).
int run_block(aaclip_ctx* c, const Tower& t, int i, int B, int L, int causal, float adapt_w, bool* xn_ready,
              cudaStream_t st, bool vv = false) {
  const LayerW& l = t.lw[i];
  const int rows = B * L, w = t.width, ff = t.mlp;
  const int cg = c->cta_group;
  const float eps = 1e-5f;
  if (!*xn_ready) {
    RUN(PC_LAYERNORM, k::launch_layernorm(c->x, l.ln1_g, l.ln1_b, eps, rows, w, 0, 0, 0, c->xn, nullptr, st));
  }
  *xn_ready = false;
  if (vv) {
    // surgery block: only the value third of in_proj is needed (q and k feed the discarded `attn_ori`)
    RUN(PC_GEMM_QKV, k::launch_gemm(c->xn, w, l.qkv_w + 2LL * w * w, w, rows, w, w, l.qkv_b + 2 * w, c->qkv, w, gemm::ACT_NONE,
                       gemm::OUT_BF16, nullptr, 0, cg, st));
    RUN(PC_ATTENTION, k::launch_vv_attention(c->qkv, w, c->att, w, B, L, t.heads, st));
  } else {
    RUN(PC_GEMM_QKV, k::launch_gemm(c->xn, w, l.qkv_w, w, rows, 3 * w, w, l.qkv_b, c->qkv, 3 * w, gemm::ACT_NONE, gemm::OUT_BF16,
                       nullptr, 0, cg, st));
    RUN(PC_ATTENTION, k::launch_attention(c->qkv, c->att, B, L, t.heads, causal, st));
  }
  RUN(PC_GEMM_OUT, k::launch_gemm(c->att, w, l.out_w, w, rows, w, w, l.out_b, c->x, w, gemm::ACT_NONE, gemm::OUT_F32_RESID,
                     nullptr, 0, cg, st));
  RUN(PC_LAYERNORM, k::launch_layernorm(c->x, l.ln2_g, l.ln2_b, eps, rows, w, 0, 0, 0, c->xn, nullptr, st));
  RUN(PC_GEMM_FC, k::launch_gemm(c->xn, w, l.fc_w, w, rows, ff, w, l.fc_b, c->h, ff, c->cfg.act, gemm::OUT_BF16, nullptr, 0, cg,
                     st));
  RUN(PC_GEMM_PROJ, k::launch_gemm(c->h, ff, l.proj_w, ff, rows, w, ff, l.proj_b, c->x, w, gemm::ACT_NONE, gemm::OUT_F32_RESID,
                     nullptr, 0, cg, st));
  if (i < (int)t.adapters.size()) {
    // adapter branch (model/adapter.py:92-99): a = LeakyReLU(x W_a^T); x <- w a |x|/|a| + (1-w) x
    RUN(PC_CAST, k::launch_cast_bf16(c->x, c->xn, (long long)rows * w, st));
    RUN(PC_GEMM_ADAPTER, k::launch_gemm(c->xn, w, t.adapters[i], w, rows, w, w, nullptr, c->a, w, gemm::ACT_LEAKY, gemm::OUT_F32,
                       nullptr, 0, cg, st));
    const bool fuse_ln = (i + 1 < t.layers);
    RUN(PC_ADAPTER_MIX, k::launch_adapter_mix(c->x, c->a, adapt_w, rows, w, fuse_ln ? t.lw[i + 1].ln1_g : nullptr,
                              fuse_ln ? t.lw[i + 1].ln1_b : nullptr, eps, fuse_ln ? c->xn : nullptr, st));
    *xn_ready = fuse_ln;
  }
  return host::OK;
}

// (Re)build the folded weights of the visual tower after any of their sources changed.
int refold(aaclip_ctx* c, cudaStream_t st) {
  if (!c->ln_fold || !c->fold_dirty) return host::OK;
  const int w = c->v.width, ff = c->v.mlp;
  for (auto& l : c->v.lw) {
    TRY(k::launch_fold_ln_weight(l.qkv_w32, l.qkv_b, l.ln1_g, l.ln1_b, 3 * w, w, l.qkv_wf, l.qkv_cs, l.qkv_bf, st));
    TRY(k::launch_fold_ln_weight(l.fc_w32, l.fc_b, l.ln2_g, l.ln2_b, ff, w, l.fc_wf, l.fc_cs, l.fc_bf, st));
    c->launches += 2;
  }
  c->fold_dirty = false;
  return host::OK;
}

// One block of the visual tower on the folded-LayerNorm schedule: on entry xn holds the bf16 copy of x and `part`
// its per-slice (sum, sum of squares); both are left in that state for the next block.  No LayerNorm launches:
//   qkv  = rstd (xb Wf_qkv^T - mean s) + b'        (consumer epilogue)
//   x   += att W_out^T + b, xb, part               (producer epilogue)
//   h    = GELU(rstd (xb Wf_fc^T - mean s) + b')
//   x   += h W_proj^T + b, xb, part
//   adapter layers: a = LeakyReLU(xb W_a^T) straight from the bf16 copy (no cast), then the mix rewrites x, xb, part
int run_block_fold(aaclip_ctx* c, const Tower& t, int i, int B, int L, float adapt_w, cudaStream_t st) {
  const LayerW& l = t.lw[i];
  const int rows = B * L, w = t.width, ff = t.mlp, slices = w / 128;
  const int cg = c->cta_group;
  k::LnFold cons;
  cons.part_in = c->part; cons.slices = slices; cons.eps = 1e-5f;
  k::LnFold prod;
  prod.xb = c->xn; prod.ldxb = w; prod.part_out = c->part;
  if (i >= c->vv_first) {
    // surgery block (model/transformer.py:123-152): value third of the folded in_proj, then the batch-coupled v-v attention
    cons.colsum = l.qkv_cs + 2 * w;
    RUN(PC_GEMM_QKV, k::launch_gemm(c->xn, w, l.qkv_wf + 2LL * w * w, w, rows, w, w, l.qkv_bf + 2 * w, c->qkv, w, gemm::ACT_NONE,
                       gemm::OUT_BF16, nullptr, 0, cg, st, nullptr, nullptr, 0, &cons));
    RUN(PC_ATTENTION, k::launch_vv_attention(c->qkv, w, c->att, w, B, L, t.heads, st));
  } else {
    cons.colsum = l.qkv_cs;
    RUN(PC_GEMM_QKV, k::launch_gemm(c->xn, w, l.qkv_wf, w, rows, 3 * w, w, l.qkv_bf, c->qkv, 3 * w, gemm::ACT_NONE,
                       gemm::OUT_BF16, nullptr, 0, cg, st, nullptr, nullptr, 0, &cons));
    RUN(PC_ATTENTION, k::launch_attention(c->qkv, c->att, B, L, t.heads, 0, st));
  }
  RUN(PC_GEMM_OUT, k::launch_gemm(c->att, w, l.out_w, w, rows, w, w, l.out_b, c->x, w, gemm::ACT_NONE, gemm::OUT_F32_RESID_LN,
                     nullptr, 0, cg, st, nullptr, nullptr, 0, &prod));
  cons.colsum = l.fc_cs;
  RUN(PC_GEMM_FC, k::launch_gemm(c->xn, w, l.fc_wf, w, rows, ff, w, l.fc_bf, c->h, ff, c->cfg.act, gemm::OUT_BF16, nullptr, 0, cg,
                     st, nullptr, nullptr, 0, &cons));
  RUN(PC_GEMM_PROJ, k::launch_gemm(c->h, ff, l.proj_w, ff, rows, w, ff, l.proj_b, c->x, w, gemm::ACT_NONE, gemm::OUT_F32_RESID_LN,
                     nullptr, 0, cg, st, nullptr, nullptr, 0, &prod));
  if (i < (int)t.adapters.size()) {
    // the branch a is kept in bf16: it enters x with weight adapt_w = 0.1, a tenth of the rounding xb itself carries
    RUN(PC_GEMM_ADAPTER, k::launch_gemm(c->xn, w, t.adapters[i], w, rows, w, w, nullptr, c->a, w, gemm::ACT_LEAKY, gemm::OUT_BF16,
                       nullptr, 0, cg, st));
    RUN(PC_ADAPTER_MIX, k::launch_adapter_mix(c->x, c->a, adapt_w, rows, w, nullptr, nullptr, 1e-5f, nullptr, st, c->xn, c->part,
                              slices, true));
  }
  return host::OK;
}

// seg_out[level] (fp32 [B,P,E], optional), det_out (fp32 [B,E], optional), dots (optional, [levels][B*P][2] with
// anchors) for one chunk of B <= max_batch images.
// raw_out[level] (fp32 [B,L,width], optional): the residual stream after block levels[level] (Transformer.forward's
// out_tokens, model/transformer.py:296-318); pooled_out (fp32 [B,E], optional): ln_post(class token) @ proj, L2-normalised
// when pooled_normalize (VisionTransformer.forward :542-546, CLIP.encode_image model/model.py:185-188).
int visual_chunk(aaclip_ctx* c, const float* image, int B, void* const* seg_out, int seg_is_bf16, long long seg_off,
                 float* det_out, const float* anchors, float* dots, cudaStream_t st, float* const* raw_out = nullptr,
                 long long raw_off = 0, float* pooled_out = nullptr, int pooled_normalize = 0) {
  const aaclip_cfg& cfg = c->cfg;
  const int w = cfg.width, L = c->L, P = c->P, E = c->E, rows = B * L, prow = B * P;
  const int cg = c->cta_group;
  // stem: conv1 as im2col GEMM, +pos, class token, ln_pre (model/adapter.py:68-85)
  RUN(PC_STEM_MISC, k::launch_im2col(image, B, cfg.image_size, cfg.patch_size, c->Kpad, c->col, st));
  RUN(PC_GEMM_PATCH, k::launch_gemm(c->col, c->Kpad, c->conv_w, c->Kpad, prow, w, c->Kpad, nullptr, c->x, w, gemm::ACT_NONE,
                     gemm::OUT_F32_PATCH, c->pos, P, cg, st));
  RUN(PC_STEM_MISC, k::launch_cls_rows(c->x, c->cls, c->pos, B, L, w, st));
  const bool fold = c->ln_fold;
  if (fold) {
    // ln_pre also leaves the bf16 copy of its output and the row sums the first in_proj epilogue needs
    TRY(refold(c, st));
    RUN(PC_LAYERNORM, k::launch_layernorm(c->x, c->ln_pre_g, c->ln_pre_b, 1e-5f, rows, w, 0, 0, 0, c->xn, c->x, st, c->part,
                            w / 128));
  } else {
    RUN(PC_LAYERNORM, k::launch_layernorm(c->x, c->ln_pre_g, c->ln_pre_b, 1e-5f, rows, w, 0, 0, 0, nullptr, c->x, st));
  }
  bool xn_ready = false;
  int level = 0;
  for (int i = 0; i < cfg.layers; ++i) {
    if (fold) TRY(run_block_fold(c, c->v, i, B, L, cfg.image_adapt_weight, st));
    else TRY(run_block(c, c->v, i, B, L, 0, cfg.image_adapt_weight, &xn_ready, st, i >= c->vv_first));
    if (level < cfg.n_levels && cfg.levels[level] == i + 1) {
      const bool last = (level == cfg.n_levels - 1);
      if (raw_out && raw_out[level])
        AACLIP_CUDA_CHECK(cudaMemcpyAsync(raw_out[level] + raw_off, c->x, (size_t)rows * w * sizeof(float),
                                          cudaMemcpyDeviceToDevice, st));
      const bool want_det = last && (det_out != nullptr);
      const int n_out = want_det ? 2 * E : E;
      // seg tokens leave as fp32 (the reference's dtype) or bf16 (half the bytes for the head to stream)
      void* so = (seg_out && seg_out[level])
                     ? static_cast<void*>(static_cast<uint8_t*>(seg_out[level]) + seg_off * (seg_is_bf16 ? 2 : 4)) : nullptr;
      if (!so && !dots && !want_det) { ++level; continue; }   // encode_image: nothing is projected at this level
      // tap: x[:, 1:, :] -> ln_post -> seg_proj (and det_proj on the last tap)   (model/adapter.py:100-111)
      RUN(PC_LAYERNORM, k::launch_layernorm(c->x, c->ln_post_g, c->ln_post_b, 1e-5f, prow, w, P, 1, (long long)L * w, c->tap,
                              nullptr, st));
      const int act = cfg.proj_relu ? gemm::ACT_LEAKY : gemm::ACT_NONE;
      if (dots && !so && E % 128 == 0) {
        // fused path: the normalised seg tokens are never materialised - the GEMM epilogue leaves the partial sums
        // of ||f||^2, <f,T0>, <f,T1> per 128-column slice (det_proj columns on the last level are still stored)
        RUN(PC_GEMM_SEGDET, k::launch_gemm(c->tap, w, c->segdet_w[level], w, prow, n_out, w, nullptr, c->s, 2 * E, act,
                           gemm::OUT_DOTS, nullptr, 0, cg, st, anchors,
                           c->partials + (size_t)level * prow * (E / 128) * 4, E));
      } else {
        RUN(PC_GEMM_SEGDET, k::launch_gemm(c->tap, w, c->segdet_w[level], w, prow, n_out, w, nullptr, c->s, 2 * E, act,
                           gemm::OUT_F32, nullptr, 0, cg, st));
        float* dl = dots ? dots + (size_t)level * prow * 2 : nullptr;
        if (so || dl) {
          RUN(PC_L2NORM, k::launch_l2norm_rows(c->s, 2 * E, 0, prow, E, seg_is_bf16 ? nullptr : static_cast<float*>(so),
                                               seg_is_bf16 ? so : nullptr, dl ? anchors : nullptr, dl, st));
        }
      }
      if (want_det) { RUN(PC_DET_MEAN, k::launch_det_mean(c->s, 2 * E, E, B, P, E, c->rownorm, det_out, st)); c->launches++; }
      ++level;
    }
  }
  if (dots && !seg_out && E % 128 == 0) {
    RUN(PC_L2NORM, k::launch_dots_finish(c->partials, cfg.n_levels, prow, E / 128, dots, st));
  }
  if (pooled_out) {
    // pooled = ln_post(x[:, 0]) @ proj: the class token of every image through the projection the last seg_proj slot
    // holds (a context built for CLIP.encode_image carries visual.proj^T there)
    RUN(PC_LAYERNORM, k::launch_layernorm(c->x, c->ln_post_g, c->ln_post_b, 1e-5f, B, w, 1, 0, (long long)L * w, c->tap, nullptr,
                                          st));
    RUN(PC_GEMM_SEGDET, k::launch_gemm(c->tap, w, c->segdet_w[cfg.n_levels - 1], w, B, E, w, nullptr, c->s, 2 * E, gemm::ACT_NONE,
                                       gemm::OUT_F32, nullptr, 0, cg, st));
    if (pooled_normalize) {
      RUN(PC_L2NORM, k::launch_l2norm_rows(c->s, 2 * E, 0, B, E, pooled_out, nullptr, nullptr, nullptr, st));
    } else {
      AACLIP_CUDA_CHECK(cudaMemcpy2DAsync(pooled_out, (size_t)E * sizeof(float), c->s, (size_t)2 * E * sizeof(float),
                                          (size_t)E * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    }
  }
  return host::OK;
}

int check_ready(const aaclip_ctx* c) {
  if (!c) return host::fail(host::ERR_INVALID, "null context");
  return host::OK;
}

// The v-v attention couples the images of a batch (vv_attn.cu): such a context never splits a batch into chunks.
int check_vv_batch(const aaclip_ctx* c, int B, const char* who) {
  if (c->vv_first >= c->cfg.layers) return host::OK;
  const int cap = std::min(c->cfg.max_batch, k::vv_attention_max_batch());
  if (B > cap)
    return host::fail(host::ERR_INVALID, "%s: batch %d on a context with v-v (surgery) attention, which couples the images "
                                         "of a batch: at most min(max_batch, %d) = %d images per call", who, B,
                      k::vv_attention_max_batch(), cap);
  return host::OK;
}

}  // namespace

extern "C" const char* aaclip_last_error(void) { return host::last_error().c_str(); }
extern "C" int aaclip_abi_version(void) { return 2; }

extern "C" int aaclip_create(aaclip_ctx** out, const aaclip_cfg* cfg, int device) {
  if (!out || !cfg) return host::fail(host::ERR_INVALID, "aaclip_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return host::fail(host::ERR_NO_DEVICE, "no CUDA device: aaclip_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return host::fail(host::ERR_INVALID, "device %d out of range", device);
  cudaDeviceProp prop;
  AACLIP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return host::fail(host::ERR_NO_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major,
                      prop.minor);
  if (cfg->width <= 0 || cfg->heads <= 0 || cfg->width != cfg->heads * 64)
    return host::fail(host::ERR_INVALID, "width %d must equal heads %d * 64", cfg->width, cfg->heads);
  if (cfg->width % 256 != 0 || cfg->mlp_width % 256 != 0 || cfg->embed_dim % 256 != 0)
    return host::fail(host::ERR_INVALID, "width / mlp_width / embed_dim must be multiples of 256");
  if (cfg->patch_size <= 0 || cfg->image_size % cfg->patch_size != 0)
    return host::fail(host::ERR_INVALID, "image_size %d not divisible by patch_size %d", cfg->image_size, cfg->patch_size);
  if (cfg->n_levels < 1 || cfg->n_levels > 8) return host::fail(host::ERR_INVALID, "n_levels %d", cfg->n_levels);
  for (int i = 0; i < cfg->n_levels; ++i)
    if (cfg->levels[i] < 1 || cfg->levels[i] > cfg->layers || (i > 0 && cfg->levels[i] <= cfg->levels[i - 1]))
      return host::fail(host::ERR_INVALID, "levels must be strictly ascending within [1, layers]");
  if (cfg->image_adapt_until < 0 || cfg->image_adapt_until > cfg->layers)
    return host::fail(host::ERR_INVALID, "image_adapt_until %d", cfg->image_adapt_until);
  if (cfg->act != AACLIP_ACT_GELU_ERF && cfg->act != AACLIP_ACT_QUICK_GELU)
    return host::fail(host::ERR_INVALID, "act must be GELU_ERF or QUICK_GELU");
  if (cfg->max_batch < 1) return host::fail(host::ERR_INVALID, "max_batch %d", cfg->max_batch);
  if (cfg->t_layers > 0 && (cfg->t_width != cfg->t_heads * 64 || cfg->t_width % 256 != 0 || cfg->t_context < 1 ||
                            cfg->t_vocab < 1 || cfg->text_adapt_until < 0 || cfg->text_adapt_until > cfg->t_layers))
    return host::fail(host::ERR_INVALID, "text tower config invalid");

  DeviceGuard guard(device);
  if (!guard.ok) return host::fail(host::ERR_CUDA, "cudaSetDevice(%d) failed", device);
  aaclip_ctx* c = new aaclip_ctx();
  c->cfg = *cfg;
  c->device = device;
  c->cta_group = (cfg->cta_group >= 1 && cfg->cta_group <= 3) ? cfg->cta_group : 0;
  c->G = cfg->image_size / cfg->patch_size;
  c->P = c->G * c->G;
  c->L = c->P + 1;
  c->E = cfg->embed_dim;
  const int K = 3 * cfg->patch_size * cfg->patch_size;
  c->Kpad = (K + 63) / 64 * 64;
  const long long w = cfg->width, E = c->E;

  int rc = host::OK;
  auto A = [&](int r) { if (rc == host::OK) rc = r; };
  c->v.width = cfg->width; c->v.heads = cfg->heads; c->v.layers = cfg->layers; c->v.mlp = cfg->mlp_width;
  A(alloc_tower(c, c->v, cfg->image_adapt_until));
  c->ln_fold = cfg->ln_fold == 1 ? true : cfg->ln_fold == 2 ? false
               : (getenv("AACLIP_LN_FOLD") ? atoi(getenv("AACLIP_LN_FOLD")) != 0 : true);
  if (w % 256 != 0 || w / 128 > 32) c->ln_fold = false;
  c->use_graphs = getenv("AACLIP_GRAPH") ? atoi(getenv("AACLIP_GRAPH")) != 0 : true;
  if (c->ln_fold) {
    const long long ffv = cfg->mlp_width;
    for (auto& l : c->v.lw) {
      A(c->alloc(&l.qkv_w32, 3 * w * w)); A(c->alloc(&l.fc_w32, ffv * w));
      A(c->alloc(&l.qkv_wf, 3 * w * w)); A(c->alloc(&l.fc_wf, ffv * w));
      A(c->alloc(&l.qkv_cs, 3 * w)); A(c->alloc(&l.qkv_bf, 3 * w)); A(c->alloc(&l.fc_cs, ffv)); A(c->alloc(&l.fc_bf, ffv));
    }
  }
  A(c->alloc(&c->conv_w, w * c->Kpad)); A(c->alloc(&c->cls, w)); A(c->alloc(&c->pos, (long long)c->L * w));
  A(c->alloc(&c->ln_pre_g, w)); A(c->alloc(&c->ln_pre_b, w)); A(c->alloc(&c->ln_post_g, w)); A(c->alloc(&c->ln_post_b, w));
  c->segdet_w.resize(cfg->n_levels);
  for (int i = 0; i < cfg->n_levels; ++i) A(c->alloc(&c->segdet_w[i], (i == cfg->n_levels - 1 ? 2 : 1) * E * w));
  long long max_w = w, max_ff = cfg->mlp_width;
  long long rows = (long long)cfg->max_batch * c->L;
  if (cfg->t_layers > 0) {
    const long long tw = cfg->t_width;
    c->t.width = cfg->t_width; c->t.heads = cfg->t_heads; c->t.layers = cfg->t_layers; c->t.mlp = 4 * cfg->t_width;
    A(alloc_tower(c, c->t, cfg->text_adapt_until));
    A(c->alloc(&c->tok_emb, (long long)cfg->t_vocab * tw)); A(c->alloc(&c->t_pos, (long long)cfg->t_context * tw));
    A(c->alloc(&c->ln_final_g, tw)); A(c->alloc(&c->ln_final_b, tw)); A(c->alloc(&c->t_final, tw * tw));
    max_w = std::max(max_w, tw); max_ff = std::max(max_ff, 4 * tw);
    rows = std::max(rows, (long long)std::max(cfg->max_text, 1) * cfg->t_context);
  }
  c->cap_rows = (int)rows;
  const long long prow = (long long)cfg->max_batch * c->P;
  A(c->alloc(&c->x, rows * max_w)); A(c->alloc(&c->a, rows * max_w)); A(c->alloc(&c->xn, rows * max_w));
  A(c->alloc(&c->qkv, rows * 3 * max_w)); A(c->alloc(&c->att, rows * max_w)); A(c->alloc(&c->h, rows * max_ff));
  A(c->alloc(&c->col, prow * c->Kpad)); A(c->alloc(&c->tap, prow * w)); A(c->alloc(&c->s, prow * 2 * E));
  A(c->alloc(&c->dots, (long long)cfg->n_levels * prow * 2)); A(c->alloc(&c->det, (long long)cfg->max_batch * E)); A(c->alloc(&c->rownorm, prow));
  A(c->alloc(&c->partials, (long long)cfg->n_levels * prow * ((E + 127) / 128) * 4));
  if (c->ln_fold) A(c->alloc(&c->part, rows * (w / 128)));
  if (rc != host::OK) { aaclip_destroy(c); return rc; }
  *out = c;
  return host::OK;
}

extern "C" void aaclip_destroy(aaclip_ctx* c) {
  if (!c) return;
  DeviceGuard guard(c->device);
  cudaDeviceSynchronize();
  for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
  for (void* p : c->allocs) cudaFree(p);
  for (auto& sl : c->slots) { if (sl.raw) cudaFree(sl.raw); if (sl.raw_scratch) cudaFree(sl.raw_scratch); }
  if (c->stage) cudaFree(c->stage);
  for (auto& sl : c->slots) {
    if (sl.in_done) cudaEventDestroy(sl.in_done);
    if (sl.comp_done) cudaEventDestroy(sl.comp_done);
    if (sl.out_done) cudaEventDestroy(sl.out_done);
  }
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->in_stream) cudaStreamDestroy(c->in_stream);
  if (c->out_stream) cudaStreamDestroy(c->out_stream);
  delete c;
}

extern "C" int aaclip_profile_enable(aaclip_ctx* c, int on) {
  if (!c) return host::fail(host::ERR_INVALID, "null context");
  c->prof_on = (on != 0);
  return host::OK;
}

// Sums the CUDA-event durations recorded since the last read into ms[cls] / counts[cls] (n_classes entries,
// class order = ProfClass) and clears the records.  Synchronises the device.
extern "C" int aaclip_profile_read(aaclip_ctx* c, double* ms, long long* counts, int n_classes) {
  if (!c || !ms || !counts) return host::fail(host::ERR_INVALID, "profile_read: null argument");
  ENTER_DEVICE(c);
  AACLIP_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < n_classes; ++i) { ms[i] = 0.0; counts[i] = 0; }
  c->prof_span_ms = 0.0;
  if (!c->prof.empty()) {
    float span = 0.f;
    if (cudaEventElapsedTime(&span, c->prof.front().a, c->prof.back().b) == cudaSuccess) c->prof_span_ms = span;
  }
  for (auto& r : c->prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cls < n_classes) { ms[r.cls] += t; counts[r.cls]++; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  c->prof.clear();
  return host::OK;
}

extern "C" double aaclip_profile_span_ms(const aaclip_ctx* c) { return c ? c->prof_span_ms : 0.0; }

extern "C" long long aaclip_device_bytes(const aaclip_ctx* c) { return c ? c->bytes : 0; }
extern "C" long long aaclip_launch_count(const aaclip_ctx* c) { return c ? c->launches : 0; }

extern "C" int aaclip_set_weight(aaclip_ctx* c, int id, int layer, const float* src, long long numel, int src_is_host,
                                 void* stream_) {
  TRY(check_ready(c));
  if (!src || numel <= 0) return host::fail(host::ERR_INVALID, "set_weight: empty source");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const aaclip_cfg& cfg = c->cfg;
  const long long w = cfg.width, ff = cfg.mlp_width, E = c->E, tw = cfg.t_width, tff = 4 * tw;
  void* dst = nullptr;
  long long expect = 0;
  bool to_bf16 = false;
  int pad_rows = 0, pad_cols = 0, pad_ld = 0;  // conv1 only
  const bool is_t = (id >= 40);
  Tower& T = is_t ? c->t : c->v;
  auto layer_ok = [&](int n) { return layer >= 0 && layer < n; };
#define PER_LAYER(FIELD, N, BF)                                                          \
  {                                                                                      \
    if (!layer_ok(T.layers)) return host::fail(host::ERR_INVALID, "set_weight: layer %d", layer); \
    dst = T.lw[layer].FIELD; expect = (N); to_bf16 = (BF);                               \
  }
  const long long W_ = is_t ? tw : w, FF_ = is_t ? tff : ff;
  switch (id) {
    case AACLIP_W_V_CONV1:
      dst = c->conv_w; expect = w * 3 * cfg.patch_size * cfg.patch_size; to_bf16 = true;
      pad_rows = (int)w; pad_cols = 3 * cfg.patch_size * cfg.patch_size; pad_ld = c->Kpad; break;
    case AACLIP_W_V_CLS: dst = c->cls; expect = w; break;
    case AACLIP_W_V_POS: dst = c->pos; expect = (long long)c->L * w; break;
    case AACLIP_W_V_LN_PRE_G: dst = c->ln_pre_g; expect = w; break;
    case AACLIP_W_V_LN_PRE_B: dst = c->ln_pre_b; expect = w; break;
    case AACLIP_W_V_LN_POST_G: dst = c->ln_post_g; expect = w; break;
    case AACLIP_W_V_LN_POST_B: dst = c->ln_post_b; expect = w; break;
    case AACLIP_W_V_LN1_G: case AACLIP_W_T_LN1_G: PER_LAYER(ln1_g, W_, false) break;
    case AACLIP_W_V_LN1_B: case AACLIP_W_T_LN1_B: PER_LAYER(ln1_b, W_, false) break;
    case AACLIP_W_V_QKV_W: case AACLIP_W_T_QKV_W: PER_LAYER(qkv_w, 3 * W_ * W_, true) break;
    case AACLIP_W_V_QKV_B: case AACLIP_W_T_QKV_B: PER_LAYER(qkv_b, 3 * W_, false) break;
    case AACLIP_W_V_OUT_W: case AACLIP_W_T_OUT_W: PER_LAYER(out_w, W_ * W_, true) break;
    case AACLIP_W_V_OUT_B: case AACLIP_W_T_OUT_B: PER_LAYER(out_b, W_, false) break;
    case AACLIP_W_V_LN2_G: case AACLIP_W_T_LN2_G: PER_LAYER(ln2_g, W_, false) break;
    case AACLIP_W_V_LN2_B: case AACLIP_W_T_LN2_B: PER_LAYER(ln2_b, W_, false) break;
    case AACLIP_W_V_FC_W: case AACLIP_W_T_FC_W: PER_LAYER(fc_w, FF_ * W_, true) break;
    case AACLIP_W_V_FC_B: case AACLIP_W_T_FC_B: PER_LAYER(fc_b, FF_, false) break;
    case AACLIP_W_V_PROJ_W: case AACLIP_W_T_PROJ_W: PER_LAYER(proj_w, W_ * FF_, true) break;
    case AACLIP_W_V_PROJ_B: case AACLIP_W_T_PROJ_B: PER_LAYER(proj_b, W_, false) break;
    case AACLIP_W_I_ADAPTER: case AACLIP_W_T_ADAPTER:
      if (!layer_ok((int)T.adapters.size())) return host::fail(host::ERR_INVALID, "set_weight: adapter %d", layer);
      dst = T.adapters[layer]; expect = W_ * W_; to_bf16 = true; break;
    case AACLIP_W_I_SEG_PROJ:
      if (!layer_ok(cfg.n_levels)) return host::fail(host::ERR_INVALID, "set_weight: level %d", layer);
      dst = c->segdet_w[layer]; expect = E * w; to_bf16 = true; break;
    case AACLIP_W_I_DET_PROJ:
      dst = c->segdet_w[cfg.n_levels - 1] + E * w; expect = E * w; to_bf16 = true; break;
    case AACLIP_W_T_TOKEN_EMB: dst = c->tok_emb; expect = (long long)cfg.t_vocab * tw; break;
    case AACLIP_W_T_POS: dst = c->t_pos; expect = (long long)cfg.t_context * tw; break;
    case AACLIP_W_T_LN_FINAL_G: dst = c->ln_final_g; expect = tw; break;
    case AACLIP_W_T_LN_FINAL_B: dst = c->ln_final_b; expect = tw; break;
    case AACLIP_W_T_FINAL_PROJ: dst = c->t_final; expect = tw * tw; to_bf16 = true; break;
    default: return host::fail(host::ERR_INVALID, "set_weight: unknown weight id %d", id);
  }
#undef PER_LAYER
  if (is_t && cfg.t_layers <= 0) return host::fail(host::ERR_STATE, "set_weight: context has no text tower");
  if (dst == nullptr) return host::fail(host::ERR_STATE, "set_weight: tensor %d not allocated", id);
  if (numel != expect)
    return host::fail(host::ERR_INVALID, "set_weight: id %d layer %d expects %lld values, got %lld", id, layer, expect, numel);
  ENTER_DEVICE(c);
  const float* dsrc = src;
  if (src_is_host) {
    if (c->stage_cap < numel) {
      AACLIP_CUDA_CHECK(cudaStreamSynchronize(st));
      if (c->stage) cudaFree(c->stage);
      c->stage = nullptr;
      AACLIP_CUDA_CHECK(cudaMalloc(&c->stage, numel * sizeof(float)));
      c->stage_cap = numel;
    }
    AACLIP_CUDA_CHECK(cudaMemcpyAsync(c->stage, src, numel * sizeof(float), cudaMemcpyHostToDevice, st));
    dsrc = c->stage;
  }
  if (c->ln_fold && !is_t) {
    float* master = nullptr;
    if (id == AACLIP_W_V_QKV_W) master = c->v.lw[layer].qkv_w32;
    if (id == AACLIP_W_V_FC_W) master = c->v.lw[layer].fc_w32;
    if (master) AACLIP_CUDA_CHECK(cudaMemcpyAsync(master, dsrc, numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    switch (id) {
      case AACLIP_W_V_LN1_G: case AACLIP_W_V_LN1_B: case AACLIP_W_V_QKV_W: case AACLIP_W_V_QKV_B:
      case AACLIP_W_V_LN2_G: case AACLIP_W_V_LN2_B: case AACLIP_W_V_FC_W: case AACLIP_W_V_FC_B:
        c->fold_dirty = true; break;
      default: break;
    }
  }
  if (!to_bf16) {
    AACLIP_CUDA_CHECK(cudaMemcpyAsync(dst, dsrc, numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else if (pad_ld) {
    cast_pad_kernel<<<1024, 256, 0, st>>>(dsrc, pad_rows, pad_cols, static_cast<bf16*>(dst), pad_ld);
    AACLIP_CUDA_CHECK(cudaGetLastError());
  } else {
    cast_pad_kernel<<<1024, 256, 0, st>>>(dsrc, 1, (int)std::min<long long>(numel, INT_MAX), static_cast<bf16*>(dst),
                                          (int)std::min<long long>(numel, INT_MAX));
    AACLIP_CUDA_CHECK(cudaGetLastError());
  }
  if (src_is_host) AACLIP_CUDA_CHECK(cudaStreamSynchronize(st));  // staging buffer is reused by the next call
  return host::OK;
}

extern "C" int aaclip_visual_forward(aaclip_ctx* c, const float* image, int B, void* const* seg_out, int seg_is_bf16,
                                     float* det_out, void* stream_) {
  TRY(check_ready(c));
  if (B < 0 || (B > 0 && !image)) return host::fail(host::ERR_INVALID, "visual_forward: B=%d image=%p", B, (const void*)image);
  TRY(check_vv_batch(c, B, "visual_forward"));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  ENTER_DEVICE(c);
  const long long img_elems = 3LL * c->cfg.image_size * c->cfg.image_size;
  for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch) {
    const int nb = std::min(c->cfg.max_batch, B - b0);
    TRY(visual_chunk(c, image + b0 * img_elems, nb, seg_out, seg_is_bf16 != 0, (long long)b0 * c->P * c->E,
                     det_out ? det_out + (long long)b0 * c->E : nullptr, nullptr, nullptr, st));
  }
  return host::OK;
}

// VisionTransformer.DAPM_replace(DPAM_layer) (model/transformer.py:406-425; train.py:243): the last DPAM_layer - 1
// blocks of the visual tower use the v-v `Attention` (transformer.py:123-152) with the block's own in_proj / out_proj
// weights.  dpam_layer <= 1 restores the ordinary attention everywhere.
extern "C" int aaclip_dapm_replace(aaclip_ctx* c, int dpam_layer) {
  TRY(check_ready(c));
  const int n = dpam_layer > 1 ? dpam_layer - 1 : 0;
  if (n > c->cfg.layers)
    return host::fail(host::ERR_INVALID, "dapm_replace: DPAM_layer %d reaches past the %d blocks of the visual tower",
                      dpam_layer, c->cfg.layers);
  const int first = n > 0 ? c->cfg.layers - n : INT_MAX;
  if (first == c->vv_first) return host::OK;
  ENTER_DEVICE(c);
  // graphs captured with the other attention are stale
  AACLIP_CUDA_CHECK(cudaDeviceSynchronize());
  for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
  c->graphs.clear();
  c->seen_once.clear();
  c->vv_first = first;
  return host::OK;
}

// CLIP.encode_image(image, out_layers, normalize) (model/model.py:185-188) = VisionTransformer.forward
// (model/transformer.py:490-551): tokens_out[i] fp32 [B, L, width] = the residual stream after block cfg.levels[i]
// (class token first), pooled_out fp32 [B, E] = ln_post(class token) @ proj.  Either may be null.  The context's levels
// are the out_layers; its seg_proj slots hold visual.proj^T (see aaclip_b200/surgery.py).
extern "C" int aaclip_encode_image(aaclip_ctx* c, const float* image, int B, float* const* tokens_out, float* pooled_out,
                                   int normalize, void* stream_) {
  TRY(check_ready(c));
  if (B < 0 || (B > 0 && !image)) return host::fail(host::ERR_INVALID, "encode_image: B=%d image=%p", B, (const void*)image);
  TRY(check_vv_batch(c, B, "encode_image"));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  ENTER_DEVICE(c);
  const long long img_elems = 3LL * c->cfg.image_size * c->cfg.image_size;
  for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch) {
    const int nb = std::min(c->cfg.max_batch, B - b0);
    TRY(visual_chunk(c, image + b0 * img_elems, nb, nullptr, 0, 0, nullptr, nullptr, nullptr, st, tokens_out,
                     (long long)b0 * c->L * c->cfg.width, pooled_out ? pooled_out + (long long)b0 * c->E : nullptr, normalize));
  }
  return host::OK;
}

namespace {
// one chunk (<= max_batch images) of the fused forward, enqueued kernel by kernel
int fused_chunk(aaclip_ctx* c, const float* image, int nb, const float* anchors, int mode, float* maps, float* scores,
                float* minmax, cudaStream_t st) {
  const int S = c->cfg.image_size;
  TRY(visual_chunk(c, image, nb, nullptr, 0, 0, c->det, anchors, c->dots, st));
  // level sum -> blur -> upsample -> map rows, with the image's extrema and score from the same CTA
  if (maps || scores) {
    RUN(PC_HEAD_MAPS, k::launch_maps_from_dots(c->dots, nullptr, c->cfg.n_levels, nb, c->G, S, mode, c->det, anchors, c->E, maps,
                                               scores, maps ? minmax : nullptr, false, st));
  }
  return host::OK;
}

// The same chunk through a CUDA graph when its key repeats (see aaclip_ctx::FusedGraph).  Not on the legacy default
// stream (capture is not allowed there), not while per-launch profiling is on, and not when the CALLER is already
// capturing this stream (a nested capture would invalidate theirs): then the launches simply join the caller's graph.
int fused_chunk_graphed(aaclip_ctx* c, const float* image, int nb, const float* anchors, int mode, float* maps,
                        float* scores, float* minmax, cudaStream_t st) {
  bool eligible = c->use_graphs && !c->prof_on && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
  bool caller_capturing = false;
  if (st != nullptr && st != cudaStreamLegacy) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    caller_capturing = (cs != cudaStreamCaptureStatusNone);
  }
  if (caller_capturing) {
    if (c->ln_fold && c->fold_dirty)
      return host::fail(host::ERR_STATE, "forward_fused: weights changed since the last forward; run one forward outside the "
                                         "stream capture first (the weight re-fold is not part of the captured step)");
    eligible = false;
  }
  if (!eligible) return fused_chunk(c, image, nb, anchors, mode, maps, scores, minmax, st);
  TRY(refold(c, st));   // never inside a capture: it runs only when weights changed
  const aaclip_ctx::FusedKey key{image, anchors, maps, scores, minmax, nb, mode};
  ++c->graph_clock;
  for (auto& g : c->graphs)
    if (g.key == key) {
      g.last_use = c->graph_clock;
      AACLIP_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
      c->launches += g.launches;
      return host::OK;
    }
  bool seen = false;
  for (auto& k2 : c->seen_once) seen = seen || (k2 == key);
  if (!seen) {   // first sight: run eagerly
    if (c->seen_once.size() >= 64) c->seen_once.erase(c->seen_once.begin());
    c->seen_once.push_back(key);
    return fused_chunk(c, image, nb, anchors, mode, maps, scores, minmax, st);
  }
  // second sight: capture, instantiate, launch
  const long long before = c->launches;
  AACLIP_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = fused_chunk(c, image, nb, anchors, mode, maps, scores, minmax, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  const long long n_launches = c->launches - before;
  c->launches = before;
  if (rc != host::OK || ce != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    c->use_graphs = false;   // something on the path is not capturable here: stay eager from now on
    if (rc != host::OK) return rc;
    return fused_chunk(c, image, nb, anchors, mode, maps, scores, minmax, st);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    cudaGetLastError();
    c->use_graphs = false;
    return fused_chunk(c, image, nb, anchors, mode, maps, scores, minmax, st);
  }
  if (c->graphs.size() >= 16) {   // evict the least recently used
    size_t v = 0;
    for (size_t i = 1; i < c->graphs.size(); ++i) if (c->graphs[i].last_use < c->graphs[v].last_use) v = i;
    cudaGraphExecDestroy(c->graphs[v].exec);
    c->graphs.erase(c->graphs.begin() + v);
  }
  c->graphs.push_back({key, exec, n_launches, c->graph_clock});
  AACLIP_CUDA_CHECK(cudaGraphLaunch(exec, st));
  c->launches += n_launches;
  return host::OK;
}
}  // namespace

extern "C" int aaclip_forward_fused(aaclip_ctx* c, const float* image, int B, const float* anchors, int mode,
                                    float* maps_out, float* scores_out, float* minmax_out, void* stream_) {
  TRY(check_ready(c));
  if (B < 0 || (B > 0 && (!image || !anchors)))
    return host::fail(host::ERR_INVALID, "forward_fused: null argument");
  if (mode != AACLIP_HEAD_TEST_INDUSTRIAL && mode != AACLIP_HEAD_TEST_MEDICAL)
    return host::fail(host::ERR_INVALID, "forward_fused: only the test modes are fused (mode=%d)", mode);
  if (minmax_out && !maps_out) return host::fail(host::ERR_INVALID, "forward_fused: extrema are produced with the maps");
  TRY(check_vv_batch(c, B, "forward_fused"));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  ENTER_DEVICE(c);
  const int S = c->cfg.image_size;
  const long long img_elems = 3LL * S * S;
  for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch) {
    const int nb = std::min(c->cfg.max_batch, B - b0);
    TRY(fused_chunk_graphed(c, image + b0 * img_elems, nb, anchors, mode,
                            maps_out ? maps_out + (long long)b0 * S * S : nullptr, scores_out ? scores_out + b0 : nullptr,
                            minmax_out ? minmax_out + 2LL * b0 : nullptr, st));
  }
  return host::OK;
}

namespace {
int ensure_host_pipeline(aaclip_ctx* c) {
  if (c->in_stream) return host::OK;
  const int S = c->cfg.image_size, mb = c->cfg.max_batch;
  AACLIP_CUDA_CHECK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  AACLIP_CUDA_CHECK(cudaStreamCreateWithFlags(&c->in_stream, cudaStreamNonBlocking));
  AACLIP_CUDA_CHECK(cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
  for (auto& sl : c->slots) {
    TRY(c->alloc(&sl.img, 3LL * mb * S * S)); TRY(c->alloc(&sl.maps, (long long)mb * S * S));
    TRY(c->alloc(&sl.scores, mb)); TRY(c->alloc(&sl.anchors, 2LL * c->E)); TRY(c->alloc(&sl.minmax, 2LL * mb));
    AACLIP_CUDA_CHECK(cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming));
    AACLIP_CUDA_CHECK(cudaEventCreateWithFlags(&sl.comp_done, cudaEventDisableTiming));
    AACLIP_CUDA_CHECK(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
  }
  return host::OK;
}

int submit_host_impl(aaclip_ctx* c, const float* host_image, const uint8_t* host_u8, int H0, int W0, int B,
                     const float* host_anchors, int mode, float* host_maps_out, float* host_scores_out,
                     float* host_minmax_out, long long* ticket) {
  TRY(check_ready(c));
  if (!host_anchors || !ticket) return host::fail(host::ERR_INVALID, "submit_host: null argument");
  if (B < 1 || B > c->cfg.max_batch)
    return host::fail(host::ERR_INVALID, "submit_host: B=%d outside [1, max_batch=%d]", B, c->cfg.max_batch);
  if (mode != AACLIP_HEAD_TEST_INDUSTRIAL && mode != AACLIP_HEAD_TEST_MEDICAL)
    return host::fail(host::ERR_INVALID, "submit_host: only the test modes are fused (mode=%d)", mode);
  if (host_minmax_out && !host_maps_out) return host::fail(host::ERR_INVALID, "submit_host: extrema are produced with the maps");
  ENTER_DEVICE(c);
  TRY(ensure_host_pipeline(c));
  aaclip_ctx::HostSlot& sl = c->slots[c->next_ticket & 1];
  if (sl.busy)
    return host::fail(host::ERR_STATE, "submit_host: ticket %lld is still in flight (at most 2 batches may be pending)",
                      sl.ticket);
  const int S = c->cfg.image_size;
  // copy-in: the slot's previous batch was waited for on the host, so its staging buffers are free
  AACLIP_CUDA_CHECK(cudaMemcpyAsync(sl.anchors, host_anchors, 2LL * c->E * sizeof(float), cudaMemcpyHostToDevice, c->in_stream));
  if (host_u8) {
    // raw uint8 images: the staging buffers grow to the largest batch seen (the slot is idle: its last ticket was waited for)
    const long long raw_bytes = 3LL * B * H0 * W0, scratch_bytes = aaclip_preprocess_scratch_bytes(B, H0, W0, S);
    if (sl.raw_cap < raw_bytes) {
      if (sl.raw) cudaFree(sl.raw);
      sl.raw = nullptr; sl.raw_cap = 0;
      AACLIP_CUDA_CHECK(cudaMalloc(&sl.raw, raw_bytes));
      sl.raw_cap = raw_bytes;
    }
    if (sl.raw_scratch_cap < scratch_bytes) {
      if (sl.raw_scratch) cudaFree(sl.raw_scratch);
      sl.raw_scratch = nullptr; sl.raw_scratch_cap = 0;
      AACLIP_CUDA_CHECK(cudaMalloc(&sl.raw_scratch, scratch_bytes));
      sl.raw_scratch_cap = scratch_bytes;
    }
    AACLIP_CUDA_CHECK(cudaMemcpyAsync(sl.raw, host_u8, raw_bytes, cudaMemcpyHostToDevice, c->in_stream));
  } else {
    AACLIP_CUDA_CHECK(cudaMemcpyAsync(sl.img, host_image, 3LL * B * S * S * sizeof(float), cudaMemcpyHostToDevice, c->in_stream));
  }
  AACLIP_CUDA_CHECK(cudaEventRecord(sl.in_done, c->in_stream));
  // compute (serial over batches on one stream: the workspaces are shared)
  AACLIP_CUDA_CHECK(cudaStreamWaitEvent(c->own_stream, sl.in_done, 0));
  if (host_u8) {   // dataset/__init__.py:127-136 on the device
    TRY(k::launch_preprocess_u8(sl.raw, B, H0, W0, S, nullptr, nullptr, sl.raw_scratch, sl.img, c->own_stream));
    c->launches += (W0 != S) ? 2 : 1;
  }
  TRY(aaclip_forward_fused(c, sl.img, B, sl.anchors, mode, host_maps_out ? sl.maps : nullptr,
                           host_scores_out ? sl.scores : nullptr, host_minmax_out ? sl.minmax : nullptr, c->own_stream));
  AACLIP_CUDA_CHECK(cudaEventRecord(sl.comp_done, c->own_stream));
  // copy-out
  AACLIP_CUDA_CHECK(cudaStreamWaitEvent(c->out_stream, sl.comp_done, 0));
  if (host_maps_out)
    AACLIP_CUDA_CHECK(cudaMemcpyAsync(host_maps_out, sl.maps, (long long)B * S * S * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
  if (host_scores_out)
    AACLIP_CUDA_CHECK(cudaMemcpyAsync(host_scores_out, sl.scores, B * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
  if (host_minmax_out)
    AACLIP_CUDA_CHECK(cudaMemcpyAsync(host_minmax_out, sl.minmax, 2LL * B * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
  AACLIP_CUDA_CHECK(cudaEventRecord(sl.out_done, c->out_stream));
  sl.busy = true;
  sl.ticket = c->next_ticket;
  *ticket = c->next_ticket++;
  return host::OK;
}
}  // namespace

extern "C" int aaclip_submit_host(aaclip_ctx* c, const float* host_image, int B, const float* host_anchors, int mode,
                                  float* host_maps_out, float* host_scores_out, float* host_minmax_out, long long* ticket) {
  if (!host_image) return host::fail(host::ERR_INVALID, "submit_host: null argument");
  return submit_host_impl(c, host_image, nullptr, 0, 0, B, host_anchors, mode, host_maps_out, host_scores_out,
                          host_minmax_out, ticket);
}
extern "C" int aaclip_submit_host_u8(aaclip_ctx* c, const uint8_t* host_u8, int B, int H0, int W0,
                                     const float* host_anchors, int mode, float* host_maps_out, float* host_scores_out,
                                     float* host_minmax_out, long long* ticket) {
  if (!host_u8 || H0 < 1 || W0 < 1) return host::fail(host::ERR_INVALID, "submit_host_u8: null image or bad size");
  return submit_host_impl(c, nullptr, host_u8, H0, W0, B, host_anchors, mode, host_maps_out, host_scores_out,
                          host_minmax_out, ticket);
}

extern "C" int aaclip_wait_host(aaclip_ctx* c, long long ticket) {
  TRY(check_ready(c));
  aaclip_ctx::HostSlot& sl = c->slots[ticket & 1];
  if (ticket < 0 || sl.ticket != ticket || !sl.busy)
    return host::fail(host::ERR_STATE, "wait_host: ticket %lld is not pending", ticket);
  ENTER_DEVICE(c);
  AACLIP_CUDA_CHECK(cudaEventSynchronize(sl.out_done));
  sl.busy = false;
  return host::OK;
}

extern "C" int aaclip_forward_fused_host(aaclip_ctx* c, const float* host_image, int B, const float* host_anchors,
                                         int mode, float* host_maps_out, float* host_scores_out, float* host_minmax_out) {
  TRY(check_ready(c));
  if (B <= 0) return host::OK;
  TRY(check_vv_batch(c, B, "forward_fused_host"));
  for (const auto& sl : c->slots)
    if (sl.busy) return host::fail(host::ERR_STATE, "forward_fused_host: ticket %lld is still pending", sl.ticket);
  const int S = c->cfg.image_size, mb = c->cfg.max_batch;
  const long long img_elems = 3LL * S * S;
  // chunks of max_batch images run through the two-slot pipeline: chunk k+1 uploads while chunk k computes
  long long pending[2] = {-1, -1};
  int n_pending = 0;
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    if (n_pending == 2) { TRY(aaclip_wait_host(c, pending[0])); pending[0] = pending[1]; n_pending = 1; }
    long long t = -1;
    TRY(aaclip_submit_host(c, host_image + b0 * img_elems, nb, host_anchors, mode,
                           host_maps_out ? host_maps_out + (long long)b0 * S * S : nullptr,
                           host_scores_out ? host_scores_out + b0 : nullptr,
                           host_minmax_out ? host_minmax_out + 2LL * b0 : nullptr, &t));
    pending[n_pending++] = t;
  }
  for (int i = 0; i < n_pending; ++i) TRY(aaclip_wait_host(c, pending[i]));
  return host::OK;
}

// What follows ln_final on the EOT row: leaky != 0 (default) - `text_adapter[-1]` = Linear + LeakyReLU (model/adapter.py:140);
// leaky == 0 - a plain projection, i.e. `@ text_projection` of the un-adapted CLIP.encode_text (model/model.py:190-200) when
// the final-projection slot holds text_projection^T and the context has no text adapters (text_adapt_until = 0).
extern "C" int aaclip_set_text_final(aaclip_ctx* c, int leaky) {
  TRY(check_ready(c));
  if (c->cfg.t_layers <= 0) return host::fail(host::ERR_STATE, "set_text_final: context has no text tower");
  c->t_final_act = leaky ? gemm::ACT_LEAKY : gemm::ACT_NONE;
  return host::OK;
}

extern "C" int aaclip_text_forward(aaclip_ctx* c, const int32_t* tokens, int n, float* out, void* stream_) {
  TRY(check_ready(c));
  const aaclip_cfg& cfg = c->cfg;
  if (cfg.t_layers <= 0) return host::fail(host::ERR_STATE, "text_forward: context has no text tower");
  if (n < 0 || (n > 0 && (!tokens || !out))) return host::fail(host::ERR_INVALID, "text_forward: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  ENTER_DEVICE(c);
  const int ctx = cfg.t_context, tw = cfg.t_width;
  const int chunk = std::max(1, c->cap_rows / ctx);
  for (int n0 = 0; n0 < n; n0 += chunk) {
    const int nn = std::min(chunk, n - n0);
    const int32_t* tk = tokens + (long long)n0 * ctx;
    text_embed_kernel<<<nn * ctx, 192, 0, st>>>(tk, c->tok_emb, c->t_pos, ctx, tw, cfg.t_vocab, c->x);
    AACLIP_CUDA_CHECK(cudaGetLastError()); c->launches++;
    bool xn_ready = false;
    for (int i = 0; i < cfg.t_layers; ++i)
      TRY(run_block(c, c->t, i, nn, ctx, 1, cfg.text_adapt_weight, &xn_ready, st));
    // ln_final is row-wise, so gather the EOT rows first, then normalise only those (model/adapter.py:138-140)
    eot_gather_kernel<<<nn, 256, 0, st>>>(tk, c->x, ctx, tw, c->a);
    AACLIP_CUDA_CHECK(cudaGetLastError()); c->launches++;
    RUN(PC_LAYERNORM, k::launch_layernorm(c->a, c->ln_final_g, c->ln_final_b, 1e-5f, nn, tw, 0, 0, 0, c->xn, nullptr, st));
    RUN(PC_OTHER, k::launch_gemm(c->xn, tw, c->t_final, tw, nn, tw, tw, nullptr, out + (long long)n0 * tw, tw, c->t_final_act,
                       gemm::OUT_F32, nullptr, 0, c->cta_group, st));
  }
  return host::OK;
}
