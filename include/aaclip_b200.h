/* aaclip_b200 — C ABI of the B200-native AA-CLIP inference hot path.
 *
 * The reference (wei-paul/AA-CLIP) has no FFI of its own: the boundary its callers (test.py:get_predictions,
 * forward_utils.get_adapted_text_embedding) see is the Python surface  AdaptedCLIP.forward / .encode_text
 * (model/adapter.py:67-145)  and  calculate_similarity_map (forward_utils.py:196-216).  aaclip_b200/ mirrors
 * that surface in Python and binds THIS library with ctypes (INTEGRATION.md shows the stub); every entry point
 * below names the reference code it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *   - unless a parameter is named host_*, pointers are DEVICE pointers owned by the caller; the context
 *     owns only its packed weights and workspaces.
 *   - calls enqueue on `stream` and return 0 (AACLIP_OK) or a negative code; aaclip_last_error() returns the
 *     message for the calling thread.  There is no CPU fallback: without an sm_100 device calls fail.
 *   - any number of contexts per process, on any devices; every call runs on its context's device and restores the
 *     caller's current device before it returns.  A context is not thread-safe.
 *   - ABI version 2 (aaclip_abi_version): v2 added seg_is_bf16 to aaclip_visual_forward, minmax_out to the fused
 *     entries and minmax_out + a caller-owned workspace to aaclip_anomaly_head.
 */
#ifndef AACLIP_B200_H_
#define AACLIP_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AACLIP_OK 0
#define AACLIP_ERR_INVALID (-1)
#define AACLIP_ERR_CUDA (-2)
#define AACLIP_ERR_NO_DEVICE (-3)
#define AACLIP_ERR_STATE (-4)

/* activation of the MLP (model/model.py:84: nn.GELU unless quick_gelu; QuickGELU model/transformer.py:46-49) */
#define AACLIP_ACT_NONE 0
#define AACLIP_ACT_GELU_ERF 1
#define AACLIP_ACT_QUICK_GELU 2
#define AACLIP_ACT_LEAKY 3 /* nn.LeakyReLU(0.01), model/adapter_modules.py:9,20 */

/* GEMM output modes */
#define AACLIP_OUT_BF16 0
#define AACLIP_OUT_F32 1
#define AACLIP_OUT_F32_RESID 2 /* out += result (fp32 residual stream, model/transformer.py:256-257) */
#define AACLIP_OUT_F32_PATCH 3 /* patch-embed scatter + positional embedding (model/adapter.py:68-82) */

/* anomaly-map head modes (forward_utils.py:196-216; DOMAINS in dataset/constants.py) */
#define AACLIP_HEAD_TEST_INDUSTRIAL 0 /* (s1+1-s0)/2, gaussian 7x7 sigma 1, bilinear align_corners */
#define AACLIP_HEAD_TEST_MEDICAL 1    /* same with gaussian 9x9 sigma 1.5 */
#define AACLIP_HEAD_TRAIN_SOFTMAX 2   /* bilinear on both channels, softmax over channels, no blur */

/* tensors accepted by aaclip_set_weight; names are the reference state_dict keys */
typedef enum {
  /* visual tower: clipmodel.visual.* (model/transformer.py:359-404) */
  AACLIP_W_V_CONV1 = 0,       /* conv1.weight [width,3,ps,ps] */
  AACLIP_W_V_CLS = 1,         /* class_embedding [width] */
  AACLIP_W_V_POS = 2,         /* positional_embedding [L,width] */
  AACLIP_W_V_LN_PRE_G = 3, AACLIP_W_V_LN_PRE_B = 4,
  AACLIP_W_V_LN_POST_G = 5, AACLIP_W_V_LN_POST_B = 6,
  /* per layer: transformer.resblocks.{i}.* (model/transformer.py:183-224) */
  AACLIP_W_V_LN1_G = 10, AACLIP_W_V_LN1_B = 11,
  AACLIP_W_V_QKV_W = 12, AACLIP_W_V_QKV_B = 13,    /* attn.in_proj_weight [3w,w], in_proj_bias */
  AACLIP_W_V_OUT_W = 14, AACLIP_W_V_OUT_B = 15,    /* attn.out_proj.weight [w,w], bias */
  AACLIP_W_V_LN2_G = 16, AACLIP_W_V_LN2_B = 17,
  AACLIP_W_V_FC_W = 18, AACLIP_W_V_FC_B = 19,      /* mlp.c_fc [4w,w] */
  AACLIP_W_V_PROJ_W = 20, AACLIP_W_V_PROJ_B = 21,  /* mlp.c_proj [w,4w] */
  /* image_adapter.* (model/adapter.py:27-39) */
  AACLIP_W_I_ADAPTER = 30,    /* layer_adapters.{i}.fc.0.weight [w,w] */
  AACLIP_W_I_SEG_PROJ = 31,   /* seg_proj.{i}.fc(.0).weight [E,w] */
  AACLIP_W_I_DET_PROJ = 32,   /* det_proj.fc(.0).weight [E,w] */
  /* text tower: clipmodel.* (model/model.py:165-174) */
  AACLIP_W_T_TOKEN_EMB = 40,  /* token_embedding.weight [vocab,tw] */
  AACLIP_W_T_POS = 41,        /* positional_embedding [ctx,tw] */
  AACLIP_W_T_LN_FINAL_G = 42, AACLIP_W_T_LN_FINAL_B = 43,
  AACLIP_W_T_LN1_G = 50, AACLIP_W_T_LN1_B = 51,
  AACLIP_W_T_QKV_W = 52, AACLIP_W_T_QKV_B = 53,
  AACLIP_W_T_OUT_W = 54, AACLIP_W_T_OUT_B = 55,
  AACLIP_W_T_LN2_G = 56, AACLIP_W_T_LN2_B = 57,
  AACLIP_W_T_FC_W = 58, AACLIP_W_T_FC_B = 59,
  AACLIP_W_T_PROJ_W = 60, AACLIP_W_T_PROJ_B = 61,
  /* text_adapter.* (model/adapter.py:41-44) */
  AACLIP_W_T_ADAPTER = 70,    /* {i}.fc.0.weight [tw,tw], i < text_adapt_until */
  AACLIP_W_T_FINAL_PROJ = 71  /* {text_adapt_until}.fc.0.weight [tw,tw] (Linear + LeakyReLU) */
} aaclip_weight_id;

typedef struct {
  /* visual tower (model/model_configs/ViT-L-14-336.json) */
  int image_size;   /* 336 (518 for the paper setting, test.py:111) */
  int patch_size;   /* 14 */
  int width;        /* 1024 */
  int heads;        /* 16 (head dim must be 64) */
  int layers;       /* 24 */
  int mlp_width;    /* 4096 */
  int embed_dim;    /* 768 */
  int act;          /* AACLIP_ACT_GELU_ERF or AACLIP_ACT_QUICK_GELU */
  /* AdaptedCLIP ctor arguments (model/adapter.py:7-17) */
  int image_adapt_until;    /* 6 */
  float image_adapt_weight; /* 0.1 */
  int n_levels;             /* 4 */
  int levels[8];            /* {6,12,18,24}: tap after block levels[i] (1-based) */
  int proj_relu;            /* SimpleProj relu flag (test.py --relu, default 0) */
  /* text tower; t_layers == 0 disables it */
  int t_context;    /* 77 */
  int t_vocab;      /* 49408 */
  int t_width;      /* 768 */
  int t_heads;      /* 12 */
  int t_layers;     /* 12 */
  int text_adapt_until;     /* 3 */
  float text_adapt_weight;  /* 0.1 */
  /* capacity */
  int max_batch;    /* images per visual forward the workspaces are sized for */
  int max_text;     /* sentences per text forward */
  int cta_group;    /* GEMM tile shape: 0 = chosen per launch (wave model); 1 = cta_group::1, 128x256 tiles; 2 = cta_group::2
                       CTA pairs, 256x256 tiles; 3 = cta_group::1, 128x128 tiles (small batches) */
  int ln_fold;      /* visual tower: 0 = library default (on, AACLIP_LN_FOLD=0 turns it off), 1 = fold ln_1 / ln_2 into
                       the in_proj / c_fc GEMMs (no LayerNorm launches inside the blocks), 2 = separate LayerNorm kernels */
} aaclip_cfg;

typedef struct aaclip_ctx aaclip_ctx;

const char* aaclip_last_error(void);
int aaclip_abi_version(void);

/* ---- context ---------------------------------------------------------------------------------------- */
int aaclip_create(aaclip_ctx** out, const aaclip_cfg* cfg, int device);
void aaclip_destroy(aaclip_ctx* ctx);
/* Copies `numel` fp32 values from DEVICE or HOST memory (src_is_host) into the context's packed layout
 * (bf16 for GEMM operands, fp32 otherwise).  `layer` indexes per-layer / per-level tensors, else 0.
 * Replaces nn.Module.load_state_dict for clipmodel.visual.*, image_adapter.*, text_adapter.* (test.py:163-176). */
int aaclip_set_weight(aaclip_ctx* ctx, int weight_id, int layer, const float* src, long long numel, int src_is_host,
                      void* stream);
/* Bytes of device memory the context holds (weights + workspaces). */
long long aaclip_device_bytes(const aaclip_ctx* ctx);
/* Number of kernels the library launched on behalf of ctx since creation (for bench.py gpu_launches). */
long long aaclip_launch_count(const aaclip_ctx* ctx);

/* Optional per-launch timing: while enabled every kernel the context launches is bracketed by CUDA events on
 * its stream.  aaclip_profile_read sums them per kernel class into ms[i] / counts[i] (i < n_classes <= 16) and
 * clears the log.  Class order: gemm_qkv, gemm_out, gemm_fc, gemm_proj, gemm_adapter, gemm_segdet, gemm_patch,
 * attention, layernorm, adapter_mix, cast, l2norm, det_mean, stem_misc, head_maps, other. */
#define AACLIP_PROFILE_CLASSES 16
int aaclip_profile_enable(aaclip_ctx* ctx, int on);
int aaclip_profile_read(aaclip_ctx* ctx, double* ms, long long* counts, int n_classes);
/* Milliseconds from the start of the first to the end of the last launch of the log aaclip_profile_read consumed
 * last (span - sum of durations = idle time between kernels). */
double aaclip_profile_span_ms(const aaclip_ctx* ctx);

/* ---- AdaptedCLIP.forward (model/adapter.py:67-112) ------------------------------------------------- */
/* image fp32 [B,3,S,S] (CLIP-normalised).  seg_out[i]: [B,P,E] L2-normalised patch tokens of level i, fp32 (the
 * reference's dtype; seg_is_bf16 = 0) or bf16 (half the bytes for aaclip_anomaly_head to stream), or NULL to skip
 * materialising them; det_out fp32 [B,E] (or NULL). */
int aaclip_visual_forward(aaclip_ctx* ctx, const float* image, int B, void* const* seg_out, int seg_is_bf16,
                          float* det_out, void* stream);

/* ---- calculate_similarity_map + test.py:83-93 ------------------------------------------------------- */
/* seg[i]: [B,P,E] fp32 (seg_is_bf16 = 0) or bf16 normalised patch tokens, n_levels of them (1 = one
 * calculate_similarity_map call, 4 = the whole test.py:89-93 loop); anchors fp32 [E,2] (anchors_batched = 0) or
 * [B,E,2]; det fp32 [B,E] or NULL.
 * maps_out: test modes -> fp32 [B,S,S] = sum over levels of the per-level map (test.py:93);
 *           train mode -> fp32 [n_levels,B,2,S,S] softmaxed per level (forward_utils.py:214-215).
 * scores_out: fp32 [B] = ((det . anchors)[:,1] + 1)/2 (test.py:83-84) or NULL.
 * minmax_out: fp32 [B,2] = per-image (min, max) of maps_out (test modes, shared anchors; what metrics_eval's
 *           normalisation needs, forward_utils.py:241-252) or NULL.
 * workspace: aaclip_anomaly_head_workspace_bytes(n_levels, B, P) bytes of 16-byte aligned device memory owned by the
 *           caller (needed when maps_out != NULL): the call allocates nothing, never synchronises and is capturable. */
long long aaclip_anomaly_head_workspace_bytes(int n_levels, int B, int P);
int aaclip_anomaly_head(const void* const* seg, int n_levels, int seg_is_bf16, const float* anchors,
                        int anchors_batched, const float* det, int B, int P, int E, int img_size, int mode,
                        float* maps_out, float* scores_out, float* minmax_out, void* workspace,
                        long long workspace_bytes, void* stream);

/* Per-image extrema of anomaly maps: maps fp32 [B, n_pix] -> out fp32 [B][2] = (min, max).  The pixel-side input of
 * metrics_eval's image-level score (forward_utils.py:241-254: global min-max normalisation + per-image max); exact,
 * so callers combine them over batches on the host. */
int aaclip_map_minmax(const float* maps, int B, long long n_pix, float* out, void* stream);

/* ---- fused image -> anomaly map (AdaptedCLIP.forward + head, no seg-token materialisation) ---------- */
/* minmax_out: fp32 [B,2] per-image (min, max) of the maps, written by the head's epilogue (no second pass over the
 * maps), or NULL. */
int aaclip_forward_fused(aaclip_ctx* ctx, const float* image, int B, const float* anchors /*[E,2]*/, int mode,
                         float* maps_out /*[B,S,S]*/, float* scores_out /*[B]*/, float* minmax_out /*[B,2]*/,
                         void* stream);
/* Same with HOST buffers (pinned or pageable): H2D of the images, D2H of maps and scores, synchronous.
 * B > max_batch is processed in chunks through the two-slot pipeline below. */
int aaclip_forward_fused_host(aaclip_ctx* ctx, const float* host_image, int B, const float* host_anchors, int mode,
                              float* host_maps_out, float* host_scores_out, float* host_minmax_out /*[B,2] or NULL*/);

/* Pipelined form of the host-buffer entry, for a loop over batches (test.py:get_predictions, test.py:53-99):
 * submit enqueues H2D (copy-in stream) -> forward (compute stream) -> D2H (copy-out stream) for one batch of
 * B <= max_batch images and returns a ticket at once; wait blocks until that batch's maps and scores have landed in
 * the host buffers.  Two batches may be in flight, so the copies of neighbouring batches overlap the compute.
 * Host buffers must stay valid (and should be pinned) until the ticket has been waited for. */
int aaclip_submit_host(aaclip_ctx* ctx, const float* host_image, int B, const float* host_anchors, int mode,
                       float* host_maps_out, float* host_scores_out, float* host_minmax_out /*[B,2] or NULL*/,
                       long long* ticket);
int aaclip_wait_host(aaclip_ctx* ctx, long long ticket);

/* ---- loader-side image transform (dataset/__init__.py:127-136, :53-62) ------------------------------ */
/* transforms.Resize((S,S), Image.BICUBIC) + ToTensor + Normalize on the device, bit-exact with PIL + torchvision
 * (Pillow's fixed-point antialiased resample, IEEE float normalisation).
 * images: uint8 [B,H0,W0,3] RGB (device); out: fp32 [B,3,S,S]; host_mean / host_std: HOST float[3] or NULL for the
 * CLIP constants of dataset/__init__.py:130-133; scratch: device, aaclip_preprocess_scratch_bytes() bytes (the
 * uint8 result of the horizontal pass; may be NULL when W0 == S). */
long long aaclip_preprocess_scratch_bytes(int B, int H0, int W0, int S);
int aaclip_preprocess_u8(const uint8_t* images, int B, int H0, int W0, int S, const float* host_mean,
                         const float* host_std, uint8_t* scratch, float* out, void* stream);
/* The bare PIL.Image.resize((S,S), BICUBIC): out_u8 uint8 [B,S,S,3] (what transform_mask-style callers and the
 * parity tests compare byte for byte). */
int aaclip_resize_bicubic_u8(const uint8_t* images, int B, int H0, int W0, int S, uint8_t* scratch, uint8_t* out_u8,
                             void* stream);
/* aaclip_submit_host with RAW images: host_u8 uint8 [B,H0,W0,3]; H2D of the bytes, transform on the device, then
 * the fused forward.  Same ticket / wait protocol (and the same two slots) as aaclip_submit_host. */
int aaclip_submit_host_u8(aaclip_ctx* ctx, const uint8_t* host_u8, int B, int H0, int W0, const float* host_anchors,
                          int mode, float* host_maps_out, float* host_scores_out, float* host_minmax_out,
                          long long* ticket);

/* ---- stage-1 feature extraction of train.py (forward only, runs under no_grad) ------------ */
/* VisionTransformer.DAPM_replace(DPAM_layer) (model/transformer.py:406-425; train.py:243, default 20): the last
 * DPAM_layer - 1 blocks of the visual tower switch to the v-v `Attention` (model/transformer.py:123-152) with their
 * own in_proj / out_proj weights; DPAM_layer <= 1 switches back.  That attention reads the block's [L, batch, D]
 * tensor as (B, N, C), so its softmax runs over the IMAGES OF THE BATCH for every token position and head: results
 * depend on the batch composition, and such a context refuses batches it would have to split
 * (B > min(max_batch, 128)).  Graphs cached by the fused forward are dropped. */
int aaclip_dapm_replace(aaclip_ctx* ctx, int dpam_layer);
/* CLIP.encode_image(image, out_layers, normalize) (model/model.py:185-188) = VisionTransformer.forward
 * (model/transformer.py:490-551) on a context whose cfg.levels are the out_layers, whose adapters are off
 * (image_adapt_until = 0) and whose seg_proj slots hold visual.proj^T:
 *   tokens_out[i]  fp32 [B, L, width] or NULL: residual stream after block levels[i], class token first
 *                  (Transformer.forward out_tokens, model/transformer.py:296-318); tokens_out itself may be NULL
 *   pooled_out     fp32 [B, E] or NULL: ln_post(class token) @ proj, L2-normalised when normalize != 0.
 * On such a context aaclip_visual_forward returns normalize(ln_post(tokens[:, 1:]) @ proj) per level: train.py:78-84. */
int aaclip_encode_image(aaclip_ctx* ctx, const float* image, int B, float* const* tokens_out, float* pooled_out,
                        int normalize, void* stream);
/* The v-v attention alone: v bf16 [B*L, ldv] (value projection in columns [0, heads*64)), out bf16 [B*L, ldo];
 * out[b, l, h] = sum_b' softmax_b'(<v[b,l,h], v[b',l,h]> / 8) v[b',l,h].  B <= 128. */
int aaclip_vv_attention(const void* v, int ldv, void* out, int ldo, int B, int L, int heads, void* stream);
/* tokens fp32 [B, P, E] += vec fp32 [B, E] broadcast over the P patches (train.py:85, `t + cls_token.unsqueeze(1)`). */
int aaclip_add_image_vector(float* tokens, const float* vec, int B, int P, int E, void* stream);

/* ---- AdaptedCLIP.encode_text(adapt_text=True) (model/adapter.py:114-145) ---------------------------- */
/* tokens int32 [n, context] (model/tokenizer.py:150-185); out fp32 [n, t_width], un-normalised. */
int aaclip_text_forward(aaclip_ctx* ctx, const int32_t* tokens, int n, float* out, void* stream);
/* What follows ln_final on the EOT row of aaclip_text_forward: leaky != 0 (default) - text_adapter[-1] = Linear + LeakyReLU
 * (model/adapter.py:140); leaky == 0 - a plain projection: with text_projection^T in the final-projection slot and
 * text_adapt_until = 0 the entry is the un-adapted CLIP.encode_text (model/model.py:190-200; test.py:197-200 builds the
 * anchors from it when no text adapter is used). */
int aaclip_set_text_final(aaclip_ctx* ctx, int leaky);

/* forward_utils.py:155-161: emb fp32 [n, width] (encode_text output of one prompt state) -> rows L2-normalised,
 * averaged, re-normalised, written to column `col` (0 = normal, 1 = abnormal) of anchors fp32 [width, 2]. */
int aaclip_text_anchor(const float* emb, int n, int width, float* anchors, int col, void* stream);

/* ---- building blocks (exported so the parity tests can pin each kernel on its own) ------------------ */
/* out = epilogue(A[M,K] . W[N,K]^T), bf16 operands (pitches lda/ldw elements), fp32 accumulation.
 * cta_group: tile shape, same values as the field of that name in the context configuration (0 = chosen from M, N and the
 * SM count). */
int aaclip_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                     void* out, int ldo, int act, int out_mode, const float* pos, int P, int cta_group, void* stream);
/* The folded-LayerNorm schedule of the visual tower (LN(x) W^T + b == rstd (x (W o gamma)^T - mean s) + b'): the
 * residual GEMM leaves the new fp32 rows, their bf16 copy and per-128-column (sum, sum of squares); the next GEMM
 * consumes the bf16 copy and applies the row statistics in its epilogue - no LayerNorm kernel in between.
 *   aaclip_gemm_resid_ln : x <- x + A W^T + bias (fp32 [M,N], in place), xb <- bf16(x), part[M][N/128] float2
 *   aaclip_gemm_lnfold   : out bf16 [M,N] <- act(rstd_r (A Wf^T - mean_r colsum[n]) + bias[n]); statistics over K from
 *                          part [M][slices]
 *   aaclip_rowstats_cast : xb <- bf16(x), part[r][0] <- whole-row sums (other slices 0)
 *   aaclip_fold_ln_weight: Wf = bf16(W o gamma) [N,K], colsum[n] = sum_k Wf[n,k], bias_f = bias + W beta */
int aaclip_gemm_resid_ln(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, float* x,
                         int ldx, void* xb, int ldxb, void* part_out, int cta_group, void* stream);
int aaclip_gemm_lnfold(const void* A, int lda, const void* Wf, int ldw, int M, int N, int K, const float* bias,
                       const float* colsum, const void* part, int slices, float eps, void* out, int ldo, int act,
                       int cta_group, void* stream);
int aaclip_rowstats_cast(const float* x, int rows, int width, void* xb, void* part, int part_slices, void* stream);
int aaclip_fold_ln_weight(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K, void* Wf,
                          float* colsum, float* bias_f, void* stream);
/* LayerNorm (model/transformer.py:37-43), fp32 in -> bf16 and/or fp32 out. */
int aaclip_layernorm(const float* x, const float* gamma, const float* beta, float eps, int rows, int width,
                     void* out_bf16, float* out_f32, void* stream);
/* softmax(q k^T / 8) v per head, head dim 64 (nn.MultiheadAttention core, model/transformer.py:237);
 * qkv bf16 [B*L, 3*heads*64], out bf16 [B*L, heads*64]; causal != 0 applies CLIP's text mask (model/model.py:172). */
int aaclip_attention(const void* qkv, void* out, int B, int L, int heads, int causal, void* stream);
/* Diagnostics: aaclip_attention + clock64() stamps of CTA `cta` (softmax warp 0: slots 0..7, MMA thread: slots
 * 16..20) per key tile into trace[slot*16 + tile] (device int64[24*16]). */
int aaclip_attention_trace(const void* qkv, void* out, int B, int L, int heads, int causal, long long* trace, int cta,
                           void* stream);
/* adapter mix x <- w*a*||x||/||a|| + (1-w)*x (model/adapter.py:93-99) */
int aaclip_adapter_mix(float* x, const float* a, float w, int rows, int width, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AACLIP_B200_H_ */
