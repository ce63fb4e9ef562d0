// Cross-translation-unit launch functions (host side).  Every function enqueues on `stream`, returns
// host::OK or a negative error code and records a message retrievable by aaclip_last_error().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace k {

// out = epilogue(A[M,K] . W[N,K]^T); A, W bf16 K-contiguous (pitches lda/ldw elements, multiples of 8).
// act: gemm::Act, out_mode: gemm::Out, cta_group: 1 or 2.
// out_mode OUT_DOTS: see gemm::Args (anchors [dots_cols,2], partials float4 [M][dots_cols/128]).
// LnFold: the folded-LayerNorm schedule (gemm_sm100.cuh).  Producer (out_mode OUT_F32_RESID_LN): xb / ldxb / part_out.
// Consumer (bf16 outputs): part_in [M][slices] float2, colsum [N], eps; A is the bf16 copy, W the gamma-folded weight.
struct LnFold {
  void* xb = nullptr; int ldxb = 0; void* part_out = nullptr;
  const void* part_in = nullptr; int slices = 0; float eps = 1e-5f; const float* colsum = nullptr;
};
int launch_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, void* out,
                int ldo, int act, int out_mode, const float* pos, int P, int cta_group, cudaStream_t stream,
                const float* anchors = nullptr, void* partials = nullptr, int dots_cols = 0, const LnFold* ln = nullptr);

// dots[l][r] = (d0, d1) / max(sqrt(ss), 1e-12) summed over the n_slices 128-column partials of row r, level l
// partials: [n_levels][rows][n_slices] float4 (ss, d0, d1, -)
int launch_dots_finish(const void* partials, int n_levels, int rows, int n_slices, float* dots, cudaStream_t stream);

// LayerNorm over the last dim (width % 128 == 0, <= 4096), fp32 in, affine, eps; one warp per row.
//   rows are read at  x + (r / rows_per_group) * group_stride + (r % rows_per_group + row_offset) * width
//   (lets ln_post skip the CLS row of every image without a copy).  out_bf16 and/or out_f32 may be null.
int launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, int rows, int width,
                     int rows_per_group, int row_offset, long long group_stride, void* out_bf16, float* out_f32,
                     cudaStream_t stream, void* part_out = nullptr, int part_slices = 0);
//   part_out ([rows][part_slices] float2): (sum, sum of squares) of every OUTPUT row in slice 0, zeros elsewhere - with
//   out_bf16 this is what the folded-LayerNorm schedule needs from ln_pre.

// fp32 -> bf16 cast of a [rows, width] matrix.
int launch_cast_bf16(const float* x, void* out_bf16, long long n, cudaStream_t stream);

// AA-CLIP adapter mix (model/adapter.py:92-99): x <- w * a * ||x|| / ||a|| + (1-w) * x, row-wise over `width`.
// Optionally fuses the next LayerNorm: if ln_gamma != null writes LN(x_new) as bf16 to ln_out.
// Folded-LayerNorm schedule: xb_out (bf16 copy of the new x) + part_out ([rows][part_slices] float2, whole-row sums
// in slice 0) instead of ln_out.
int launch_adapter_mix(float* x, const void* a, float w, int rows, int width, const float* ln_gamma,
                       const float* ln_beta, float eps, void* ln_out_bf16, cudaStream_t stream, void* xb_out = nullptr,
                       void* part_out = nullptr, int part_slices = 0, bool a_is_bf16 = false);
// xb <- bf16(x), part[r] <- (sum, sum of squares) of row r in slice 0 (other slices zero)
int launch_rowstats_cast(const float* x, int rows, int width, void* xb, void* part, int part_slices, cudaStream_t stream);
// Wf = bf16(W o gamma), colsum[n] = sum_k Wf[n,k], bias_f = bias + W beta   (W fp32 [N,K])
int launch_fold_ln_weight(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K, void* Wf,
                          float* colsum, float* bias_f, cudaStream_t stream);

// Row-wise L2 normalise (F.normalize, eps 1e-12) of s[rows, ld] columns [col0, col0+width).
//   out_f32 / out_bf16: normalised rows [rows, width] (either may be null)
//   anchors != null ([width, 2] fp32): also dots[r] = (<f_r, T[:,0]>, <f_r, T[:,1]>) as float2
int launch_l2norm_rows(const float* s, int ld, int col0, int rows, int width, float* out_f32, void* out_bf16,
                       const float* anchors, float* dots, cudaStream_t stream);

// det token: mean over P patches of the L2-normalised rows of s[b*P + p, col0:col0+width] -> det[b, width]
// (two launches; inv_scratch: B*P floats for the inverse row norms)
int launch_det_mean(const float* s, int ld, int col0, int B, int P, int width, float* inv_scratch, float* det,
                    cudaStream_t stream);

// patch-embedding im2col: image fp32 [B,3,S,S] -> bf16 [B*G*G, Kpad], k = c*ps*ps + i*ps + j, zero padded
int launch_im2col(const float* image, int B, int S, int ps, int Kpad, void* out_bf16, cudaStream_t stream);

// class-token rows: x[b*L + 0, :] = cls + pos[0]
int launch_cls_rows(float* x, const float* cls, const float* pos, int B, int L, int width, cudaStream_t stream);

// Multi-head attention, head dim 64, bf16 qkv[M = B*L, 3*width] (q | k | v, heads contiguous 64-wide),
// out bf16 [M, width].  causal: additive -inf above the diagonal (text tower).  scale = 1/sqrt(64).
int launch_attention(const void* qkv, void* out, int B, int L, int heads, int causal, cudaStream_t stream);

// v-v attention of the surgery feature extractor (model/transformer.py:123-152 as installed by DAPM_replace :406-425):
// v bf16 [B*L, ldv] (value projection, heads contiguous 64-wide), out bf16 [B*L, ldo]; for every (token, head) the
// B images of the batch attend to each other: softmax(v v^T / 8) v.  B <= vv_attention_max_batch().  See vv_attn.cu.
int launch_vv_attention(const void* v, int ldv, void* out, int ldo, int B, int L, int heads, cudaStream_t stream);
int vv_attention_max_batch();

// tokens[b, p, :] += vec[b, :]   (train.py:85: patch features + the image's class feature)
int launch_add_image_vector(float* tokens, const float* vec, int B, int P, int E, cudaStream_t stream);

// Anomaly-map head (forward_utils.py:196-216 + test.py:83-93), see head.cu
int launch_patch_dots(const void* const* seg, int n_levels, int seg_is_bf16, const float* anchors, int anchors_batched,
                      int B, int P, int E, float* dots /*[n_levels][B*P][2]*/, cudaStream_t stream);
int launch_head_maps(const float* dots, int B, int G, int S, int mode, int n_levels, float* maps, cudaStream_t stream);
int launch_scores(const float* det, const float* anchors, int anchors_batched, int B, int E, float* scores,
                  cudaStream_t stream);
// test modes from dots [n_levels][B*P][2] (level sum formed here) or msum [B][P] (already level-summed): blur ->
// upsample -> maps [B,S,S] (+ minmax [B,2], + scores [B] from det / anchors); any of maps / scores / minmax may be null.
// A cluster of 1, 2 or 4 CTAs per image (head_epilogue.cuh); pdl chains it to the stream's previous kernel.
int launch_maps_from_dots(const float* dots, const float* msum, int n_levels, int B, int G, int S, int mode, const float* det,
                          const float* anchors, int E, float* maps, float* scores, float* minmax, bool pdl,
                          cudaStream_t stream);
// The token-streaming half of the A7 contract (head_stream.cu): test modes, shared anchors, E = 768, bf16 or fp32
// tokens, 16-byte aligned level pointers -> level-summed scalars [B, P] at the start of `workspace`
// (head_stream_workspace_bytes(B, P) bytes) and scores [B].
long long head_stream_workspace_bytes(int B, int P);
bool head_stream_supported(int n_levels, int seg_is_bf16, int anchors_batched, int E, int P, int G, int S, int mode,
                           const void* const* seg);
int launch_head_stream(const void* const* seg, int n_levels, int seg_is_bf16, const float* anchors, const float* det,
                       int B, int P, float* scores, void* workspace, cudaStream_t stream);

// Loader-side transform (dataset/__init__.py:127-136): u8 [B,H0,W0,3] -> PIL-bicubic resize -> /255 -> normalise ->
// fp32 [B,3,S,S]; scratch: 3*B*H0*S bytes (unused when W0 == S); mean/std: host float[3] or null (CLIP constants).
int launch_preprocess_u8(const uint8_t* images, int B, int H0, int W0, int S, const float* mean, const float* stdv,
                         uint8_t* scratch, float* out, cudaStream_t stream);

}  // namespace k
