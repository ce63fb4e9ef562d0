// Anomaly-map head: calculate_similarity_map (forward_utils.py:196-216) + the caller arithmetic of
// test.py:83-93 (level cat + sum, image score), as two HBM-bound kernels:
//
//   patch_dots_kernel : one warp per patch token; 16-byte coalesced loads of the (bf16 or fp32) normalised
//                       tokens of every level, warp-shuffle dot products with both text anchors.
//   head_maps_kernel  : per (image, band of output rows): builds A = U_bilinear . G_blur  (S x G, the blur
//                       with reflect padding and the align_corners=True upsample are both separable and
//                       linear, so map = A . m . A^T), then writes the fp32 map with coalesced stores.
//                       Train mode (test=False): no blur, both channels, softmax over channels, per level.
//
// Test mode, summed over levels (test.py:93):   m[h][w] = sum_l (100*(d1-d0) + 1)/2 = 50*sum_l(d1-d0) + n_levels/2
#include <stdarg.h>
#include <algorithm>
#include "common.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace {

constexpr int MAX_LEVELS = 8;
struct SegPtrs { const void* p[MAX_LEVELS]; };

// dots[l][row][c] = <seg_l[row, :], anchors[:, c]>   (row = b*P + p), c in {0,1}
template <bool BF16>
__global__ void __launch_bounds__(256)
patch_dots_kernel(SegPtrs seg, int n_levels, const float* __restrict__ anchors, int anchors_batched, int rows, int P,
                  int E, float* __restrict__ dots) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* T = anchors + (anchors_batched ? (size_t)(row / P) * E * 2 : 0);
  for (int l = 0; l < n_levels; ++l) {
    float d0 = 0.f, d1 = 0.f;
    if constexpr (BF16) {
      const __nv_bfloat16* f = static_cast<const __nv_bfloat16*>(seg.p[l]) + (size_t)row * E;
      for (int c = lane * 8; c < E; c += 256) {
        const uint4 raw = *reinterpret_cast<const uint4*>(f + c);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 v = __bfloat1622float2(h[j]);
          const float4 t = __ldg(reinterpret_cast<const float4*>(T + (c + 2 * j) * 2));  // T[c][0..1], T[c+1][0..1]
          d0 += v.x * t.x + v.y * t.z;
          d1 += v.x * t.y + v.y * t.w;
        }
      }
    } else {
      const float* f = static_cast<const float*>(seg.p[l]) + (size_t)row * E;
      for (int c = lane * 4; c < E; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(f + c);
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(T + c * 2));
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(T + c * 2 + 4));
        d0 += (v.x * t0.x + v.y * t0.z) + (v.z * t1.x + v.w * t1.z);
        d1 += (v.x * t0.y + v.y * t0.w) + (v.z * t1.y + v.w * t1.w);
      }
    }
    d0 = ptx::warp_sum(d0);
    d1 = ptx::warp_sum(d1);
    if (lane == 0) {
      float2 o = make_float2(d0, d1);
      *reinterpret_cast<float2*>(dots + ((size_t)l * rows + row) * 2) = o;
    }
  }
}

__device__ __forceinline__ int reflect_idx(int i, int G) {
  if (i < 0) i = -i;
  if (i >= G) i = 2 * (G - 1) - i;
  return i;
}

constexpr int BAND = 16;       // output rows per block
constexpr int HEAD_THREADS = 256;

// mode: AACLIP_HEAD_*.  dots: [n_levels][B*G*G][2].
// test modes : maps [B][S][S]; level-summed.  grid = (bands, B, 1)
// train mode : maps [n_levels][B][2][S][S]; softmax over the 2 channels. grid = (bands, B, n_levels)
__global__ void __launch_bounds__(HEAD_THREADS)
head_maps_kernel(const float* __restrict__ dots, int n_levels, int B, int G, int S, int mode,
                 float* __restrict__ maps) {
  extern __shared__ float sm[];
  float* A = sm;                    // [S][G]
  float* m = A + (size_t)S * G;     // [C][G][G]
  float* t = m + 2 * G * G;         // [G][HEAD_THREADS]
  __shared__ float wk[16];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int lvl = blockIdx.z;
  const bool train = (mode == AACLIP_HEAD_TRAIN_SOFTMAX);
  const int ksize = train ? 1 : (mode == AACLIP_HEAD_TEST_INDUSTRIAL ? 7 : 9);
  const float sigma = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 1.0f : 1.5f;
  const int P = G * G;
  const size_t rows = (size_t)B * P;

  // 1-D gaussian taps, normalised to sum 1 (kornia 0.6.9 gaussian(); odd window)
  if (tid == 0) {
    float sum = 0.f;
    for (int i = 0; i < ksize; ++i) {
      const float x = float(i - ksize / 2);
      wk[i] = train ? 1.0f : expf(-(x * x) / (2.0f * sigma * sigma));
      sum += wk[i];
    }
    for (int i = 0; i < ksize; ++i) wk[i] /= sum;
  }
  // per-patch scalars
  for (int i = tid; i < P; i += HEAD_THREADS) {
    if (train) {
      const float2 d = *reinterpret_cast<const float2*>(dots + ((size_t)lvl * rows + (size_t)b * P + i) * 2);
      m[i] = 100.0f * d.x;
      m[P + i] = 100.0f * d.y;
    } else {
      float acc = 0.f;
      for (int l = 0; l < n_levels; ++l) {
        const float2 d = *reinterpret_cast<const float2*>(dots + ((size_t)l * rows + (size_t)b * P + i) * 2);
        // per level exactly as the reference: (s1 + 1 - s0) / 2 with s = 100 * dot
        acc += (100.0f * d.y + 1.0f - 100.0f * d.x) * 0.5f;
      }
      m[i] = acc;
    }
  }
  __syncthreads();
  // A[x][g] = l0 * Gm[r0][g] + l1 * Gm[r1][g],  Gm = blur matrix with reflect padding (identity in train mode)
  const float scale = (S > 1) ? float(G - 1) / float(S - 1) : 0.f;
  for (int i = tid; i < S * G; i += HEAD_THREADS) {
    const int x = i / G, g = i - x * G;
    const float src = scale * float(x);
    int r0 = int(src);
    if (r0 > G - 1) r0 = G - 1;
    const int r1 = r0 + ((r0 < G - 1) ? 1 : 0);
    const float l1 = src - float(r0), l0 = 1.0f - l1;
    float g0 = 0.f, g1 = 0.f;
    for (int tt = 0; tt < ksize; ++tt) {
      const int off = tt - ksize / 2;
      if (reflect_idx(r0 + off, G) == g) g0 += wk[tt];
      if (reflect_idx(r1 + off, G) == g) g1 += wk[tt];
    }
    A[i] = l0 * g0 + l1 * g1;
  }
  __syncthreads();

  const int y0 = blockIdx.x * BAND;
  const int ny = min(BAND, S - y0);
  const int C = train ? 2 : 1;
  for (int x = tid; x < S; x += HEAD_THREADS) {
    float o0[BAND];
    const float* Ax = A + (size_t)x * G;
    for (int c = 0; c < C; ++c) {
      // t[g1] = sum_g2 m_c[g1][g2] * A[x][g2]
      for (int g1 = 0; g1 < G; ++g1) {
        float acc = 0.f;
        const float* mr = m + c * P + g1 * G;
        for (int g2 = 0; g2 < G; ++g2) acc += mr[g2] * Ax[g2];
        t[g1 * HEAD_THREADS + tid] = acc;
      }
#pragma unroll
      for (int yy = 0; yy < BAND; ++yy) {
        if (yy < ny) {
          const float* Ay = A + (size_t)(y0 + yy) * G;
          float acc = 0.f;
          for (int g1 = 0; g1 < G; ++g1) acc += Ay[g1] * t[g1 * HEAD_THREADS + tid];
          if (!train) {
            maps[((size_t)b * S + (y0 + yy)) * S + x] = acc;
          } else if (c == 0) {
            o0[yy] = acc;
          } else {
            const float mx = fmaxf(o0[yy], acc);
            const float e0 = expf(o0[yy] - mx), e1 = expf(acc - mx);
            const float inv = 1.0f / (e0 + e1);
            float* base = maps + (((size_t)lvl * B + b) * 2) * S * S + (size_t)(y0 + yy) * S + x;
            base[0] = e0 * inv;
            base[(size_t)S * S] = e1 * inv;
          }
        }
      }
    }
  }
}

// scores[b] = (<det[b], anchors[:,1]> + 1) / 2      (test.py:83-84)
__global__ void scores_kernel(const float* __restrict__ det, const float* __restrict__ anchors, int anchors_batched,
                              int B, int E, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* T = anchors + (anchors_batched ? (size_t)b * E * 2 : 0);
  float acc = 0.f;
  for (int c = lane; c < E; c += 32) acc += det[(size_t)b * E + c] * T[c * 2 + 1];
  acc = ptx::warp_sum(acc);
  if (lane == 0) scores[b] = (acc + 1.0f) * 0.5f;
}


// forward_utils.py:155-161: rows L2-normalised, averaged, average L2-normalised; written to column `col`
// of anchors [width, 2].  One block; n is a handful of prompt sentences.
__global__ void __launch_bounds__(256)
text_anchor_kernel(const float* __restrict__ emb, int n, int width, float* __restrict__ anchors, int col) {
  extern __shared__ float sh[];  // [n] inverse row norms, then [8] partial sums
  float* inv = sh;
  float* part = sh + n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < n; r += 8) {
    float ss = 0.f;
    for (int c = lane; c < width; c += 32) { const float v = emb[(size_t)r * width + c]; ss += v * v; }
    ss = ptx::warp_sum(ss);
    if (lane == 0) inv[r] = 1.0f / sqrtf(ss);
  }
  __syncthreads();
  float local = 0.f;
  for (int c = threadIdx.x; c < width; c += 256) {
    float acc = 0.f;
    for (int r = 0; r < n; ++r) acc += emb[(size_t)r * width + c] * inv[r];
    acc /= float(n);
    anchors[c * 2 + col] = acc;  // un-normalised mean, fixed up below
    local += acc * acc;
  }
  local = ptx::warp_sum(local);
  if (lane == 0) part[warp] = local;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += part[i];
  const float invn = 1.0f / sqrtf(tot);
  for (int c = threadIdx.x; c < width; c += 256) anchors[c * 2 + col] *= invn;
}

}  // namespace

int k::launch_patch_dots(const void* const* seg, int n_levels, int seg_is_bf16, const float* anchors,
                         int anchors_batched, int B, int P, int E, float* dots, cudaStream_t stream) {
  if (n_levels < 1 || n_levels > MAX_LEVELS) return host::fail(host::ERR_INVALID, "head: n_levels=%d", n_levels);
  if (E % 8 != 0) return host::fail(host::ERR_INVALID, "head: embed dim %d must be a multiple of 8", E);
  SegPtrs sp;
  for (int l = 0; l < n_levels; ++l) {
    if (seg[l] == nullptr) return host::fail(host::ERR_INVALID, "head: seg[%d] is NULL", l);
    sp.p[l] = seg[l];
  }
  const int rows = B * P;
  const int blocks = (rows + 7) / 8;
  if (seg_is_bf16)
    patch_dots_kernel<true><<<blocks, 256, 0, stream>>>(sp, n_levels, anchors, anchors_batched, rows, P, E, dots);
  else
    patch_dots_kernel<false><<<blocks, 256, 0, stream>>>(sp, n_levels, anchors, anchors_batched, rows, P, E, dots);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

int k::launch_head_maps(const float* dots, int B, int G, int S, int mode, int n_levels, float* maps,
                        cudaStream_t stream) {
  if (mode < 0 || mode > 2) return host::fail(host::ERR_INVALID, "head: mode %d", mode);
  const int pad = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 3 : (mode == AACLIP_HEAD_TEST_MEDICAL ? 4 : 0);
  if (G <= pad) return host::fail(host::ERR_INVALID, "head: grid %d too small for reflect padding %d", G, pad);
  const size_t smem = ((size_t)S * G + 2 * G * G + (size_t)G * HEAD_THREADS) * sizeof(float);
  if (smem > 227 * 1024) return host::fail(host::ERR_INVALID, "head: img_size %d / grid %d needs %zu B smem", S, G, smem);
  static size_t configured = 0;
  if (smem > configured) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(head_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((S + BAND - 1) / BAND, B, mode == AACLIP_HEAD_TRAIN_SOFTMAX ? n_levels : 1);
  head_maps_kernel<<<grid, HEAD_THREADS, smem, stream>>>(dots, n_levels, B, G, S, mode, maps);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

int k::launch_scores(const float* det, const float* anchors, int anchors_batched, int B, int E, float* scores,
                     cudaStream_t stream) {
  scores_kernel<<<(B + 7) / 8, 256, 0, stream>>>(det, anchors, anchors_batched, B, E, scores);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

// Scratch for the standalone head entry (dots are tiny: n_levels * B * P * 2 floats).
namespace {
struct DotScratch { float* p = nullptr; size_t cap = 0; int dev = -1; };
DotScratch g_scratch;
}

extern "C" int aaclip_anomaly_head(const void* const* seg, int n_levels, int seg_is_bf16, const float* anchors,
                                   int anchors_batched, const float* det, int B, int P, int E, int img_size, int mode,
                                   float* maps_out, float* scores_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B <= 0) return host::OK;  // empty batch: nothing to do
  int G = 0;
  while ((G + 1) * (G + 1) <= P) ++G;  // H = int(sqrt(L)), forward_utils.py:201
  if (G * G != P) return host::fail(host::ERR_INVALID, "head: P=%d is not a square grid", P);
  if (maps_out != nullptr) {
    int dev = 0;
    AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
    const size_t need = (size_t)n_levels * B * P * 2 * sizeof(float);
    if (g_scratch.dev != dev || g_scratch.cap < need) {
      if (g_scratch.p) { cudaStreamSynchronize(stream); cudaFree(g_scratch.p); g_scratch.p = nullptr; }
      AACLIP_CUDA_CHECK(cudaMalloc(&g_scratch.p, need));
      g_scratch.cap = need; g_scratch.dev = dev;
    }
    int rc = k::launch_patch_dots(seg, n_levels, seg_is_bf16, anchors, anchors_batched, B, P, E, g_scratch.p, stream);
    if (rc) return rc;
    rc = k::launch_head_maps(g_scratch.p, B, G, img_size, mode, n_levels, maps_out, stream);
    if (rc) return rc;
  }
  if (scores_out != nullptr) {
    if (det == nullptr) return host::fail(host::ERR_INVALID, "head: scores requested without det");
    int rc = k::launch_scores(det, anchors, anchors_batched, B, E, scores_out, stream);
    if (rc) return rc;
  }
  return host::OK;
}

extern "C" int aaclip_text_anchor(const float* emb, int n, int width, float* anchors, int col, void* stream_) {
  if (!emb || !anchors || n <= 0 || width <= 0 || (col != 0 && col != 1))
    return host::fail(host::ERR_INVALID, "text_anchor: bad argument");
  text_anchor_kernel<<<1, 256, (n + 8) * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(emb, n, width, anchors, col);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}
