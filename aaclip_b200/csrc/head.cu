// Anomaly-map head: calculate_similarity_map (forward_utils.py:196-216) + the caller arithmetic of
// test.py:83-93 (level cat + sum, image score), as two HBM-bound kernels:
//
//   patch_dots_kernel : one warp per patch token; 16-byte coalesced loads of the (bf16 or fp32) normalised
//                       tokens of every level, warp-shuffle dot products with both text anchors.
//   head_maps_kernel  : per (image, band of output rows): level-summed patch map in smem, separable gaussian
//                       blur with reflect padding on the G x G grid, then the align_corners=True bilinear
//                       upsample as 4 taps per output pixel, float4 coalesced stores of the fp32 map.
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
// General form (any E % 8 == 0, shared or per-image anchors): one warp per row, levels in turn.
template <bool BF16>
__global__ void __launch_bounds__(256)
patch_dots_kernel(SegPtrs seg, int n_levels, const float* __restrict__ anchors, int anchors_batched, int rows, int P,
                  int E, float* __restrict__ dots) {
  ptx::grid_dep_sync();
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

// Streaming form for the A7 contract (bf16 tokens, E = 768, shared anchors): the lane's 24 anchor pairs live in
// registers for the whole kernel, each warp walks rows with a grid stride and issues the 16-byte loads of ALL
// levels of a row (up to 12 per lane) before it touches any of them, so enough bytes are in flight to run at HBM
// speed (the general form above had 3).
template <int NL>
__global__ void __launch_bounds__(256)
patch_dots_stream_kernel(SegPtrs seg, const float* __restrict__ anchors, int rows, float* __restrict__ dots) {
  ptx::grid_dep_sync();
  constexpr int E = 768, CH = E / 256;   // 3 chunks of 8 elements per lane
  const int lane = threadIdx.x & 31;
  float2 t[CH][8];                        // anchors of this lane's columns: (T[c][0], T[c][1])
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) t[i][j] = __ldg(reinterpret_cast<const float2*>(anchors + (size_t)(i * 256 + lane * 8 + j) * 2));
  const int warps = gridDim.x * 8;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += warps) {
    uint4 raw[NL][CH];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const __nv_bfloat16* f = static_cast<const __nv_bfloat16*>(seg.p[l]) + (size_t)row * E + lane * 8;
#pragma unroll
      for (int i = 0; i < CH; ++i) raw[l][i] = __ldcs(reinterpret_cast<const uint4*>(f + i * 256));   // streamed once
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[l][i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 v = __bfloat1622float2(h[j]);
          d0 += v.x * t[i][2 * j].x + v.y * t[i][2 * j + 1].x;
          d1 += v.x * t[i][2 * j].y + v.y * t[i][2 * j + 1].y;
        }
      }
      d0 = ptx::warp_sum(d0);
      d1 = ptx::warp_sum(d1);
      if (lane == 0) *reinterpret_cast<float2*>(dots + ((size_t)l * rows + row) * 2) = make_float2(d0, d1);
    }
  }
}

__device__ __forceinline__ int reflect_idx(int i, int G) {
  if (i < 0) i = -i;
  if (i >= G) i = 2 * (G - 1) - i;
  return i;
}

constexpr int BAND = 48;       // output rows per block
constexpr int HEAD_THREADS = 256;

// mode: AACLIP_HEAD_*.  dots: [n_levels][B*G*G][2].
// test modes : maps [B][S][S]; level-summed.  grid = (bands, B, 1)
// train mode : maps [n_levels][B][2][S][S]; softmax over the 2 channels. grid = (bands, B, n_levels)
//
// The blur (reflect padding) and the align_corners=True bilinear upsample are applied in the reference's order
// on the G x G patch map: blur rows, blur columns (kornia's separable filter), then each output pixel is the
// 4-tap bilinear blend of the blurred map, float4 stores along x.  Every block redoes the tiny blur of the whole
// G x G map (2 * G*G*ksize FMAs) rather than share it through global memory.
__global__ void __launch_bounds__(HEAD_THREADS)
head_maps_kernel(const float* __restrict__ dots, int n_levels, int B, int G, int S, int mode,
                 float* __restrict__ maps) {
  ptx::grid_dep_sync();
  extern __shared__ float sm[];
  const int P = G * G;
  float* m = sm;                 // [C][G][G]  per-patch scalars
  float* t = m + 2 * P;          // [C][G][G]  after the row blur
  float* mb = t + 2 * P;         // [C][G][G]  blurred map
  __shared__ float wk[16];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int lvl = blockIdx.z;
  const bool train = (mode == AACLIP_HEAD_TRAIN_SOFTMAX);
  const int ksize = train ? 1 : (mode == AACLIP_HEAD_TEST_INDUSTRIAL ? 7 : 9);
  const float sigma = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 1.0f : 1.5f;
  const size_t rows = (size_t)B * P;
  const int C = train ? 2 : 1;

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
  const float* src_map = m;
  if (!train) {
    // blur along x, then along y (reflect padding: forward_utils.py:208-210)
    const int half = ksize / 2;
    for (int i = tid; i < P; i += HEAD_THREADS) {
      const int gy = i / G, gx = i - gy * G;
      float acc = 0.f;
      for (int tt = 0; tt < ksize; ++tt) acc += wk[tt] * m[gy * G + reflect_idx(gx + tt - half, G)];
      t[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < P; i += HEAD_THREADS) {
      const int gy = i / G, gx = i - gy * G;
      float acc = 0.f;
      for (int tt = 0; tt < ksize; ++tt) acc += wk[tt] * t[reflect_idx(gy + tt - half, G) * G + gx];
      mb[i] = acc;
    }
    __syncthreads();
    src_map = mb;
  }

  // bilinear, align_corners=True: src = dst * (G-1)/(S-1)   (forward_utils.py:211-213)
  const float scale = (S > 1) ? float(G - 1) / float(S - 1) : 0.f;
  const int y0 = blockIdx.x * BAND;
  const int ny = min(BAND, S - y0);
  const int xq = (S + 3) / 4;               // float4 groups per row
  for (int i = tid; i < ny * xq; i += HEAD_THREADS) {
    const int yy = i / xq, x4 = (i - yy * xq) * 4;
    const int y = y0 + yy;
    const float sy = scale * float(y);
    int ry0 = min(int(sy), G - 1);
    const int ry1 = ry0 + ((ry0 < G - 1) ? 1 : 0);
    const float ly1 = sy - float(ry0), ly0 = 1.0f - ly1;
    float o[2][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int x = min(x4 + e, S - 1);
      const float sx = scale * float(x);
      int rx0 = min(int(sx), G - 1);
      const int rx1 = rx0 + ((rx0 < G - 1) ? 1 : 0);
      const float lx1 = sx - float(rx0), lx0 = 1.0f - lx1;
      for (int c = 0; c < C; ++c) {
        const float* mc = src_map + c * P;
        const float top = lx0 * mc[ry0 * G + rx0] + lx1 * mc[ry0 * G + rx1];
        const float bot = lx0 * mc[ry1 * G + rx0] + lx1 * mc[ry1 * G + rx1];
        o[c][e] = ly0 * top + ly1 * bot;
      }
    }
    if (!train) {
      float* dst = maps + ((size_t)b * S + y) * S + x4;
      if (x4 + 3 < S && (S & 3) == 0) *reinterpret_cast<float4*>(dst) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
      else for (int e = 0; e < 4 && x4 + e < S; ++e) dst[e] = o[0][e];
    } else {
      float p0[4], p1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float mx = fmaxf(o[0][e], o[1][e]);
        const float e0 = expf(o[0][e] - mx), e1 = expf(o[1][e] - mx);
        const float inv = 1.0f / (e0 + e1);
        p0[e] = e0 * inv; p1[e] = e1 * inv;
      }
      float* base = maps + (((size_t)lvl * B + b) * 2) * S * S + (size_t)y * S + x4;
      if (x4 + 3 < S && (S & 3) == 0) {
        *reinterpret_cast<float4*>(base) = make_float4(p0[0], p0[1], p0[2], p0[3]);
        *reinterpret_cast<float4*>(base + (size_t)S * S) = make_float4(p1[0], p1[1], p1[2], p1[3]);
      } else {
        for (int e = 0; e < 4 && x4 + e < S; ++e) { base[e] = p0[e]; base[(size_t)S * S + e] = p1[e]; }
      }
    }
  }
}

// ---- the A7 contract as ONE kernel (test modes, bf16 tokens, E = 768, shared anchors) -----------------------------
// A thread-block cluster of 8 CTAs per image.  Phase 1: CTA r reads patches [r*P/8, (r+1)*P/8) of all levels
// (every 16-byte load of a row issued before the first use; the lane's 24 anchor pairs stay in registers) and
// leaves the level-summed scalars m in ITS shared memory.  After a cluster barrier every CTA gathers the whole
// G x G map through distributed shared memory, blurs it, and writes 1/8 of the image's output rows with float4
// stores; CTA 0 also produces the image score.  Bytes moved = the algorithmic 3.99 MB/image, once.
constexpr int FUSED_CLUSTER = 8;
constexpr int FUSED_THREADS = 128;
template <int NL>
__global__ void __cluster_dims__(FUSED_CLUSTER, 1, 1) __launch_bounds__(FUSED_THREADS, 4)
head_fused_kernel(SegPtrs seg, const float* __restrict__ anchors, const float* __restrict__ det, int P, int G, int S,
                  int mode, float* __restrict__ maps, float* __restrict__ scores) {
  ptx::grid_dep_sync();
  constexpr int E = 768, CH = E / 256;
  extern __shared__ float sm[];
  float* m_part = sm;                 // [chunk]   this CTA's patches
  float* m = sm + 96;                 // [P]       whole map (gathered), chunk <= 96 keeps the offset fixed
  float* t = m + P;                   // [P]       after the row blur
  float* mb = t + P;                  // [P]       blurred map
  __shared__ float wk[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const int b = blockIdx.x / FUSED_CLUSTER;
  const int chunk = (P + FUSED_CLUSTER - 1) / FUSED_CLUSTER;
  const int ksize = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 7 : 9;
  const float sigma = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 1.0f : 1.5f;
  if (tid == 0) {
    float sum = 0.f;
    for (int i = 0; i < ksize; ++i) {
      const float x = float(i - ksize / 2);
      wk[i] = expf(-(x * x) / (2.0f * sigma * sigma));
      sum += wk[i];
    }
    for (int i = 0; i < ksize; ++i) wk[i] /= sum;
  }
  // ---- phase 1: anchor dots of this CTA's patches
  // the test-mode map needs only (s1 - s0) per level: ((100 d1) + 1 - (100 d0)) / 2 = 50 <f, T1 - T0> + 0.5, so a lane
  // keeps the 24 DIFFERENCES of its anchor pairs (half the registers and FMAs of two separate dot products)
  float td[CH][8];
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 t2 = __ldg(reinterpret_cast<const float2*>(anchors + (size_t)(i * 256 + lane * 8 + j) * 2));
      td[i][j] = t2.y - t2.x;
    }
  const int p_end = min(P, int(rank + 1) * chunk);
  // PP patches per warp and iteration: all PP * NL * CH 16-byte loads of a lane are in flight before the first use
  // (memory-level parallelism is what bounds this phase: 6 KB per warp and patch)
  constexpr int PP = 2, WARPS = FUSED_THREADS / 32;
  for (int p0 = int(rank) * chunk + warp; p0 < p_end; p0 += WARPS * PP) {
    uint4 raw[PP][NL][CH];
#pragma unroll
    for (int u = 0; u < PP; ++u) {
      const int p = min(p0 + u * WARPS, p_end - 1);   // a clamped duplicate is computed and dropped
      const size_t row = (size_t)b * P + p;
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const __nv_bfloat16* f = static_cast<const __nv_bfloat16*>(seg.p[l]) + row * E + lane * 8;
#pragma unroll
        for (int i = 0; i < CH; ++i) raw[u][l][i] = __ldcs(reinterpret_cast<const uint4*>(f + i * 256));
      }
    }
#pragma unroll
    for (int u = 0; u < PP; ++u) {
      float acc = 0.f;
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u][l][i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 v = __bfloat1622float2(h[j]);
            d = fmaf(v.x, td[i][2 * j], fmaf(v.y, td[i][2 * j + 1], d));
          }
        }
        d = ptx::warp_sum(d);
        acc += fmaf(50.0f, d, 0.5f);   // (s1 + 1 - s0) / 2 of this level (test.py:85)
      }
      const int p = p0 + u * WARPS;
      if (lane == 0 && p < p_end) m_part[p - int(rank) * chunk] = acc;
    }
  }
  // image score (test.py:83-84) by one warp of the cluster's first CTA
  if (rank == 0 && warp == 0 && scores != nullptr) {
    float acc = 0.f;
    for (int c = lane; c < E; c += 32) acc += det[(size_t)b * E + c] * __ldg(anchors + c * 2 + 1);
    acc = ptx::warp_sum(acc);
    if (lane == 0) scores[b] = (acc + 1.0f) * 0.5f;
  }
  ptx::cluster_sync();
  // ---- gather the whole map through distributed shared memory
  for (int i = tid; i < P; i += FUSED_THREADS) {
    const uint32_t src_rank = uint32_t(i / chunk);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(ptx::smem_u32(m_part + (i - int(src_rank) * chunk))), "r"(src_rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    m[i] = v;
  }
  ptx::cluster_sync();   // nobody's m_part may disappear (CTA exit) before every peer has read it
  // ---- blur along x, then along y (reflect padding: forward_utils.py:208-210)
  const int half = ksize / 2;
  for (int i = tid; i < P; i += FUSED_THREADS) {
    const int gy = i / G, gx = i - gy * G;
    float acc = 0.f;
    for (int k = 0; k < ksize; ++k) acc += wk[k] * m[gy * G + reflect_idx(gx + k - half, G)];
    t[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < P; i += FUSED_THREADS) {
    const int gy = i / G, gx = i - gy * G;
    float acc = 0.f;
    for (int k = 0; k < ksize; ++k) acc += wk[k] * t[reflect_idx(gy + k - half, G) * G + gx];
    mb[i] = acc;
  }
  __syncthreads();
  // ---- bilinear, align_corners=True (forward_utils.py:211-213): this CTA's share of the output rows
  const float scale = (S > 1) ? float(G - 1) / float(S - 1) : 0.f;
  const int rows_per = (S + FUSED_CLUSTER - 1) / FUSED_CLUSTER;
  const int y0 = int(rank) * rows_per, ny = max(0, min(rows_per, S - y0));
  const int xq = (S + 3) / 4;
  for (int i = tid; i < ny * xq; i += FUSED_THREADS) {
    const int yy = i / xq, x4 = (i - yy * xq) * 4;
    const int y = y0 + yy;
    const float sy = scale * float(y);
    const int ry0 = min(int(sy), G - 1);
    const int ry1 = ry0 + ((ry0 < G - 1) ? 1 : 0);
    const float ly1 = sy - float(ry0), ly0 = 1.0f - ly1;
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int x = min(x4 + e, S - 1);
      const float sx = scale * float(x);
      const int rx0 = min(int(sx), G - 1);
      const int rx1 = rx0 + ((rx0 < G - 1) ? 1 : 0);
      const float lx1 = sx - float(rx0), lx0 = 1.0f - lx1;
      const float top = lx0 * mb[ry0 * G + rx0] + lx1 * mb[ry0 * G + rx1];
      const float bot = lx0 * mb[ry1 * G + rx0] + lx1 * mb[ry1 * G + rx1];
      o[e] = ly0 * top + ly1 * bot;
    }
    float* dst = maps + ((size_t)b * S + y) * S + x4;
    if (x4 + 3 < S && (S & 3) == 0) __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
    else for (int e = 0; e < 4 && x4 + e < S; ++e) dst[e] = o[e];
  }
}

// scores[b] = (<det[b], anchors[:,1]> + 1) / 2      (test.py:83-84)
__global__ void scores_kernel(const float* __restrict__ det, const float* __restrict__ anchors, int anchors_batched,
                              int B, int E, float* __restrict__ scores) {
  ptx::grid_dep_sync();
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
  ptx::grid_dep_sync();
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
  if (seg_is_bf16 && !anchors_batched && E == 768 && (n_levels == 4 || n_levels == 1)) {
    // streaming form: each warp walks rows with a grid stride
    int dev = 0;
    AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
    const int sms = host::sm_count(dev) > 0 ? host::sm_count(dev) : 148;
    const int grid = std::min(blocks, sms * 2);   // 105 registers: two resident blocks per SM
    if (n_levels == 4)
      AACLIP_CUDA_CHECK(host::launch(patch_dots_stream_kernel<4>, dim3(grid), dim3(256), 0, stream, sp, anchors, rows, dots));
    else
      AACLIP_CUDA_CHECK(host::launch(patch_dots_stream_kernel<1>, dim3(grid), dim3(256), 0, stream, sp, anchors, rows, dots));
    return host::OK;
  }
  if (seg_is_bf16)
    AACLIP_CUDA_CHECK(host::launch(patch_dots_kernel<true>, dim3(blocks), dim3(256), 0, stream, sp, n_levels, anchors,
                                   anchors_batched, rows, P, E, dots));
  else
    AACLIP_CUDA_CHECK(host::launch(patch_dots_kernel<false>, dim3(blocks), dim3(256), 0, stream, sp, n_levels, anchors,
                                   anchors_batched, rows, P, E, dots));
  return host::OK;
}

int k::launch_head_maps(const float* dots, int B, int G, int S, int mode, int n_levels, float* maps,
                        cudaStream_t stream) {
  if (mode < 0 || mode > 2) return host::fail(host::ERR_INVALID, "head: mode %d", mode);
  const int pad = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 3 : (mode == AACLIP_HEAD_TEST_MEDICAL ? 4 : 0);
  if (G <= pad) return host::fail(host::ERR_INVALID, "head: grid %d too small for reflect padding %d", G, pad);
  const size_t smem = (size_t)6 * G * G * sizeof(float);
  if (smem > 200 * 1024) return host::fail(host::ERR_INVALID, "head: img_size %d / grid %d needs %zu B smem", S, G, smem);
  static size_t configured = 0;
  if (smem > configured) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(head_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((S + BAND - 1) / BAND, B, mode == AACLIP_HEAD_TRAIN_SOFTMAX ? n_levels : 1);
  AACLIP_CUDA_CHECK(host::launch(head_maps_kernel, grid, dim3(HEAD_THREADS), smem, stream, dots, n_levels, B, G, S, mode, maps));
  return host::OK;
}

int k::launch_scores(const float* det, const float* anchors, int anchors_batched, int B, int E, float* scores,
                     cudaStream_t stream) {
  AACLIP_CUDA_CHECK(host::launch(scores_kernel, dim3((B + 7) / 8), dim3(256), 0, stream, det, anchors, anchors_batched, B, E,
                                 scores));
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
  const bool test_mode = (mode == AACLIP_HEAD_TEST_INDUSTRIAL || mode == AACLIP_HEAD_TEST_MEDICAL);
  const int pad = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 3 : 4;
  if (maps_out != nullptr && test_mode && seg_is_bf16 && !anchors_batched && E == 768 && (n_levels == 4 || n_levels == 1) &&
      (P + FUSED_CLUSTER - 1) / FUSED_CLUSTER <= 96 && G > pad && (scores_out == nullptr || det != nullptr)) {
    // the whole head in one launch: cluster of 8 CTAs per image (head_fused_kernel)
    SegPtrs sp;
    for (int l = 0; l < n_levels; ++l) {
      if (seg[l] == nullptr) return host::fail(host::ERR_INVALID, "head: seg[%d] is NULL", l);
      sp.p[l] = seg[l];
    }
    const size_t smem = (96 + 3 * (size_t)P) * sizeof(float);
    if (n_levels == 4)
      AACLIP_CUDA_CHECK(host::launch(head_fused_kernel<4>, dim3(B * FUSED_CLUSTER), dim3(FUSED_THREADS), smem, stream, sp, anchors,
                                     det, P, G, img_size, mode, maps_out, scores_out));
    else
      AACLIP_CUDA_CHECK(host::launch(head_fused_kernel<1>, dim3(B * FUSED_CLUSTER), dim3(FUSED_THREADS), smem, stream, sp, anchors,
                                     det, P, G, img_size, mode, maps_out, scores_out));
    return host::OK;
  }
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

// ---- per-image extrema of the anomaly maps: what metrics_eval needs from the pixels for its image-level score
// (forward_utils.py:241-254: global min-max normalisation, then pmax = max over each image's pixels).  min / max are
// exact, so the host can combine per-image values over any number of batches.  HBM-bound: 4 * n_pix bytes per image.
namespace {
constexpr int MM_THREADS = 1024;
__global__ void __launch_bounds__(MM_THREADS) map_minmax_kernel(const float* __restrict__ maps, long long n_pix,
                                                                 float* __restrict__ out) {
  const float* m = maps + (long long)blockIdx.x * n_pix;
  float lo = INFINITY, hi = -INFINITY;
  const bool vec = ((reinterpret_cast<uintptr_t>(m) & 15u) == 0);
  const long long n4 = vec ? (n_pix >> 2) : 0;
  const float4* m4 = reinterpret_cast<const float4*>(m);
  for (long long i = threadIdx.x; i < n4; i += MM_THREADS) {
    const float4 v = __ldg(m4 + i);
    lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
    hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  for (long long i = (n4 << 2) + threadIdx.x; i < n_pix; i += MM_THREADS) {
    const float v = __ldg(m + i);
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[MM_THREADS / 32], s_hi[MM_THREADS / 32];
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = s_lo[threadIdx.x]; hi = s_hi[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = lo; out[2 * blockIdx.x + 1] = hi; }
  }
}
}  // namespace

extern "C" int aaclip_map_minmax(const float* maps, int B, long long n_pix, float* out, void* stream_) {
  if (B <= 0) return host::OK;
  if (!maps || !out || n_pix <= 0) return host::fail(host::ERR_INVALID, "map_minmax: bad argument");
  map_minmax_kernel<<<B, MM_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(maps, n_pix, out);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

extern "C" int aaclip_text_anchor(const float* emb, int n, int width, float* anchors, int col, void* stream_) {
  if (!emb || !anchors || n <= 0 || width <= 0 || (col != 0 && col != 1))
    return host::fail(host::ERR_INVALID, "text_anchor: bad argument");
  text_anchor_kernel<<<1, 256, (n + 8) * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(emb, n, width, anchors, col);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}
