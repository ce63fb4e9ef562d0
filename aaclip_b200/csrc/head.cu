// Anomaly-map head: calculate_similarity_map (forward_utils.py:196-216) + the caller arithmetic of
// test.py:83-93 (level cat + sum, image score).
//
//   test modes, shared anchors, E = 768 (what test.py runs): head_stream_kernel (head_stream.cu), one launch.
//   fused engine path (dots already left by the seg_proj GEMM epilogue): maps_from_dots_kernel below - one CTA per
//                       image: level sum, blur, upsample, map rows, extrema, score (head_epilogue.cuh).
//   general form (train mode = softmax over the two channels per level, per-image anchors, other E):
//     patch_dots_kernel : one warp per patch token; 16-byte coalesced loads of the (bf16 or fp32) normalised
//                         tokens of every level, warp-shuffle dot products with both text anchors.
//     head_maps_kernel  : per (image, band of output rows): patch map in smem, optional separable gaussian blur
//                         with reflect padding, align_corners=True bilinear upsample as 4 taps per output pixel.
//
// Test mode, summed over levels (test.py:93):   m[h][w] = sum_l (100*(d1-d0) + 1)/2 = 50*sum_l(d1-d0) + n_levels/2
#include <stdarg.h>
#include <algorithm>
#include "common.cuh"
#include "head_epilogue.cuh"
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

// ---- test-mode image tails: per-patch scalars -> blur -> upsample -> map rows, extrema, image score.
// Input: either dots [n_levels][B*P][2] (fused engine path: left by the seg_proj GEMM epilogue + dots_finish; the level
// sum (s1 + 1 - s0) / 2 is formed here) or msum [B][P] (already level-summed: head_stream_kernel).  A cluster of R CTAs
// (8 warps each) shares one image: every CTA blurs the tiny G x G map itself and writes 1/R of the output rows; the
// cluster's extrema meet in rank 0 through distributed shared memory.  R = 1 when the batch alone fills the SMs.
__global__ void __launch_bounds__(headepi::THREADS)
maps_from_dots_kernel(const float* __restrict__ dots, const float* __restrict__ msum, int n_levels, int B, int G, int S,
                      int ksize, const headepi::Taps taps, int R, const float* __restrict__ det, const float* __restrict__ anchors, int E,
                      float* __restrict__ maps, float* __restrict__ scores, float* __restrict__ minmax) {
  extern __shared__ __align__(16) float epi_sm[];
  __shared__ __align__(8) float cl_red[2 * 8];   // rank 0: the cluster's partial (min, max) pairs
  const int P = G * G, tid = threadIdx.x;
  const int b = blockIdx.x / R, rank = blockIdx.x - b * R;
  const headepi::Smem e = headepi::carve(epi_sm, P, G, S);
  headepi::setup(e, tid, headepi::THREADS, G, S, ksize, taps);
  ptx::grid_dep_sync();   // everything above overlapped the producing kernel's tail (programmatic dependent launch)
  const size_t rows = (size_t)B * P;
  for (int i0 = tid; i0 < P; i0 += 4 * headepi::THREADS) {
    // four patches x all levels per thread with every load issued before the first use: the inputs are cold (HBM / L2
    // round trips of a microsecond each), a dependent chain of n_levels * P / 256 of them would dominate the kernel
    if (msum != nullptr) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q * headepi::THREADS;
        v[q] = (i < P) ? __ldcg(msum + (size_t)b * P + i) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q * headepi::THREADS;
        if (i < P) e.m[i] = v[q];
      }
    } else {
      float2 d[4][MAX_LEVELS];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q * headepi::THREADS;
#pragma unroll
        for (int l = 0; l < MAX_LEVELS; ++l)
          d[q][l] = (i < P && l < n_levels) ? __ldg(reinterpret_cast<const float2*>(dots + ((size_t)l * rows + (size_t)b * P + i) * 2))
                                            : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q * headepi::THREADS;
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < MAX_LEVELS; ++l)
          if (l < n_levels) acc += (100.0f * d[q][l].y + 1.0f - 100.0f * d[q][l].x) * 0.5f;   // per level as the reference (test.py:85)
        if (i < P) e.m[i] = acc;
      }
    }
  }
  if (rank == 0 && tid < 32 && scores != nullptr) {
    if (E == 768) headepi::image_score_unrolled<24>(det, anchors, b, tid, scores);
    else headepi::image_score(det, anchors, E, b, tid, scores);
  }
  __syncthreads();
  if (maps == nullptr) return;   // uniform over the cluster
  const int rows_per = (S + R - 1) / R;
  const int y0 = min(S, rank * rows_per), y1 = min(S, y0 + rows_per);
  float lo, hi;
  headepi::image(e, tid, b, G, S, ksize, y0, y1, maps, lo, hi, [] { __syncthreads(); });
  if (minmax == nullptr) return;
  headepi::block_minmax(e, tid, lo, hi, [] { __syncthreads(); });
  if (R == 1) {
    if (tid == 0) { minmax[2 * b] = lo; minmax[2 * b + 1] = hi; }
    return;
  }
  if (tid == 0) {   // my (min, max) -> rank 0's cl_red[rank]
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(ptx::smem_u32(&cl_red[2 * rank])), "r"(0));
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(lo), "f"(hi) : "memory");
  }
  ptx::cluster_sync();   // release / acquire: every partial has landed in rank 0's shared memory
  if (rank == 0 && tid == 0) {
    for (int r = 1; r < R; ++r) { lo = fminf(lo, cl_red[2 * r]); hi = fmaxf(hi, cl_red[2 * r + 1]); }
    minmax[2 * b] = lo;
    minmax[2 * b + 1] = hi;
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
  // per call: the attribute belongs to the current device's context (one process may hold several contexts)
  if (smem > 48 * 1024)
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(head_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((S + BAND - 1) / BAND, B, mode == AACLIP_HEAD_TRAIN_SOFTMAX ? n_levels : 1);
  AACLIP_CUDA_CHECK(host::launch(head_maps_kernel, grid, dim3(HEAD_THREADS), smem, stream, dots, n_levels, B, G, S, mode, maps));
  return host::OK;
}

// Test modes from precomputed dots (or level-summed scalars `msum`): blur -> upsample -> maps (+ extrema, + image
// scores), one launch.  pdl: chain to the preceding kernel of the stream by programmatic dependent launch.
int k::launch_maps_from_dots(const float* dots, const float* msum, int n_levels, int B, int G, int S, int mode, const float* det,
                             const float* anchors, int E, float* maps, float* scores, float* minmax, bool pdl,
                             cudaStream_t stream) {
  if (B <= 0) return host::OK;
  if (mode != AACLIP_HEAD_TEST_INDUSTRIAL && mode != AACLIP_HEAD_TEST_MEDICAL)
    return host::fail(host::ERR_INVALID, "head: maps_from_dots is test-mode only (mode %d)", mode);
  const int pad = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 3 : 4;
  if (G <= pad) return host::fail(host::ERR_INVALID, "head: grid %d too small for reflect padding %d", G, pad);
  if (scores != nullptr && (det == nullptr || anchors == nullptr))
    return host::fail(host::ERR_INVALID, "head: scores requested without det / anchors");
  if ((dots == nullptr) == (msum == nullptr)) return host::fail(host::ERR_INVALID, "head: exactly one of dots / msum");
  if (n_levels < 1 || n_levels > MAX_LEVELS) return host::fail(host::ERR_INVALID, "head: n_levels=%d", n_levels);
  const size_t smem = headepi::smem_floats(G * G, G, S) * sizeof(float);
  if (smem > 200 * 1024) return host::fail(host::ERR_INVALID, "head: img_size %d / grid %d needs %zu B smem", S, G, smem);
  if (smem > 48 * 1024)
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(maps_from_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  const int sms = host::sm_count(dev) > 0 ? host::sm_count(dev) : 148;
  // rows of one image over a cluster of R CTAs while the batch alone does not fill the SMs
  // the tail is latency bound (8 warps per CTA, short dependent chains), not store bound: split an image's rows over
  // R CTAs while that keeps the grid within ~2 CTAs per SM (measured at batch 64: R = 1 / 2 / 4 / 8 -> 67.7 / 59.3 /
  // 55.0 / 68.8 us for the whole head; every CTA repeats the table set-up, the gather and the blur)
  int R = 1;
  while (maps != nullptr && R < 4 && 2 * R * B <= 2 * sms) R *= 2;
  if (getenv("AACLIP_HEAD_R")) R = std::max(1, std::min(8, atoi(getenv("AACLIP_HEAD_R"))));   // diagnostics
  if (getenv("AACLIP_HEAD_NOPDL")) pdl = false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * R);
  cfg.blockDim = dim3(headepi::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (R > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = R; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl || host::pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const headepi::Taps taps = headepi::gaussian_taps(pad * 2 + 1, mode == AACLIP_HEAD_TEST_INDUSTRIAL ? 1.0f : 1.5f);
  AACLIP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, maps_from_dots_kernel, dots, msum, n_levels, B, G, S, pad * 2 + 1, taps, R, det,
                                       anchors, E, maps, scores, minmax));
  return host::OK;
}

int k::launch_scores(const float* det, const float* anchors, int anchors_batched, int B, int E, float* scores,
                     cudaStream_t stream) {
  AACLIP_CUDA_CHECK(host::launch(scores_kernel, dim3((B + 7) / 8), dim3(256), 0, stream, det, anchors, anchors_batched, B, E,
                                 scores));
  return host::OK;
}

// Device scratch the head needs from its caller (the library allocates nothing per call, so the entry is capturable
// and safe on any stream): the streaming kernel's per-patch scalars + counters, or the general form's dots.
extern "C" long long aaclip_anomaly_head_workspace_bytes(int n_levels, int B, int P) {
  if (n_levels < 1 || B < 0 || P < 0) return 0;
  const long long dots = (long long)n_levels * B * P * 2 * (long long)sizeof(float);
  return std::max(dots, k::head_stream_workspace_bytes(B, P)) + 256;
}

extern "C" int aaclip_anomaly_head(const void* const* seg, int n_levels, int seg_is_bf16, const float* anchors,
                                   int anchors_batched, const float* det, int B, int P, int E, int img_size, int mode,
                                   float* maps_out, float* scores_out, float* minmax_out, void* workspace,
                                   long long workspace_bytes, void* stream_) {
  host::PointerDeviceGuard dev_guard((seg != nullptr && n_levels > 0) ? seg[0] : nullptr);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B <= 0) return host::OK;  // empty batch: nothing to do
  if (!seg || !anchors || n_levels < 1 || n_levels > MAX_LEVELS)
    return host::fail(host::ERR_INVALID, "head: seg / anchors missing or n_levels=%d", n_levels);
  int G = 0;
  while ((G + 1) * (G + 1) <= P) ++G;  // H = int(sqrt(L)), forward_utils.py:201
  if (G * G != P) return host::fail(host::ERR_INVALID, "head: P=%d is not a square grid", P);
  if (scores_out != nullptr && det == nullptr) return host::fail(host::ERR_INVALID, "head: scores requested without det");
  for (int l = 0; l < n_levels; ++l)
    if (seg[l] == nullptr || (reinterpret_cast<uintptr_t>(seg[l]) & 15u) != 0)
      return host::fail(host::ERR_INVALID, "head: seg[%d] must be a 16-byte aligned device pointer", l);
  const bool test_mode = (mode == AACLIP_HEAD_TEST_INDUSTRIAL || mode == AACLIP_HEAD_TEST_MEDICAL);
  if (minmax_out != nullptr && (!test_mode || maps_out == nullptr))
    return host::fail(host::ERR_INVALID, "head: extrema are produced with the test-mode maps only");
  if (maps_out != nullptr) {
    if (workspace == nullptr || workspace_bytes < aaclip_anomaly_head_workspace_bytes(n_levels, B, P) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15u) != 0)
      return host::fail(host::ERR_INVALID, "head: workspace of %lld bytes (16-byte aligned) required, got %lld",
                        aaclip_anomaly_head_workspace_bytes(n_levels, B, P), workspace_bytes);
    if (k::head_stream_supported(n_levels, seg_is_bf16, anchors_batched, E, P, G, img_size, mode, seg)) {
      // tokens streamed once (head_stream.cu: one scalar per patch + the image scores), then the image tails, chained
      // by programmatic dependent launch: maps + extrema written once
      int rc = k::launch_head_stream(seg, n_levels, seg_is_bf16, anchors, det, B, P, scores_out, workspace, stream);
      if (rc) return rc;
      return k::launch_maps_from_dots(nullptr, static_cast<const float*>(workspace), n_levels, B, G, img_size, mode, nullptr,
                                      nullptr, E, maps_out, nullptr, minmax_out, true, stream);
    }
    float* dots = static_cast<float*>(workspace);
    int rc = k::launch_patch_dots(seg, n_levels, seg_is_bf16, anchors, anchors_batched, B, P, E, dots, stream);
    if (rc) return rc;
    if (test_mode && !anchors_batched)
      return k::launch_maps_from_dots(dots, nullptr, n_levels, B, G, img_size, mode, det, anchors, E, maps_out, scores_out,
                                      minmax_out, false, stream);
    if (minmax_out != nullptr) return host::fail(host::ERR_INVALID, "head: extrema need shared anchors");
    rc = k::launch_head_maps(dots, B, G, img_size, mode, n_levels, maps_out, stream);
    if (rc) return rc;
  }
  if (scores_out != nullptr) {
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
  host::PointerDeviceGuard dev_guard(maps);
  if (B <= 0) return host::OK;
  if (!maps || !out || n_pix <= 0) return host::fail(host::ERR_INVALID, "map_minmax: bad argument");
  map_minmax_kernel<<<B, MM_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(maps, n_pix, out);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

extern "C" int aaclip_text_anchor(const float* emb, int n, int width, float* anchors, int col, void* stream_) {
  host::PointerDeviceGuard dev_guard(emb);
  if (!emb || !anchors || n <= 0 || width <= 0 || (col != 0 && col != 1))
    return host::fail(host::ERR_INVALID, "text_anchor: bad argument");
  text_anchor_kernel<<<1, 256, (n + 8) * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(emb, n, width, anchors, col);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}
