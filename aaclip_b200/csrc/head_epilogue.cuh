// Per-image tail of the test-mode anomaly-map head, shared by head_stream_kernel (head_stream.cu) and
// maps_from_dots_kernel (head.cu):   G x G level-summed patch map  ->  separable gaussian blur, reflect padding
// (kornia gaussian_blur2d, forward_utils.py:208-210)  ->  bilinear upsample to S x S, align_corners=True
// (forward_utils.py:211-213)  ->  fp32 map rows + the image's (min, max)  (metrics_eval, forward_utils.py:241-252).
//
// 256 threads (8 warps) cooperate on one image.  The upsample is done row-wise: a warp blends the two source rows
// of output row y once (G values), then every output pixel is two shared-memory reads and two FMAs against a
// per-column (left index, right weight) table that is built once per kernel - 5x fewer instructions than the
// 4-tap-per-pixel form it replaces (which made the head issue bound).
#pragma once
#include <math.h>
#include "ptx.cuh"

namespace headepi {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

__device__ __forceinline__ int reflect_idx(int i, int G) {
  if (i < 0) i = -i;
  if (i >= G) i = 2 * (G - 1) - i;
  return i;
}

__host__ __device__ inline int padded_row(int G) { return (G + 1 + 31) & ~31; }   // one duplicate of the last column
// floats of shared memory the epilogue needs (m, t, mb; rx / lx tables; one blended row per warp; taps; reduction)
__host__ __device__ inline size_t smem_floats(int P, int G, int S) {
  return (size_t)3 * P + 2 * (size_t)((S + 3) & ~3) + (size_t)WARPS * 4 * padded_row(G) + 16 + 2 * WARPS;
}

struct Smem {
  float* m;     // [P]  level-summed per-patch scalars (input)
  float* t;     // [P]  after the blur along x
  float* mb;    // [P]  blurred map
  int* rx;      // [S]  left source column of output column x
  float* lx;    // [S]  weight of the right source column
  float* vbuf;  // [WARPS][2][padded_row(G)] float2: vertically blended source row of each warp as (v[g], v[g+1]) pairs,
                // double buffered
  float* wk;    // [16] gaussian taps
  float* red;   // [2 * WARPS] per-warp minima, maxima
};

__device__ __forceinline__ Smem carve(float* base, int P, int G, int S) {
  Smem e;
  const int S4 = (S + 3) & ~3;
  e.rx = reinterpret_cast<int*>(base);          // tables first: 16-byte aligned for the vector reads
  e.lx = base + S4;
  e.vbuf = e.lx + S4;
  e.m = e.vbuf + WARPS * 4 * padded_row(G);
  e.t = e.m + P;
  e.mb = e.t + P;
  e.wk = e.mb + P;
  e.red = e.wk + 16;
  return e;
}

// kornia 0.6.9 gaussian(): exp(-x^2 / (2 sigma^2)) over an odd window, normalised to sum 1 (host side: the taps travel
// as kernel arguments, so that no thread computes seven expf in front of everybody else)
struct Taps { float w[16]; };
inline Taps gaussian_taps(int ksize, float sigma) {
  Taps t = {};
  float sum = 0.f;
  for (int i = 0; i < ksize; ++i) {
    const float x = float(i - ksize / 2);
    t.w[i] = expf(-(x * x) / (2.0f * sigma * sigma));
    sum += t.w[i];
  }
  for (int i = 0; i < ksize; ++i) t.w[i] /= sum;
  return t;
}

// Column tables + gaussian taps; `tid` in [0, nthreads).  Caller synchronises afterwards.
__device__ __forceinline__ void setup(const Smem& e, int tid, int nthreads, int G, int S, int ksize, const Taps& taps) {
  const float scale = (S > 1) ? float(G - 1) / float(S - 1) : 0.f;   // area_pixel_compute_scale, align_corners
  for (int x = tid; x < S; x += nthreads) {
    const float sx = scale * float(x);
    const int rx0 = min(int(sx), G - 1);
    e.rx[x] = rx0;
    e.lx[x] = sx - float(rx0);
  }
  if (tid < ksize) e.wk[tid] = taps.w[tid];
}

// one output row from its vertically blended source row v: pixel x = lx0[x] v[rx[x]] + lx1[x] v[rx[x] + 1]
// (v[G] duplicates v[G-1]; its weight is 0 or an ulp).  Generic form: tables read from shared memory per row.
template <int VW>
__device__ __forceinline__ void upsample_row(const Smem& e, const float* v, float* dst, int lane, int S, float& lo,
                                             float& hi) {
  for (int q = lane; q * VW < S; q += 32) {
    float o[VW];
    int r[VW];
    float l[VW];
    if constexpr (VW == 4) {
      const int4 r4 = *reinterpret_cast<const int4*>(e.rx + 4 * q);
      const float4 l4 = *reinterpret_cast<const float4*>(e.lx + 4 * q);
      r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w;
      l[0] = l4.x; l[1] = l4.y; l[2] = l4.z; l[3] = l4.w;
    } else if constexpr (VW == 2) {
      const int2 r2 = *reinterpret_cast<const int2*>(e.rx + 2 * q);
      const float2 l2 = *reinterpret_cast<const float2*>(e.lx + 2 * q);
      r[0] = r2.x; r[1] = r2.y; l[0] = l2.x; l[1] = l2.y;
    } else {
      r[0] = e.rx[q]; l[0] = e.lx[q];
    }
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      o[k] = (1.0f - l[k]) * v[r[k]] + l[k] * v[r[k] + 1];
      lo = fminf(lo, o[k]);
      hi = fmaxf(hi, o[k]);
    }
    if constexpr (VW == 4) __stcs(reinterpret_cast<float4*>(dst + 4 * q), make_float4(o[0], o[1], o[2], o[3]));
    else if constexpr (VW == 2) __stcs(reinterpret_cast<float2*>(dst + 2 * q), make_float2(o[0], o[1]));
    else dst[q] = o[0];
  }
}

// vertically blended source row of output row y into v[0..G]
__device__ __forceinline__ void blend_row(const Smem& e, float* v, int y, float scale, int G, int lane) {
  const float sy = scale * float(y);
  const int ry0 = min(int(sy), G - 1);
  const int ry1 = ry0 + ((ry0 < G - 1) ? 1 : 0);
  const float ly1 = sy - float(ry0), ly0 = 1.0f - ly1;
  for (int gx = lane; gx <= G; gx += 32) {
    const int g = min(gx, G - 1);
    v[gx] = ly0 * e.mb[ry0 * G + g] + ly1 * e.mb[ry1 * G + g];
  }
}

// vertically blended source row of output row y as PAIRS: vp[g] = (v[g], v[g + 1]) for g < G (v[G] duplicates v[G-1]:
// its weight is 0 or an ulp), so that one 8-byte shared-memory read serves one output pixel
__device__ __forceinline__ void blend_row_pairs(const Smem& e, float2* vp, int y, float scale, int G, int lane) {
  const float sy = scale * float(y);
  const int ry0 = min(int(sy), G - 1);
  const int ry1 = ry0 + ((ry0 < G - 1) ? 1 : 0);
  const float ly1 = sy - float(ry0), ly0 = 1.0f - ly1;
  const float* m0 = e.mb + ry0 * G;
  const float* m1 = e.mb + ry1 * G;
  for (int g = lane; g < G; g += 32) {
    const int g1 = min(g + 1, G - 1);
    vp[g] = make_float2(ly0 * m0[g] + ly1 * m1[g], ly0 * m0[g1] + ly1 * m1[g1]);
  }
}

// Output rows y0 + warp, y0 + warp + WARPS, ... < y1 of one image by one warp.  The lane's NG column groups (VW pixels
// each) keep their (pair offset, weights) in registers for the whole image, the row blend is double buffered (one
// __syncwarp per row; the next row's blend is issued before this row's pixels), and every pixel is one 8-byte LDS +
// FMUL + FFMA + its share of a 3-input min / max: the loop is instruction bound, not store bound.
template <int VW, int NG>
__device__ __forceinline__ void upsample_rows_fast(const Smem& e, float* img, int warp, int lane, int G, int S, float scale,
                                                   int y0, int y1, float& lo, float& hi) {
  int off[NG][VW];
  float l1[NG][VW], l0[NG][VW];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int q = lane + 32 * g;
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      const int x = min(q * VW + k, S - 1);
      off[g][k] = e.rx[x] * 8;
      l1[g][k] = e.lx[x];
      l0[g][k] = 1.0f - l1[g][k];
    }
  }
  const int pr = padded_row(G);
  float2* vb = reinterpret_cast<float2*>(e.vbuf) + warp * 2 * pr;
  if (y0 + warp < y1) blend_row_pairs(e, vb, y0 + warp, scale, G, lane);
  int it = 0;
  for (int y = y0 + warp; y < y1; y += WARPS, ++it) {
    const char* v = reinterpret_cast<const char*>(vb + (it & 1) * pr);
    __syncwarp();   // this row's blend is visible; the other buffer's readers (previous row) are done
    if (y + WARPS < y1) blend_row_pairs(e, vb + ((it + 1) & 1) * pr, y + WARPS, scale, G, lane);
    float* dst = img + (size_t)y * S;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int q = lane + 32 * g;
      if (q * VW < S) {
        float o[VW];
#pragma unroll
        for (int k = 0; k < VW; ++k) {
          const float2 p = *reinterpret_cast<const float2*>(v + off[g][k]);
          o[k] = l0[g][k] * p.x + l1[g][k] * p.y;
        }
        if constexpr (VW == 4) {
          lo = fminf(fminf(lo, o[0]), fminf(o[1], fminf(o[2], o[3])));
          hi = fmaxf(fmaxf(hi, o[0]), fmaxf(o[1], fmaxf(o[2], o[3])));
          __stcs(reinterpret_cast<float4*>(dst + 4 * q), make_float4(o[0], o[1], o[2], o[3]));
        } else {
          lo = fminf(lo, fminf(o[0], o[1]));
          hi = fmaxf(hi, fmaxf(o[0], o[1]));
          __stcs(reinterpret_cast<float2*>(dst + 2 * q), make_float2(o[0], o[1]));
        }
      }
    }
  }
}

// e.m holds the image's per-patch scalars and a barrier has made them visible.  `sync` is the barrier over the 256
// cooperating threads.  Blurs the whole G x G map, then writes output rows [y0, y1) of image b (a cluster of CTAs may
// split the rows of one image); (lo, hi) return this thread's extrema over the pixels it wrote.
template <typename SyncFn>
__device__ __forceinline__ void image(const Smem& e, int tid, int b, int G, int S, int ksize, int y0, int y1,
                                      float* __restrict__ maps, float& lo, float& hi, SyncFn&& sync) {
  const int P = G * G, half = ksize / 2;
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < P; i += THREADS) {
    const int gy = i / G, gx = i - gy * G;
    float acc = 0.f;
    for (int k = 0; k < ksize; ++k) acc += e.wk[k] * e.m[gy * G + reflect_idx(gx + k - half, G)];
    e.t[i] = acc;
  }
  sync();
  for (int i = tid; i < P; i += THREADS) {
    const int gy = i / G, gx = i - gy * G;
    float acc = 0.f;
    for (int k = 0; k < ksize; ++k) acc += e.wk[k] * e.t[reflect_idx(gy + k - half, G) * G + gx];
    e.mb[i] = acc;
  }
  sync();
  const float scale = (S > 1) ? float(G - 1) / float(S - 1) : 0.f;
  float* img = maps + (size_t)b * S * S;
  const bool a16 = (reinterpret_cast<uintptr_t>(img) & 15u) == 0, a8 = (reinterpret_cast<uintptr_t>(img) & 7u) == 0;
  const int vw = ((S & 3) == 0 && a16) ? 4 : ((S & 1) == 0 && a8) ? 2 : 1;
  lo = INFINITY; hi = -INFINITY;
  if (vw == 4 && S <= 384) upsample_rows_fast<4, 3>(e, img, warp, lane, G, S, scale, y0, y1, lo, hi);        // 336 px
  else if (vw == 4 && S <= 512) upsample_rows_fast<4, 4>(e, img, warp, lane, G, S, scale, y0, y1, lo, hi);
  else if (vw == 2 && S <= 576) upsample_rows_fast<2, 9>(e, img, warp, lane, G, S, scale, y0, y1, lo, hi);   // 518 px
  else {
    float* v = e.vbuf + warp * 4 * padded_row(G);
    for (int y = y0 + warp; y < y1; y += WARPS) {
      blend_row(e, v, y, scale, G, lane);
      __syncwarp();
      float* dst = img + (size_t)y * S;
      if (vw == 4) upsample_row<4>(e, v, dst, lane, S, lo, hi);
      else if (vw == 2) upsample_row<2>(e, v, dst, lane, S, lo, hi);
      else upsample_row<1>(e, v, dst, lane, S, lo, hi);
      __syncwarp();
    }
  }
}

// block-wide (min, max) of the threads' (lo, hi): valid in thread 0 after the call
template <typename SyncFn>
__device__ __forceinline__ void block_minmax(const Smem& e, int tid, float& lo, float& hi, SyncFn&& sync) {
  const int warp = tid >> 5, lane = tid & 31;
  lo = -ptx::warp_max(-lo);
  hi = ptx::warp_max(hi);
  if (lane == 0) { e.red[warp] = lo; e.red[WARPS + warp] = hi; }
  sync();
  if (tid == 0)
    for (int w = 1; w < WARPS; ++w) { lo = fminf(lo, e.red[w]); hi = fmaxf(hi, e.red[WARPS + w]); }
}

// scores[b] = (<det[b], anchors[:,1]> + 1) / 2  by one warp      (test.py:83-84)
__device__ __forceinline__ void image_score(const float* __restrict__ det, const float* __restrict__ anchors, int E, int b,
                                            int lane, float* __restrict__ scores) {
  float acc = 0.f;
  for (int c = lane; c < E; c += 32) acc += det[(size_t)b * E + c] * __ldg(anchors + c * 2 + 1);
  acc = ptx::warp_sum(acc);
  if (lane == 0) scores[b] = (acc + 1.0f) * 0.5f;
}

// scores[b] = (<det[b], anchors[:,1]> + 1) / 2  by one warp      (test.py:83-84).  E = 32 * NE: all of a lane's
// loads are issued before the first use (one HBM round trip instead of NE).
template <int NE>
__device__ __forceinline__ void image_score_unrolled(const float* __restrict__ det, const float* __restrict__ anchors, int b,
                                                     int lane, float* __restrict__ scores) {
  float d[NE], t[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    d[i] = __ldg(det + (size_t)b * (32 * NE) + i * 32 + lane);
    t[i] = __ldg(anchors + (i * 32 + lane) * 2 + 1);
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NE; ++i) acc = fmaf(d[i], t[i], acc);
  acc = ptx::warp_sum(acc);
  if (lane == 0) scores[b] = (acc + 1.0f) * 0.5f;
}

}  // namespace headepi
