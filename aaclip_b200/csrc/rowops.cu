// HBM-bound row kernels: one warp per row, 16-byte vectorised coalesced loads, warp-shuffle reductions,
// fp32 statistics.  They replace ATen's native_layer_norm / linalg_vector_norm / mul / div / add chains:
//   LayerNorm           model/transformer.py:37-43  (ln_pre, ln_1, ln_2, ln_post, ln_final)
//   adapter mix         model/adapter.py:93-99
//   F.normalize / mean  model/adapter.py:109-111
//   conv1 im2col, cls   model/adapter.py:68-82
#include <stdarg.h>
#include <algorithm>
#include "common.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace {

constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Row of VEC*128 floats held as VEC float4 per lane: element index = (i*32 + lane)*4 + {0..3}.
// The rows (the fp32 residual stream, 151 MB at B=64) are read once per kernel and are larger than L2: stream them
// (evict-first) so that they do not push out the bf16 output the next GEMM is about to read.
template <int VEC>
__device__ __forceinline__ void load_row(const float* row, int lane, float4 (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = __ldcs(reinterpret_cast<const float4*>(row + (i * 32 + lane) * 4));
}
template <int VEC>
__device__ __forceinline__ float row_sum(const float4 (&v)[VEC]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  return ptx::warp_sum(s);
}
template <int VEC>
__device__ __forceinline__ float row_sumsq(const float4 (&v)[VEC]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  return ptx::warp_sum(s);
}

// bf16 copy of the row + (sum, sum of squares) into slice 0 of the row's partial-sum record (other slices zeroed)
template <int VEC>
__device__ __forceinline__ void rowstats_store(const float4 (&v)[VEC], int lane, __nv_bfloat16* xb, float2* part,
                                               int slices) {
  const float s1 = row_sum<VEC>(v), s2 = row_sumsq<VEC>(v);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    uint2 w;
    w.x = ptx::pack_bf16x2(v[i].x, v[i].y);
    w.y = ptx::pack_bf16x2(v[i].z, v[i].w);
    *reinterpret_cast<uint2*>(xb + (i * 32 + lane) * 4) = w;
  }
  if (lane < slices) part[lane] = lane == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
}

// y = (x - mean) * rstd * gamma + beta written as bf16 (8 B per lane-chunk) and/or fp32
template <int VEC>
__device__ __forceinline__ void layernorm_store(const float4 (&v)[VEC], int lane, int width, const float* gamma,
                                                const float* beta, float eps, __nv_bfloat16* out_bf16,
                                                float* out_f32, float2* part_out = nullptr, int part_slices = 0) {
  const float inv_w = 1.0f / float(width);
  float o1 = 0.f, o2 = 0.f;   // sum / sum of squares of the OUTPUT row (folded-LayerNorm schedule, after ln_pre)
  const float mean = row_sum<VEC>(v) * inv_w;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(ptx::warp_sum(ss) * inv_w + eps);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = ldg4(gamma + col), b = ldg4(beta + col);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out_bf16) {
      uint2 w;
      w.x = ptx::pack_bf16x2(y.x, y.y);
      w.y = ptx::pack_bf16x2(y.z, y.w);
      *reinterpret_cast<uint2*>(out_bf16 + col) = w;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + col) = y;
    o1 += (y.x + y.y) + (y.z + y.w);
    o2 += (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
  }
  if (part_out != nullptr) {
    o1 = ptx::warp_sum(o1);
    o2 = ptx::warp_sum(o2);
    if (lane < part_slices) part_out[lane] = lane == 0 ? make_float2(o1, o2) : make_float2(0.f, 0.f);
  }
}

template <int VEC>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, int rows, int rows_per_group, int row_offset, long long group_stride,
                 __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, float2* __restrict__ part_out,
                 int part_slices) {
  ptx::grid_dep_sync();
  constexpr int width = VEC * 128;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int g = r / rows_per_group, within = r - g * rows_per_group;
  const float* src = x + (long long)g * group_stride + (long long)(within + row_offset) * width;
  float4 v[VEC];
  load_row<VEC>(src, lane, v);
  layernorm_store<VEC>(v, lane, width, gamma, beta, eps, out_bf16 ? out_bf16 + (size_t)r * width : nullptr,
                       out_f32 ? out_f32 + (size_t)r * width : nullptr,
                       part_out ? part_out + (size_t)r * part_slices : nullptr, part_slices);
}

template <int VEC, bool A_BF16>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
adapter_mix_kernel(float* __restrict__ x, const void* __restrict__ a_, float w, int rows,
                   const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta, float eps,
                   __nv_bfloat16* __restrict__ ln_out, __nv_bfloat16* __restrict__ xb_out, float2* __restrict__ part_out,
                   int part_slices) {
  ptx::grid_dep_sync();
  constexpr int width = VEC * 128;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* xr = x + (size_t)r * width;
  float4 xv[VEC], av[VEC];
  load_row<VEC>(xr, lane, xv);
  if constexpr (A_BF16) {   // the folded schedule keeps the adapter branch in bf16 (it enters x with weight w = 0.1)
    const __nv_bfloat16* ar = static_cast<const __nv_bfloat16*>(a_) + (size_t)r * width;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const uint2 q = __ldcs(reinterpret_cast<const uint2*>(ar + (i * 32 + lane) * 4));
      const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
      const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
      av[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
  } else {
    load_row<VEC>(static_cast<const float*>(a_) + (size_t)r * width, lane, av);
  }
  const float nx = sqrtf(row_sumsq<VEC>(xv));
  const float na = sqrtf(row_sumsq<VEC>(av));
  const float sa = w * (nx / na);  // no epsilon, as in the reference (model/adapter.py:94-98)
  const float sx = 1.0f - w;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    xv[i].x = sa * av[i].x + sx * xv[i].x;
    xv[i].y = sa * av[i].y + sx * xv[i].y;
    xv[i].z = sa * av[i].z + sx * xv[i].z;
    xv[i].w = sa * av[i].w + sx * xv[i].w;
    *reinterpret_cast<float4*>(xr + (i * 32 + lane) * 4) = xv[i];
  }
  if (ln_gamma != nullptr)
    layernorm_store<VEC>(xv, lane, width, ln_gamma, ln_beta, eps, ln_out + (size_t)r * width, nullptr);
  if (xb_out != nullptr) rowstats_store<VEC>(xv, lane, xb_out + (size_t)r * width, part_out + (size_t)r * part_slices, part_slices);
}

// Folded-LayerNorm schedule: the bf16 copy of a fp32 row + its (sum, sum of squares) in the producer's partial layout
// (whole-row sums in slice 0, zeros elsewhere).  Used where no residual GEMM produced them: after ln_pre.
template <int VEC>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
rowstats_cast_kernel(const float* __restrict__ x, int rows, __nv_bfloat16* __restrict__ xb, float2* __restrict__ part,
                     int part_slices) {
  ptx::grid_dep_sync();
  constexpr int width = VEC * 128;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  float4 v[VEC];
  load_row<VEC>(x + (size_t)r * width, lane, v);
  rowstats_store<VEC>(v, lane, xb + (size_t)r * width, part + (size_t)r * part_slices, part_slices);
}

// Fold LayerNorm's affine into the Linear that consumes it (one warp per output feature n):
//   Wf[n,k] = bf16(W[n,k] * gamma[k]);  colsum[n] = sum_k float(Wf[n,k])  (of the ROUNDED weights: it must cancel what
//   the tensor core sums);  bias_f[n] = b[n] + sum_k W[n,k] * beta[k]
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
fold_ln_weight_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int N, int K, __nv_bfloat16* __restrict__ Wf,
                      float* __restrict__ colsum, float* __restrict__ bias_f) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (n >= N) return;
  float s = 0.f, bb = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 w = ldg4(W + (size_t)n * K + k), g = ldg4(gamma + k), b = ldg4(beta + k);
    const __nv_bfloat16 f0 = __float2bfloat16(w.x * g.x), f1 = __float2bfloat16(w.y * g.y),
                        f2 = __float2bfloat16(w.z * g.z), f3 = __float2bfloat16(w.w * g.w);
    s += (__bfloat162float(f0) + __bfloat162float(f1)) + (__bfloat162float(f2) + __bfloat162float(f3));
    bb += (w.x * b.x + w.y * b.y) + (w.z * b.z + w.w * b.w);
    uint2 o;
    o.x = (uint32_t)__bfloat16_as_ushort(f0) | ((uint32_t)__bfloat16_as_ushort(f1) << 16);
    o.y = (uint32_t)__bfloat16_as_ushort(f2) | ((uint32_t)__bfloat16_as_ushort(f3) << 16);
    *reinterpret_cast<uint2*>(Wf + (size_t)n * K + k) = o;
  }
  s = ptx::warp_sum(s);
  bb = ptx::warp_sum(bb);
  if (lane == 0) { colsum[n] = s; bias_f[n] = (bias ? bias[n] : 0.f) + bb; }
}

// F.normalize(dim=-1, eps=1e-12) of s[r, col0:col0+width]; optionally also the two anchor dot products
// dots[r] = (<f, T[:,0]>, <f, T[:,1]>) of the normalised row with anchors T [width, 2]  (forward_utils.py:199)
template <int VEC>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_rows_kernel(const float* __restrict__ s, int ld, int col0, int rows, float* __restrict__ out_f32,
                   __nv_bfloat16* __restrict__ out_bf16, const float* __restrict__ anchors,
                   float* __restrict__ dots) {
  ptx::grid_dep_sync();
  constexpr int width = VEC * 128;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  float4 v[VEC];
  load_row<VEC>(s + (size_t)r * ld + col0, lane, v);
  const float inv = 1.0f / fmaxf(sqrtf(row_sumsq<VEC>(v)), 1e-12f);
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    float4 y = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)r * width + col) = y;
    if (out_bf16) {
      uint2 w;
      w.x = ptx::pack_bf16x2(y.x, y.y);
      w.y = ptx::pack_bf16x2(y.z, y.w);
      *reinterpret_cast<uint2*>(out_bf16 + (size_t)r * width + col) = w;
    }
    if (anchors) {
      const float4 t0 = ldg4(anchors + col * 2), t1 = ldg4(anchors + col * 2 + 4);
      d0 += (y.x * t0.x + y.y * t0.z) + (y.z * t1.x + y.w * t1.z);
      d1 += (y.x * t0.y + y.y * t0.w) + (y.z * t1.y + y.w * t1.w);
    }
  }
  if (anchors) {
    d0 = ptx::warp_sum(d0);
    d1 = ptx::warp_sum(d1);
    if (lane == 0) *reinterpret_cast<float2*>(dots + (size_t)r * 2) = make_float2(d0, d1);
  }
}

// det[b, :] = mean_p normalize(s[b*P + p, col0:col0+width])      (model/adapter.py:110-111), two passes:
//   row_inv_norm_kernel: one warp per row -> inv[r] = 1 / max(||row||, 1e-12)
//   det_mean_kernel    : block per (image, 128-column slice); each warp sums P/8 rows (one float4 per lane), the 8
//                        partial sums are combined through smem in a fixed order (deterministic, no atomics).
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
row_inv_norm_kernel(const float* __restrict__ s, int ld, int col0, int rows, int width, float* __restrict__ inv) {
  ptx::grid_dep_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* row = s + (size_t)r * ld + col0;
  float ss = 0.f;
  for (int c = lane * 4; c < width; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  ss = ptx::warp_sum(ss);
  if (lane == 0) inv[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
det_mean_kernel(const float* __restrict__ s, int ld, int col0, int P, int width, const float* __restrict__ inv,
                float* __restrict__ det) {
  ptx::grid_dep_sync();
  __shared__ float4 part[WARPS_PER_BLOCK][32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + lane * 4;
  const bool live = c < width;
  const float* base = s + (size_t)b * P * ld + col0 + c;
  const float* iv = inv + (size_t)b * P;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
#pragma unroll 4
    for (int p = warp; p < P; p += WARPS_PER_BLOCK) {
      const float4 v = *reinterpret_cast<const float4*>(base + (size_t)p * ld);
      const float w = __ldg(iv + p);
      acc.x += v.x * w; acc.y += v.y * w; acc.z += v.z * w; acc.w += v.w * w;
    }
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && live) {
    float4 t = part[0][lane];
#pragma unroll
    for (int w = 1; w < WARPS_PER_BLOCK; ++w) {
      const float4 q = part[w][lane];
      t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
    }
    const float invP = 1.0f / float(P);
    *reinterpret_cast<float4*>(det + (size_t)b * width + c) = make_float4(t.x * invP, t.y * invP, t.z * invP, t.w * invP);
  }
}

// dots[l][r] = (sum d0, sum d1) / max(sqrt(sum ss), 1e-12) over the 128-column partials the seg_proj GEMM epilogue left
__global__ void __launch_bounds__(256)
dots_finish_kernel(const float4* __restrict__ partials, long long total_rows, int n_slices, float2* __restrict__ dots) {
  ptx::grid_dep_sync();
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= total_rows) return;
  float ss = 0.f, d0 = 0.f, d1 = 0.f;
  for (int i = 0; i < n_slices; ++i) {
    const float4 p = partials[r * n_slices + i];
    ss += p.x; d0 += p.y; d1 += p.z;
  }
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  dots[r] = make_float2(d0 * inv, d1 * inv);
}

// image fp32 [B,3,S,S] -> A bf16 [B*G*G, Kpad]; k = c*ps*ps + i*ps + j (conv1.weight.view(width,-1) order), zero padded.
// One CTA per (image, patch row): the 3 x ps image rows it needs (3*ps*S floats, 56 KB at 336 px) are staged in shared
// memory with coalesced 16-byte loads, a k -> smem-offset table replaces the per-element divisions, and every warp store
// is 128 contiguous bytes of one output row.
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ img, int S, int ps, int G, int Kpad, __nv_bfloat16* __restrict__ out) {
  ptx::grid_dep_sync();
  extern __shared__ float im_sm[];          // [3][ps][S] floats, then Kpad ints
  const int K = 3 * ps * ps, band = ps * S;
  int* off = reinterpret_cast<int*>(im_sm + 3 * band);
  const int gy = blockIdx.x % G, b = blockIdx.x / G;
  for (int c = 0; c < 3; ++c) {
    const float* src = img + (((long long)b * 3 + c) * S + (long long)gy * ps) * S;   // ps consecutive rows: contiguous
    if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(band * sizeof(float))) & 15u) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(im_sm + c * band);
      for (int i = threadIdx.x; i < band / 4; i += blockDim.x) d4[i] = __ldcs(s4 + i);
    } else {
      for (int i = threadIdx.x; i < band; i += blockDim.x) im_sm[c * band + i] = __ldcs(src + i);
    }
  }
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    int o = -1;
    if (k < K) {
      const int c = k / (ps * ps), rem = k - c * ps * ps, i = rem / ps, j = rem - i * ps;
      o = c * band + i * S + j;
    }
    off[k] = o;
  }
  __syncthreads();
  const int kp_n = Kpad / 2;
  __nv_bfloat16* dst = out + ((long long)b * G * G + (long long)gy * G) * Kpad;
  for (int t = threadIdx.x; t < G * kp_n; t += blockDim.x) {
    const int gx = t / kp_n, kp = t - gx * kp_n;
    const int o0 = off[2 * kp], o1 = off[2 * kp + 1];
    const float v0 = o0 >= 0 ? im_sm[o0 + gx * ps] : 0.f, v1 = o1 >= 0 ? im_sm[o1 + gx * ps] : 0.f;
    *reinterpret_cast<uint32_t*>(dst + (long long)gx * Kpad + 2 * kp) = ptx::pack_bf16x2(v0, v1);
  }
}

__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                                int B, int L, int width) {
  ptx::grid_dep_sync();
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < width; c += blockDim.x) x[(size_t)b * L * width + c] = cls[c] + pos[c];
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n4) {
  ptx::grid_dep_sync();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n4;
       t += (long long)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(x + t * 4);
    uint2 w;
    w.x = ptx::pack_bf16x2(v.x, v.y);
    w.y = ptx::pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + t * 4) = w;
  }
}

inline int row_blocks(int rows) { return (rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK; }

#define DISPATCH_VEC(width, CALL)                                                              \
  switch ((width) / 128) {                                                                     \
    case 1: { constexpr int V = 1; CALL; } break;                                              \
    case 2: { constexpr int V = 2; CALL; } break;                                              \
    case 4: { constexpr int V = 4; CALL; } break;                                              \
    case 6: { constexpr int V = 6; CALL; } break;                                              \
    case 8: { constexpr int V = 8; CALL; } break;                                              \
    default: return host::fail(host::ERR_INVALID, "row kernel: unsupported width %d", (width)); \
  }

// tokens[b, p, :] += vec[b, :]   (train.py:85: `t + cls_token.unsqueeze(1)`); float4 lanes, E % 4 == 0
__global__ void add_image_vector_kernel(float* __restrict__ tokens, const float* __restrict__ vec, int P, int E4,
                                        long long total4) {
  ptx::grid_dep_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % E4);
    const long long b = i / ((long long)P * E4);
    float4 t = reinterpret_cast<float4*>(tokens)[i];
    const float4 a = reinterpret_cast<const float4*>(vec)[b * E4 + c];
    t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
    reinterpret_cast<float4*>(tokens)[i] = t;
  }
}

}  // namespace

int k::launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, int rows, int width,
                        int rows_per_group, int row_offset, long long group_stride, void* out_bf16, float* out_f32,
                        cudaStream_t stream, void* part_out, int part_slices) {
  if (rows <= 0) return host::OK;
  if (width % 128 != 0) return host::fail(host::ERR_INVALID, "layernorm: width %d must be a multiple of 128", width);
  if (part_out && (part_slices < 1 || part_slices > 32)) return host::fail(host::ERR_INVALID, "layernorm: %d slices", part_slices);
  if (rows_per_group <= 0) { rows_per_group = rows; row_offset = 0; group_stride = 0; }
  DISPATCH_VEC(width, AACLIP_CUDA_CHECK(host::launch(layernorm_kernel<V>, dim3(row_blocks(rows)), dim3(WARPS_PER_BLOCK * 32),
                                                     0, stream, x, gamma, beta, eps, rows, rows_per_group, row_offset,
                                                     group_stride, static_cast<__nv_bfloat16*>(out_bf16), out_f32,
                                                     static_cast<float2*>(part_out), part_slices)));
  return host::OK;
}

int k::launch_adapter_mix(float* x, const void* a, float w, int rows, int width, const float* ln_gamma,
                          const float* ln_beta, float eps, void* ln_out_bf16, cudaStream_t stream, void* xb_out,
                          void* part_out, int part_slices, bool a_is_bf16) {
  if (rows <= 0) return host::OK;
  if (width % 128 != 0) return host::fail(host::ERR_INVALID, "adapter_mix: width %d must be a multiple of 128", width);
  if (xb_out && (!part_out || part_slices < 1 || part_slices > 32))
    return host::fail(host::ERR_INVALID, "adapter_mix: xb_out needs part_out and 1..32 slices");
  if (a_is_bf16) {
    DISPATCH_VEC(width, AACLIP_CUDA_CHECK(host::launch(adapter_mix_kernel<V, true>, dim3(row_blocks(rows)),
                                                       dim3(WARPS_PER_BLOCK * 32), 0, stream, x, a, w, rows, ln_gamma, ln_beta,
                                                       eps, static_cast<__nv_bfloat16*>(ln_out_bf16),
                                                       static_cast<__nv_bfloat16*>(xb_out), static_cast<float2*>(part_out),
                                                       part_slices)));
  } else {
    DISPATCH_VEC(width, AACLIP_CUDA_CHECK(host::launch(adapter_mix_kernel<V, false>, dim3(row_blocks(rows)),
                                                       dim3(WARPS_PER_BLOCK * 32), 0, stream, x, a, w, rows, ln_gamma, ln_beta,
                                                       eps, static_cast<__nv_bfloat16*>(ln_out_bf16),
                                                       static_cast<__nv_bfloat16*>(xb_out), static_cast<float2*>(part_out),
                                                       part_slices)));
  }
  return host::OK;
}

int k::launch_rowstats_cast(const float* x, int rows, int width, void* xb, void* part, int part_slices, cudaStream_t stream) {
  if (rows <= 0) return host::OK;
  if (width % 128 != 0 || part_slices < 1 || part_slices > 32)
    return host::fail(host::ERR_INVALID, "rowstats_cast: width %d / slices %d", width, part_slices);
  DISPATCH_VEC(width, AACLIP_CUDA_CHECK(host::launch(rowstats_cast_kernel<V>, dim3(row_blocks(rows)),
                                                     dim3(WARPS_PER_BLOCK * 32), 0, stream, x, rows,
                                                     static_cast<__nv_bfloat16*>(xb), static_cast<float2*>(part),
                                                     part_slices)));
  return host::OK;
}

int k::launch_fold_ln_weight(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K, void* Wf,
                             float* colsum, float* bias_f, cudaStream_t stream) {
  if (N <= 0) return host::OK;
  if (K % 4 != 0) return host::fail(host::ERR_INVALID, "fold_ln_weight: K=%d must be a multiple of 4", K);
  fold_ln_weight_kernel<<<row_blocks(N), WARPS_PER_BLOCK * 32, 0, stream>>>(W, bias, gamma, beta, N, K,
                                                                         static_cast<__nv_bfloat16*>(Wf), colsum, bias_f);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

int k::launch_l2norm_rows(const float* s, int ld, int col0, int rows, int width, float* out_f32, void* out_bf16,
                          const float* anchors, float* dots, cudaStream_t stream) {
  if (rows <= 0) return host::OK;
  if (width % 128 != 0 || ld % 4 != 0 || col0 % 4 != 0)
    return host::fail(host::ERR_INVALID, "l2norm: width %d / ld %d / col0 %d alignment", width, ld, col0);
  DISPATCH_VEC(width, AACLIP_CUDA_CHECK(host::launch(l2norm_rows_kernel<V>, dim3(row_blocks(rows)),
                                                     dim3(WARPS_PER_BLOCK * 32), 0, stream, s, ld, col0, rows, out_f32,
                                                     static_cast<__nv_bfloat16*>(out_bf16), anchors, dots)));
  return host::OK;
}

int k::launch_det_mean(const float* s, int ld, int col0, int B, int P, int width, float* inv_scratch, float* det,
                       cudaStream_t stream) {
  if (B <= 0) return host::OK;
  if (width % 4 != 0 || ld % 4 != 0 || col0 % 4 != 0) return host::fail(host::ERR_INVALID, "det_mean: alignment");
  if (!inv_scratch) return host::fail(host::ERR_INVALID, "det_mean: scratch for %d row norms missing", B * P);
  AACLIP_CUDA_CHECK(host::launch(row_inv_norm_kernel, dim3(row_blocks(B * P)), dim3(WARPS_PER_BLOCK * 32), 0, stream, s, ld,
                                 col0, B * P, width, inv_scratch));
  dim3 grid(B, (width + 127) / 128);
  AACLIP_CUDA_CHECK(host::launch(det_mean_kernel, grid, dim3(WARPS_PER_BLOCK * 32), 0, stream, s, ld, col0, P, width,
                                 (const float*)inv_scratch, det));
  return host::OK;
}

int k::launch_add_image_vector(float* tokens, const float* vec, int B, int P, int E, cudaStream_t stream) {
  const long long total4 = (long long)B * P * E / 4;
  if (total4 <= 0) return host::OK;
  if (E % 4 != 0 || (reinterpret_cast<uintptr_t>(tokens) & 15u) != 0 || (reinterpret_cast<uintptr_t>(vec) & 15u) != 0)
    return host::fail(host::ERR_INVALID, "add_image_vector: E %d and the pointers must be 16-byte aligned", E);
  const unsigned grid = (unsigned)std::min<long long>((total4 + 255) / 256, 148 * 16);
  AACLIP_CUDA_CHECK(host::launch(add_image_vector_kernel, dim3(grid), dim3(256), 0, stream, tokens, vec, P, E / 4, total4));
  return host::OK;
}

int k::launch_dots_finish(const void* partials, int n_levels, int rows, int n_slices, float* dots, cudaStream_t stream) {
  const long long total = (long long)n_levels * rows;
  if (total <= 0) return host::OK;
  AACLIP_CUDA_CHECK(host::launch(dots_finish_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream,
                                 static_cast<const float4*>(partials), total, n_slices, reinterpret_cast<float2*>(dots)));
  return host::OK;
}

int k::launch_im2col(const float* image, int B, int S, int ps, int Kpad, void* out_bf16, cudaStream_t stream) {
  if (S % ps != 0 || Kpad % 8 != 0 || Kpad < 3 * ps * ps)
    return host::fail(host::ERR_INVALID, "im2col: S=%d ps=%d Kpad=%d", S, ps, Kpad);
  const int G = S / ps;
  if (B <= 0) return host::OK;
  const size_t smem = (size_t)3 * ps * S * sizeof(float) + (size_t)Kpad * sizeof(int);
  if (smem > 220 * 1024) return host::fail(host::ERR_INVALID, "im2col: %zu B of shared memory (S=%d ps=%d)", smem, S, ps);
  if (smem > 48 * 1024)
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AACLIP_CUDA_CHECK(host::launch(im2col_kernel, dim3(B * G), dim3(256), smem, stream, image, S, ps, G, Kpad,
                                 static_cast<__nv_bfloat16*>(out_bf16)));
  return host::OK;
}

int k::launch_cls_rows(float* x, const float* cls, const float* pos, int B, int L, int width, cudaStream_t stream) {
  AACLIP_CUDA_CHECK(host::launch(cls_rows_kernel, dim3(B), dim3(256), 0, stream, x, cls, pos, B, L, width));
  return host::OK;
}

int k::launch_cast_bf16(const float* x, void* out_bf16, long long n, cudaStream_t stream) {
  if (n % 4 != 0) return host::fail(host::ERR_INVALID, "cast_bf16: n %% 4 != 0");
  const long long n4 = n / 4;
  const int blocks = int(std::min<long long>((n4 + 255) / 256, 148LL * 16));
  AACLIP_CUDA_CHECK(host::launch(cast_bf16_kernel, dim3(blocks), dim3(256), 0, stream, x,
                                 static_cast<__nv_bfloat16*>(out_bf16), n4));
  return host::OK;
}

extern "C" int aaclip_layernorm(const float* x, const float* gamma, const float* beta, float eps, int rows, int width,
                                void* out_bf16, float* out_f32, void* stream) {
  host::PointerDeviceGuard dev_guard(x);
  return k::launch_layernorm(x, gamma, beta, eps, rows, width, 0, 0, 0, out_bf16, out_f32,
                             static_cast<cudaStream_t>(stream));
}

extern "C" int aaclip_adapter_mix(float* x, const float* a, float w, int rows, int width, void* stream) {
  host::PointerDeviceGuard dev_guard(x);
  return k::launch_adapter_mix(x, a, w, rows, width, nullptr, nullptr, 0.f, nullptr,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int aaclip_rowstats_cast(const float* x, int rows, int width, void* xb, void* part, int part_slices, void* stream) {
  host::PointerDeviceGuard dev_guard(x);
  return k::launch_rowstats_cast(x, rows, width, xb, part, part_slices, static_cast<cudaStream_t>(stream));
}
extern "C" int aaclip_fold_ln_weight(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K,
                                     void* Wf, float* colsum, float* bias_f, void* stream) {
  host::PointerDeviceGuard dev_guard(W);
  if (!W || !gamma || !beta || !Wf || !colsum || !bias_f) return host::fail(host::ERR_INVALID, "fold_ln_weight: null argument");
  return k::launch_fold_ln_weight(W, bias, gamma, beta, N, K, Wf, colsum, bias_f, static_cast<cudaStream_t>(stream));
}

// tokens fp32 [B, P, E] += vec fp32 [B, E] broadcast over the patches (train.py:85)
extern "C" int aaclip_add_image_vector(float* tokens, const float* vec, int B, int P, int E, void* stream) {
  if (B > 0 && (!tokens || !vec)) return host::fail(host::ERR_INVALID, "add_image_vector: null argument");
  host::PointerDeviceGuard dev_guard(tokens);
  return k::launch_add_image_vector(tokens, vec, B, P, E, static_cast<cudaStream_t>(stream));
}
