// v-v attention of the "surgery" feature extractor.
//
// Reference: `Attention.forward` (model/transformer.py:123-152) installed over the last DPAM_layer-1 blocks by
// `VisionTransformer.DAPM_replace` (model/transformer.py:406-425) and called by train.py:75 under no_grad.  The
// block hands the module its [L, batch, D] tensor and the module reads the shape as (B, N, C): what it computes is
//     v      = in_proj(x)[..., 2D:3D]                      (q and k only feed `attn_ori`, which is discarded)
//     A[l,h] = softmax_over_images( v[l,h] v[l,h]^T / sqrt(64) )          a  batch x batch  matrix
//     out    = out_proj( A[l,h] v[l,h] )
// i.e. for every token position and head, the images of ONE batch attend to each other.  It is batch-coupled: the
// result of an image depends on the other images of its batch, so a batch is never split into chunks.
//
// Work is tiny next to the GEMMs (L * heads * batch^2 * 64 * 4 flop; train.py runs batch 2): one warp per
// (token, head) item.  Batches of up to 8 images: lane = two channels, dot products by warp shuffles, no shared memory
// (vv_attention_small_kernel).  Larger batches: the item's value rows staged in shared memory as fp32 (row pitch 65
// floats: a lane reads its own query row without bank conflicts, a key row is a broadcast), lane = query image, two passes
// over the keys (row maximum, then exponentials and the weighted sum).  Rows are 128-byte bf16 segments of the GEMM output.
#include <stdarg.h>
#include <algorithm>
#include "common.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace {

typedef __nv_bfloat16 bf16;
constexpr int VV_D = 64;          // head dimension
constexpr int VV_PITCH = VV_D + 1;
constexpr int VV_WARPS = 4;
constexpr int VV_MAX_BATCH = 128;

__global__ void __launch_bounds__(VV_WARPS * 32)
vv_attention_kernel(const bf16* __restrict__ v, int ldv, bf16* __restrict__ out, int ldo, int B, int L, int heads,
                    float scale_log2e) {
  extern __shared__ float vv_smem[];
  ptx::grid_dep_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* vs = vv_smem + (size_t)warp * B * VV_PITCH;
  const long long items = (long long)L * heads;
  for (long long item = (long long)blockIdx.x * VV_WARPS + warp; item < items; item += (long long)gridDim.x * VV_WARPS) {
    const int l = int(item / heads), h = int(item % heads);
    // stage the B value rows of this (token, head): lane reads channels 2*lane, 2*lane+1 (one 128-byte row per step)
    for (int b = 0; b < B; ++b) {
      const __nv_bfloat162 p =
          *reinterpret_cast<const __nv_bfloat162*>(v + ((size_t)b * L + l) * ldv + h * VV_D + 2 * lane);
      const float2 f = __bfloat1622float2(p);
      vs[b * VV_PITCH + 2 * lane] = f.x;
      vs[b * VV_PITCH + 2 * lane + 1] = f.y;
    }
    __syncwarp();
    for (int q0 = 0; q0 < B; q0 += 32) {
      const int qb = q0 + lane;
      const bool active = qb < B;
      const float* qrow = vs + (active ? qb : 0) * VV_PITCH;
      float q[VV_D];
#pragma unroll
      for (int c = 0; c < VV_D; ++c) q[c] = qrow[c];
      float m = -INFINITY;
      for (int kb = 0; kb < B; ++kb) {
        const float* krow = vs + kb * VV_PITCH;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < VV_D; ++c) s = fmaf(q[c], krow[c], s);
        m = fmaxf(m, s);
      }
      float o[VV_D];
#pragma unroll
      for (int c = 0; c < VV_D; ++c) o[c] = 0.f;
      float sum = 0.f;
      for (int kb = 0; kb < B; ++kb) {
        const float* krow = vs + kb * VV_PITCH;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < VV_D; ++c) s = fmaf(q[c], krow[c], s);
        const float p = exp2f((s - m) * scale_log2e);
        sum += p;
#pragma unroll
        for (int c = 0; c < VV_D; ++c) o[c] = fmaf(p, krow[c], o[c]);
      }
      if (active) {
        const float inv = 1.f / sum;
        uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)qb * L + l) * ldo + h * VV_D);
#pragma unroll
        for (int c = 0; c < VV_D; c += 8) {
          uint4 w;
          __nv_bfloat162 t;
          t = __floats2bfloat162_rn(o[c] * inv, o[c + 1] * inv);     w.x = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(o[c + 2] * inv, o[c + 3] * inv); w.y = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(o[c + 4] * inv, o[c + 5] * inv); w.z = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(o[c + 6] * inv, o[c + 7] * inv); w.w = *reinterpret_cast<uint32_t*>(&t);
          dst[c / 8] = w;
        }
      }
    }
    __syncwarp();   // the rows are overwritten by the warp's next item
  }
}

// Small batches (train.py runs 2): the lane-per-query form above leaves 30 of 32 lanes idle and walks 64-long dependent
// FMA chains out of shared memory - 49 us per launch at B = 2, L = 1370.  Here a lane owns two CHANNELS of every image's
// value row (the bf16x2 it loads, one coalesced 128-byte row per image), the NB (NB + 1) / 2 distinct dot products (the
// score matrix is symmetric) are summed over the warp by butterfly shuffles, every lane evaluates the NB x NB softmax,
// and writes its two channels of the NB output rows.  No shared memory, ~40 registers.
constexpr int VV_SMALL_MAX = 8;
constexpr int VV_SMALL_WARPS = 8;

template <int NB>
__global__ void __launch_bounds__(VV_SMALL_WARPS * 32)
vv_attention_small_kernel(const bf16* __restrict__ v, int ldv, bf16* __restrict__ out, int ldo, int L, int heads,
                          float scale_log2e) {
  ptx::grid_dep_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long items = (long long)L * heads;
  for (long long item = (long long)blockIdx.x * VV_SMALL_WARPS + warp; item < items;
       item += (long long)gridDim.x * VV_SMALL_WARPS) {
    const int l = int(item / heads), h = int(item % heads);
    float2 x[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b)
      x[b] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(v + ((size_t)b * L + l) * ldv + h * VV_D + 2 * lane));
    float s[NB][NB];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int c = b; c < NB; ++c) {
        float d = fmaf(x[b].x, x[c].x, x[b].y * x[c].y);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        s[b][c] = d;
        s[c][b] = d;
      }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float m = s[b][0];
#pragma unroll
      for (int c = 1; c < NB; ++c) m = fmaxf(m, s[b][c]);
      float sum = 0.f, ox = 0.f, oy = 0.f;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const float p = exp2f((s[b][c] - m) * scale_log2e);
        sum += p;
        ox = fmaf(p, x[c].x, ox);
        oy = fmaf(p, x[c].y, oy);
      }
      const float inv = 1.f / sum;
      *reinterpret_cast<__nv_bfloat162*>(out + ((size_t)b * L + l) * ldo + h * VV_D + 2 * lane) =
          __floats2bfloat162_rn(ox * inv, oy * inv);
    }
  }
}

template <int NB>
cudaError_t launch_small(const bf16* v, int ldv, bf16* out, int ldo, int L, int heads, int sms, cudaStream_t stream) {
  const long long items = (long long)L * heads;
  const long long want = (items + VV_SMALL_WARPS - 1) / VV_SMALL_WARPS;
  const int grid = (int)std::min<long long>(want, 8LL * sms);
  return host::launch(vv_attention_small_kernel<NB>, dim3(grid), dim3(VV_SMALL_WARPS * 32), 0, stream, v, ldv, out, ldo, L, heads,
                      0.125f * 1.4426950408889634f);
}

}  // namespace

int k::vv_attention_max_batch() { return VV_MAX_BATCH; }

int k::launch_vv_attention(const void* v, int ldv, void* out, int ldo, int B, int L, int heads, cudaStream_t stream) {
  if (B <= 0) return host::OK;
  if (L <= 0 || heads <= 0) return host::fail(host::ERR_INVALID, "vv_attention: L=%d heads=%d", L, heads);
  if (B > VV_MAX_BATCH)
    return host::fail(host::ERR_INVALID, "vv_attention: batch %d; the v-v attention couples the images of a batch and "
                                         "handles at most %d of them", B, VV_MAX_BATCH);
  if (ldv % 8 != 0 || ldo % 8 != 0 || ldv < heads * VV_D || ldo < heads * VV_D ||
      (reinterpret_cast<uintptr_t>(v) & 15u) != 0 || (reinterpret_cast<uintptr_t>(out) & 15u) != 0)
    return host::fail(host::ERR_INVALID, "vv_attention: rows must be 16-byte aligned and hold heads * 64 values");
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  if (B <= VV_SMALL_MAX) {
    const int sm = host::sm_count(dev) > 0 ? host::sm_count(dev) : 148;
    const bf16* vp = static_cast<const bf16*>(v);
    bf16* op = static_cast<bf16*>(out);
    switch (B) {
#define VV_CASE(N_) case N_: AACLIP_CUDA_CHECK(launch_small<N_>(vp, ldv, op, ldo, L, heads, sm, stream)); break;
      VV_CASE(1) VV_CASE(2) VV_CASE(3) VV_CASE(4) VV_CASE(5) VV_CASE(6) VV_CASE(7) VV_CASE(8)
#undef VV_CASE
    }
    return host::OK;
  }
  const size_t smem = (size_t)VV_WARPS * B * VV_PITCH * sizeof(float);
  // the > 48 KB dynamic-smem opt-in is per (kernel, device): remembered per device
  static bool configured[64] = {false};
  if (smem > 48 * 1024 && !(dev >= 0 && dev < 64 && configured[dev])) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(vv_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           VV_WARPS * VV_MAX_BATCH * VV_PITCH * (int)sizeof(float)));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const long long items = (long long)L * heads;
  const int sms = host::sm_count(dev);
  const long long want = (items + VV_WARPS - 1) / VV_WARPS;
  const int grid = (int)std::min<long long>(want, 8LL * (sms > 0 ? sms : 148));
  AACLIP_CUDA_CHECK(host::launch(vv_attention_kernel, dim3(grid), dim3(VV_WARPS * 32), smem, stream,
                                 static_cast<const bf16*>(v), ldv, static_cast<bf16*>(out), ldo, B, L, heads,
                                 0.125f * 1.4426950408889634f));
  return host::OK;
}

// v: bf16 [B*L, ldv] with the value projection in columns [0, heads*64); out: bf16 [B*L, ldo].
extern "C" int aaclip_vv_attention(const void* v, int ldv, void* out, int ldo, int B, int L, int heads, void* stream) {
  if (B > 0 && (!v || !out)) return host::fail(host::ERR_INVALID, "vv_attention: null argument");
  host::PointerDeviceGuard dev_guard(v);
  return k::launch_vv_attention(v, ldv, out, ldo, B, L, heads, static_cast<cudaStream_t>(stream));
}
