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
// (token, head) item, the item's `batch` value rows staged in shared memory as fp32 (row pitch 65 floats: a lane
// reads its own query row without bank conflicts, a key row is a broadcast), lane = query image, two passes over the
// keys (row maximum, then exponentials and the weighted sum).  Rows are 128-byte bf16 segments of the GEMM output.
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
  const size_t smem = (size_t)VV_WARPS * B * VV_PITCH * sizeof(float);
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
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
