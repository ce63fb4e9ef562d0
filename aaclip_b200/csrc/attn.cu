// Multi-head self-attention core for head dim 64 on sm_100a (tcgen05 + TMEM + TMA), flash style: the
// [L, L] probability matrix the reference materialises (nn.MultiheadAttention slow path with
// need_weights=True, model/transformer.py:200,237) never leaves the SM.
//
// One CTA per (128-row query tile, head, image); 6 warps:
//   warp 0      TMA producer: Q tile once, then K_j / V_j tiles (128 keys each) through a 2-deep ring
//   warp 1      TMEM allocator + tcgen05.mma issuer:  S_j = Q K_j^T (M128 N128 K64),  O_j = P_j V_j (M128 N64 K128)
//   warps 2..5  softmax: thread == query row.  tcgen05.ld S_j, online max / exp2 / row-sum in fp32, P_j
//               written as bf16 into 128B-swizzled smem (the A operand of the PV MMA), O accumulated in
//               registers with the usual rescale, final O / l stored as bf16.
// Operands come straight out of the fused QKV GEMM output qkv[B*L, 3*heads*64] through ONE 2-D tensor map:
// rows past the image's last token are either the next image's tokens or TMA zero fill and are masked.
// V tiles are consumed as an MN-major B operand exactly as TMA lands them (no transpose pass).
#include <stdarg.h>
#include "common.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace attn {

constexpr int D = 64;          // head dim
constexpr int BQ = 128;        // query rows per CTA
constexpr int BKV = 128;       // keys per tile
constexpr int TILE_BYTES = 128 * D * 2;   // 16 KB: Q, K_j and V_j tiles
constexpr int P_BYTES = BQ * BKV * 2;     // 32 KB
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;            // S: [0,128), O: [128,192)
// smem: Q | K0 K1 | V0 V1 | P | barriers   (all tiles 1024-B aligned)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;
constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;

struct Bars {
  uint64_t q_full, k_full[2], v_full[2], k_empty[2], v_empty[2], s_full, p_full, o_full;
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, int L, int heads,
                 int causal) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-B alignment
  Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BAR);
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = ptx::lane_id();
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int W = heads * D;
  const int q0 = qt * BQ;
  int n_kv = (L + BKV - 1) / BKV;
  if (causal) n_kv = min(n_kv, qt + 1);
  const int row_base = b * L;  // first token row of this image in qkv / out

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQKV);
    ptx::mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->k_full[i], 1);
      ptx::mbar_init(&bars->v_full[i], 1);
      ptx::mbar_init(&bars->k_empty[i], 1);
      ptx::mbar_init(&bars->v_empty[i], 1);
    }
    ptx::mbar_init(&bars->s_full, 1);
    ptx::mbar_init(&bars->p_full, 4);  // one arrive per softmax warp
    ptx::mbar_init(&bars->o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<1>(&bars->tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_slot);

  if (warp == 0) {
    // ===================================================== TMA producer
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&bars->q_full, TILE_BYTES);
      ptx::tma_load_2d(smem + OFF_Q, &tmQKV, &bars->q_full, h * D, row_base + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(&bars->k_empty[st], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&bars->k_full[st], TILE_BYTES);
        ptx::tma_load_2d(smem + OFF_K + st * TILE_BYTES, &tmQKV, &bars->k_full[st], W + h * D, row_base + j * BKV);
        ptx::mbar_wait(&bars->v_empty[st], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&bars->v_full[st], TILE_BYTES);
        ptx::tma_load_2d(smem + OFF_V + st * TILE_BYTES, &tmQKV, &bars->v_full[st], 2 * W + h * D,
                         row_base + j * BKV);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(BQ, BKV, 0, 0);  // Q K-major, K K-major
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32(BQ, D, 0, 1);    // P K-major, V MN-major
      const uint32_t t_s = tmem_base, t_o = tmem_base + 128;
      const uint32_t q_addr = ptx::smem_u32(smem + OFF_Q);
      const uint32_t p_addr = ptx::smem_u32(smem + OFF_P);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        ptx::mbar_wait(&bars->k_full[st], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(smem + OFF_K + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::mma_f16_ss<1>(t_s, ptx::umma_desc_kmajor_sw128(q_addr + k * 32),
                             ptx::umma_desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
        ptx::mma_commit(&bars->s_full);
        ptx::mma_commit(&bars->k_empty[st]);
      };
      ptx::mbar_wait(&bars->q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        ptx::mbar_wait(&bars->p_full, j & 1);  // P_j in smem; S and O TMEM buffers drained
        ptx::tc_fence_after();
        if (j + 1 < n_kv) issue_s(j + 1);
        ptx::mbar_wait(&bars->v_full[st], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t v_addr = ptx::smem_u32(smem + OFF_V + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A: P is two 64-key swizzle atoms of 128 rows x 128 B; B: 16 keys = 16 rows x 128 B of the V tile
          const uint64_t ad = ptx::umma_desc_kmajor_sw128(p_addr + (k >> 2) * (BQ * 128) + (k & 3) * 32);
          const uint64_t bd = ptx::umma_desc_mnmajor_sw128(v_addr + k * 16 * 128, 1024);
          ptx::mma_f16_ss<1>(t_o, ad, bd, idesc_o, k != 0 ? 1u : 0u);
        }
        ptx::mma_commit(&bars->o_full);
        ptx::mma_commit(&bars->v_empty[st]);
      }
    }
  } else {
    // ===================================================== softmax warps: thread == query row
    const uint32_t quarter = warp & 3u;
    const int row = int(quarter * 32u + lane);
    const uint32_t t_s = tmem_base + ((quarter * 32u) << 16);
    const uint32_t t_o = t_s + 128;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    uint8_t* p_row = smem + OFF_P + row * 128;
    const uint32_t sw = uint32_t(row & 7);
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    const int qi = q0 + row;

    for (int j = 0; j < n_kv; ++j) {
      const int kv0 = j * BKV;
      const bool need_mask = (kv0 + BKV > L) || (causal && kv0 + BKV > q0);
      int limit = L - kv0;                          // keys >= limit are out of range
      if (causal) limit = min(limit, qi - kv0 + 1); // keys > qi are masked (CLIP text mask, model/model.py:172)
      ptx::mbar_wait(&bars->s_full, j & 1);
      ptx::tc_fence_after();
      // ---- pass 1: row max (two 32-column TMEM loads in flight per wait, 4 independent max chains)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int cp = 0; cp < 2; ++cp) {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32b_x32(t_s + cp * 64, v0);
        ptx::tmem_ld_32x32b_x32(t_s + cp * 64 + 32, v1);
        ptx::tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (cp * 64 + i < limit) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v0[i]));
            if (cp * 64 + 32 + i < limit) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v1[i]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            mx4[i & 3] = fmaxf(mx4[i & 3], fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
          }
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, mx);
      // a fully masked row (only rows >= L of a causal tile) keeps m_new = -inf: use 0 to avoid inf - inf
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ptx::ex2_approx((m - m_use) * c);
      // ---- fold in the previous tile's P.V now that its MMA has certainly been issued
      if (j > 0) {
        ptx::mbar_wait(&bars->o_full, (j - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_o + hh * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[hh * 32 + i] = fmaf(o[hh * 32 + i], alpha_prev, __uint_as_float(v[i]));
        }
      }
      // ---- pass 2: p = exp2((s - m) * c), bf16 P tile into swizzled smem, fp32 row sum (4 chains)
      const float mc = m_use * c;
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int cp = 0; cp < 2; ++cp) {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32b_x32(t_s + cp * 64, v0);
        ptx::tmem_ld_32x32b_x32(t_s + cp * 64 + 32, v1);
        ptx::tmem_ld_wait();
        // keys [cp*64, cp*64+64) = swizzle atom cp: v0 -> 16-B chunks 0..3, v1 -> chunks 4..7
        uint8_t* atom = p_row + cp * (BQ * 128);
#pragma unroll
        for (int hv = 0; hv < 2; ++hv) {
          const uint32_t (&v)[32] = hv ? v1 : v0;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0 = ptx::ex2_approx(fmaf(__uint_as_float(v[i]), c, -mc));
            float e1 = ptx::ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mc));
            if (need_mask) {
              if (cp * 64 + hv * 32 + i >= limit) e0 = 0.f;
              if (cp * 64 + hv * 32 + i + 1 >= limit) e1 = 0.f;
            }
            rs4[(i >> 1) & 3] += e0 + e1;
            pk[i >> 1] = ptx::pack_bf16x2(e0, e1);
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t chunk = uint32_t(hv * 4 + q4);
            ptx::st_shared_v4(atom + ((chunk ^ sw) << 4), pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        }
      }
      const float rs = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      l = fmaf(l, alpha, rs);
      m = m_new;
      alpha_prev = alpha;
      ptx::fence_proxy_async_smem();  // generic-proxy P stores -> visible to the tensor core (async proxy)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->p_full);
    }
    // ---- last tile's P.V
    ptx::mbar_wait(&bars->o_full, (n_kv - 1) & 1);
    ptx::tc_fence_after();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(t_o + hh * 32, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[hh * 32 + i] = fmaf(o[hh * 32 + i], alpha_prev, __uint_as_float(v[i]));
    }
    if (qi < L) {
      const float inv = 1.0f / l;
      uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(row_base + qi) * W + h * D);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        uint4 w;
        w.x = ptx::pack_bf16x2(o[q4 * 8 + 0] * inv, o[q4 * 8 + 1] * inv);
        w.y = ptx::pack_bf16x2(o[q4 * 8 + 2] * inv, o[q4 * 8 + 3] * inv);
        w.z = ptx::pack_bf16x2(o[q4 * 8 + 4] * inv, o[q4 * 8 + 5] * inv);
        w.w = ptx::pack_bf16x2(o[q4 * 8 + 6] * inv, o[q4 * 8 + 7] * inv);
        dst[q4] = w;
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

}  // namespace attn

int k::launch_attention(const void* qkv, void* out, int B, int L, int heads, int causal, cudaStream_t stream) {
  if (B <= 0) return host::OK;
  if (L <= 0 || heads <= 0) return host::fail(host::ERR_INVALID, "attention: L=%d heads=%d", L, heads);
  const int W = heads * attn::D;
  CUtensorMap tm;
  int rc = host::make_tmap_2d(&tm, qkv, (uint64_t)B * L, 3 * W, 3 * W, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(attn::attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           attn::SMEM_BYTES));
    configured = true;
  }
  dim3 grid((L + attn::BQ - 1) / attn::BQ, heads, B);
  attn::attention_kernel<<<grid, attn::THREADS, attn::SMEM_BYTES, stream>>>(
      tm, static_cast<__nv_bfloat16*>(out), L, heads, causal);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

extern "C" int aaclip_attention(const void* qkv, void* out, int B, int L, int heads, int causal, void* stream) {
  return k::launch_attention(qkv, out, B, L, heads, causal, static_cast<cudaStream_t>(stream));
}
