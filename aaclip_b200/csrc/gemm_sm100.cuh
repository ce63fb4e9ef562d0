// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   out = epilogue(A[M,K] . W[N,K]^T)
//
//   * A (activations) and W (nn.Linear weight, [out_features, in_features]) are bf16, K contiguous.
//   * TMA (cp.async.bulk.tensor, 128B swizzle) stages 64-wide K slabs into a multi-stage smem ring.
//   * One elected thread issues tcgen05.mma (kind::f16, fp32 accumulation) into TMEM; the accumulator
//     is double buffered (2 x 256 columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
//   * CG = 1: one CTA computes a 128x256 tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2)
//     computes a 256x256 tile; each CTA stages its own 128 rows of A and 128 rows of W, halving the
//     smem/L2 operand traffic per SM.
//   * 8 epilogue warps read TMEM with tcgen05.ld (lane == output row) and apply the fused epilogue:
//     bias, erf-GELU / QuickGELU / LeakyReLU.  Results are staged per warp in 128B-swizzled smem (32 rows x
//     128 B, bank-conflict free) and leave the SM as TMA tile stores (cp.async.bulk.tensor) or, for the fp32
//     residual stream, TMA reduce-adds (cp.reduce.async.bulk.tensor .add): x += tile happens in L2, the SM
//     never reads x, every global access is a full 128 B line, and ragged M/N tails are clipped by the TMA
//     unit.  The tiny patch-embedding GEMM keeps a direct register->global scatter (+positional embedding).
//
// These replace the cuBLAS/cuDNN + ATen elementwise call sites of the reference:
//   in_proj / out_proj   model/transformer.py:237 (nn.MultiheadAttention)      -> ACT_NONE + OUT_BF16 / OUT_F32_RESID
//   mlp.c_fc + gelu      model/transformer.py:211-217,257                      -> ACT_GELU_ERF|ACT_QUICK_GELU + OUT_BF16
//   mlp.c_proj + resid   model/transformer.py:257                              -> OUT_F32_RESID
//   SimpleAdapter        model/adapter_modules.py:6-13, model/adapter.py:92    -> ACT_LEAKY + OUT_F32
//   SimpleProj           model/adapter_modules.py:16-26, model/adapter.py:106  -> ACT_NONE|ACT_LEAKY + OUT_F32
//   conv1 + cls/pos      model/adapter.py:68-82                                -> OUT_F32_PATCH
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace gemm {

enum Act : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_QUICK_GELU = 2, ACT_LEAKY = 3 };
enum Out : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_RESID = 2, OUT_F32_PATCH = 3, OUT_DOTS = 4, OUT_F32_RESID_LN = 5 };

struct Args {
  int M, N, K;
  const float* bias;  // [N] or nullptr
  void* out;          // bf16 or f32, row pitch ldo elements
  int ldo;
  const float* pos;   // OUT_F32_PATCH: positional embedding [P+1, N]
  int P;              // OUT_F32_PATCH: patches per image
  // OUT_DOTS (seg_proj + F.normalize + anchor similarity, model/adapter.py:106-109 + forward_utils.py:199):
  // columns [0, dots_cols) are never stored; each 128-column slice of a row leaves (sum of squares, <row, T[:,0]>,
  // <row, T[:,1]>) in partials[row][slice]; columns >= dots_cols (det_proj on the last level) are stored as fp32.
  const float* anchors;   // [dots_cols, 2]
  float4* partials;       // [M][dots_cols / 128]
  int dots_cols;          // multiple of 128
  // LayerNorm folded into the CONSUMER GEMM (template LNF):  LN(x) W^T + b  ==  rstd_r (x (W o gamma)^T - mean_r s) + b'
  // with s[n] = sum_k (W o gamma)[n,k], b' = b + W beta.  A is the bf16 copy of the fp32 rows, W the folded weight,
  // bias = b'; the row statistics come from the partial sums the producer left per 128-column slice.
  const float2* ln_part;  // [M][ln_slices] (sum, sum of squares) of the fp32 rows
  int ln_slices;
  int ln_width;           // row width the statistics run over (= K)
  float ln_eps;
  const float* ln_colsum; // s[N]
  // OUT_F32_RESID_LN (the PRODUCER): out <- out + acc + bias (fp32, in place through TMA load / store), plus the bf16
  // copy of the new rows and their per-slice partial sums
  float2* part_out;       // [M][N / 128]
  __nv_bfloat16* xb;      // [M][ldxb] bf16 copy of the new rows
  int ldxb;
};

// RLN: the OUT_F32_RESID_LN epilogue (64 KB of x_old ring, one stage fewer); RV selects its variant (see
// epilogue_resid_ln): the working epilogue warps (8: two per TMEM lane quarter, 128 columns each; 4: one per quarter, all
// 256 columns) and the depth of each warp's x_old ring.
// BN_: tile N.  256 is the throughput shape; 128 (CG = 1 only) halves the work per tile and k-block for launches whose
// tiles would not fill the SMs (small batches: the K loop of one tile is a serial chain).
template <int CG, bool RLN = false, int RV = 0, int BN_ = 256>
struct Cfg {
  static_assert(BN_ == 256 || (BN_ == 128 && CG == 1), "tile N: 256, or 128 with cta_group 1");
  static constexpr int BM = 128;           // accumulator rows per CTA (== TMEM lanes)
  static constexpr int BN = BN_;           // tile N (per CTA pair when CG == 2)
  static constexpr int BN_CTA = BN / CG;   // rows of W staged by each CTA
  static constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
  static constexpr int UMMA_K = 16;
  static constexpr int STAGES = (BN_ == 128) ? (RLN ? 5 : 6) : RLN ? ((CG == 1) ? 3 : 5) : ((CG == 1) ? 4 : 6);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN_CTA * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int NUM_EPI_WARPS = 8;
  // RLN: only 4 of the 8 epilogue warps work (one per TMEM lane quarter, all 256 columns of its 32 rows), each with a
  // ring of XRING x_old chunks: what bounds that epilogue is bytes in flight per SM, not issue slots
  static constexpr int ACTIVE_EPI_WARPS = (BN_ == 128 || (RLN && RV != 1 && RV != 4)) ? 4 : 8;
  static constexpr int XRING = RV == 1 ? 1 : RV == 2 ? 3 : RV == 4 ? 2 : 4;
  static constexpr int STAGING_BYTES = (RLN ? 2 : 1) * 32 * 128;  // per epilogue warp: 32 rows x 128 B, 128B-swizzled
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + NUM_EPI_WARPS * STAGING_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
  static constexpr int THREADS = 128 + NUM_EPI_WARPS * 32;
  static constexpr int TMEM_COLS = 2 * BN; // 2 accumulator buffers x BN fp32 columns
};

// erf-GELU via Abramowitz-Stegun 7.1.26 (|erf err| <= 1.5e-7): 0.5x(1+erf(x/sqrt2)) = hx + |hx| * erf(|x|/sqrt2),
// erf(z) = 1 - t(a1 + t(a2 + t(a3 + t(a4 + t a5)))) exp(-z^2), t = 1/(1 + p z).  Branch free: one MUFU.RCP, one
// MUFU.EX2 and 11 FMA-pipe instructions per element (constants pre-folded onto x).
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float t = ptx::rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ptx::ex2_approx((x * x) * (-0.5f * 1.4426950408889634f));   // exp(-x^2/2)
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);
}
// x * sigmoid(1.702 x)  (model/transformer.py:46-49)
__device__ __forceinline__ float quick_gelu(float x) {
  return x * ptx::rcp_approx(1.0f + ptx::ex2_approx(-1.702f * 1.4426950408889634f * x));
}

// Packed (two fp32 lanes per instruction: FFMA2 / FMUL2 / FADD2) forms of the same functions: the c_fc epilogue is
// FMA-pipe bound (12 FMA-pipe instructions per element of GELU against an 8192-cycle mainloop per tile), the packed
// forms halve that.
__device__ __forceinline__ ptx::F2 gelu_erf2(ptx::F2 x) {
  using namespace ptx;
  const F2 ax = abs2(x);
  float d0, d1;
  f2_get(fma2(ax, f2s(0.3275911f * 0.70710678118654752f), f2s(1.0f)), d0, d1);
  const F2 t = f2(rcp_approx(d0), rcp_approx(d1));
  F2 p = fma2(t, f2s(1.061405429f), f2s(-1.453152027f));
  p = fma2(p, t, f2s(1.421413741f));
  p = fma2(p, t, f2s(-0.284496736f));
  p = fma2(p, t, f2s(0.254829592f));
  p = mul2(p, t);
  float e0, e1;
  f2_get(mul2(mul2(x, f2s(-0.5f * 1.4426950408889634f)), x), e0, e1);   // -x^2/2 * log2(e)
  const F2 e = f2(ex2_approx(e0), ex2_approx(e1));
  const F2 np = mul2(p, f2s(-1.0f));
  const F2 erf_abs = fma2(np, e, f2s(1.0f));
  const F2 hx = mul2(x, f2s(0.5f));
  return fma2(abs2(hx), erf_abs, hx);
}
__device__ __forceinline__ ptx::F2 quick_gelu2(ptx::F2 x) {
  using namespace ptx;
  float a0, a1;
  f2_get(mul2(x, f2s(-1.702f * 1.4426950408889634f)), a0, a1);
  float d0, d1;
  f2_get(add2(f2(ex2_approx(a0), ex2_approx(a1)), f2s(1.0f)), d0, d1);
  return mul2(x, f2(rcp_approx(d0), rcp_approx(d1)));
}
template <int ACT>
__device__ __forceinline__ ptx::F2 apply_act2(ptx::F2 x) {
  if constexpr (ACT == ACT_GELU_ERF) return gelu_erf2(x);
  else if constexpr (ACT == ACT_QUICK_GELU) return quick_gelu2(x);
  else if constexpr (ACT == ACT_LEAKY) {
    float a, b;
    ptx::f2_get(x, a, b);
    return ptx::f2(a > 0.f ? a : 0.01f * a, b > 0.f ? b : 0.01f * b);
  } else return x;
}

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (ACT == ACT_GELU_ERF) return gelu_erf(x);
  else if constexpr (ACT == ACT_QUICK_GELU) return quick_gelu(x);
  else if constexpr (ACT == ACT_LEAKY) return x > 0.f ? x : 0.01f * x;
  else return x;
}

template <int ACT, int OUT>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], int row, int col, const Args& a) {
  if (row >= a.M || col >= a.N) return;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (a.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(a.bias + col);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = apply_act<ACT>(f[j]);

  if constexpr (OUT == OUT_BF16) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + size_t(row) * a.ldo + col);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 w;
      w.x = ptx::pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
      w.y = ptx::pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
      w.z = ptx::pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
      w.w = ptx::pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
      o[j] = w;
    }
  } else {
    size_t orow = row;
    if constexpr (OUT == OUT_F32_PATCH) {
      // patch row m = b*P + p  ->  token row b*(P+1) + 1 + p, plus positional_embedding[1 + p]
      const int b = row / a.P, p = row - b * a.P;
      orow = size_t(b) * (a.P + 1) + 1 + p;
      const float4* p4 = reinterpret_cast<const float4*>(a.pos + size_t(p + 1) * a.N + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = __ldg(p4 + j);
        f[4 * j + 0] += q.x; f[4 * j + 1] += q.y; f[4 * j + 2] += q.z; f[4 * j + 3] += q.w;
      }
    }
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + orow * a.ldo + col);
    if constexpr (OUT == OUT_F32_RESID) {
      float4 r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = o[j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[4 * j + 0] += r[j].x; f[4 * j + 1] += r[j].y; f[4 * j + 2] += r[j].z; f[4 * j + 3] += r[j].w;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = make_float4(f[4 * j + 0], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  }
}

// bias + activation on 32 accumulator columns, two columns per instruction; LNF: the folded LayerNorm first,
//   f = rstd * acc + (-(mean * rstd) * s[col] + bias[col])
template <int ACT, int LNF = 0>
__device__ __forceinline__ void bias_act32(const uint32_t (&v)[32], float (&f)[32], const float* bias, int col,
                                           const float* colsum = nullptr, float mean = 0.f, float rstd = 1.f) {
  using namespace ptx;
  const F2 r2 = f2s(rstd), nmr2 = f2s(-mean * rstd);
  const float4* b4 = reinterpret_cast<const float4*>(bias + col);
  const float4* s4 = reinterpret_cast<const float4*>(colsum + col);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    F2 a0 = f2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]));
    F2 a1 = f2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    if constexpr (LNF != 0) {
      const float4 b = __ldg(b4 + j), s = __ldg(s4 + j);
      a0 = fma2(r2, a0, fma2(nmr2, f2(s.x, s.y), f2(b.x, b.y)));
      a1 = fma2(r2, a1, fma2(nmr2, f2(s.z, s.w), f2(b.z, b.w)));
    } else if (bias != nullptr) {
      const float4 b = __ldg(b4 + j);
      a0 = add2(a0, f2(b.x, b.y));
      a1 = add2(a1, f2(b.z, b.w));
    }
    a0 = apply_act2<ACT>(a0);
    a1 = apply_act2<ACT>(a1);
    f2_get(a0, f[4 * j + 0], f[4 * j + 1]);
    f2_get(a1, f[4 * j + 2], f[4 * j + 3]);
  }
}

// One epilogue warp drains its 32 rows x 128 accumulator columns: TMEM -> regs -> swizzled smem -> TMA.
// `stage` is this warp's private 4 KB buffer; lane == row.  Lane 0 owns the bulk async-groups.
template <int ACT, int OUT, int LNF = 0, typename ArriveFn>
__device__ __forceinline__ void epilogue_staged(uint32_t t_addr, uint8_t* stage, const CUtensorMap* tmC, int row0,
                                                int col0, uint32_t lane, const Args& a, ArriveFn&& release_tmem) {
  uint8_t* my_row = stage + lane * 128;
  const uint32_t sw = lane & 7u;
  float mean = 0.f, rstd = 1.f;
  if constexpr (LNF != 0) {   // this lane's row statistics from the producer's per-slice partial sums
    const int row = min(row0 + int(lane), a.M - 1);
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < a.ln_slices; ++i) {
      const float2 p = __ldg(a.ln_part + size_t(row) * a.ln_slices + i);
      s1 += p.x; s2 += p.y;
    }
    const float inv_n = 1.0f / float(a.ln_width);
    mean = s1 * inv_n;
    rstd = rsqrtf(fmaxf(fmaf(-mean, mean, s2 * inv_n), 0.f) + a.ln_eps);
  }
  constexpr int COLS_PER_STORE = (OUT == OUT_BF16) ? 64 : 32;   // 128 B per row either way
  constexpr int N_STORES = 128 / COLS_PER_STORE;
#pragma unroll 1
  for (int c = 0; c < N_STORES; ++c) {
    const int col = col0 + c * COLS_PER_STORE;
    uint32_t w[32];  // 128 B of output for this row
    if constexpr (OUT == OUT_BF16) {
      uint32_t v0[32], v1[32];
      ptx::tmem_ld_32x32b_x32(t_addr + c * 64, v0);
      ptx::tmem_ld_32x32b_x32(t_addr + c * 64 + 32, v1);
      ptx::tmem_ld_wait();
      if (c == N_STORES - 1) release_tmem();
      float f[32];
      bias_act32<ACT, LNF>(v0, f, a.bias, col, a.ln_colsum, mean, rstd);
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = ptx::pack_bf16x2(f[2 * j], f[2 * j + 1]);
      bias_act32<ACT, LNF>(v1, f, a.bias, col + 32, a.ln_colsum, mean, rstd);
#pragma unroll
      for (int j = 0; j < 16; ++j) w[16 + j] = ptx::pack_bf16x2(f[2 * j], f[2 * j + 1]);
    } else {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(t_addr + c * 32, v);
      ptx::tmem_ld_wait();
      if (c == N_STORES - 1) release_tmem();
      float f[32];
      bias_act32<ACT, LNF>(v, f, a.bias, col, a.ln_colsum, mean, rstd);
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(f[j]);
    }
    // the previous TMA store out of this buffer must have finished reading it
    if (lane == 0) ptx::bulk_wait_read<0>();
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
      ptx::st_shared_v4(my_row + ((uint32_t(q) ^ sw) << 4), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (col < a.N && row0 < a.M) {
        if constexpr (OUT == OUT_F32_RESID) ptx::tma_reduce_add_2d(tmC, stage, col, row0);
        else ptx::tma_store_2d(tmC, stage, col, row0);
      }
      ptx::bulk_commit();
    }
  }
}

// OUT_F32_RESID_LN epilogue:  x_new = x_old + acc + bias  for one warp's 32 rows x (NCH * 32) columns of a tile
// (lane == row), in chunks of 32 columns.  x_old is a 151 MB fp32 stream that lives in HBM: what bounds this epilogue is
// how many of its bytes are in flight, so every warp keeps a RING of XR 4 KB chunks (128B-swizzled TMA boxes) that runs
// XR - 1 chunks ahead of the arithmetic - across tile boundaries and ahead of the tile's accumulator (the loads do not
// depend on it).  A chunk is updated in place and leaves by TMA store; its slot is refilled one chunk later, when that
// store has finished reading it (cp.async.bulk.wait_group.read 1: does not wait in steady state).  The row's (sum, sum
// of squares) per 128 columns go to part_out.  The bf16 copy of x_new leaves either straight from registers (YM 0: 4 x
// 16 B, YM 1: 2 x 32 B per lane and chunk) or through a 4 KB staging buffer and a TMA store every second chunk (YM 2).
// History: one chunk in flight per warp (8 warps x 8 KB: x chunk + bf16 staging; kept as epilogue_resid_ln_v0) made
// every chunk pay a full HBM round trip: out_proj 110 us in-step against a 70 us HBM floor, DRAM 54 % busy.
template <int NCH, int XR, int YM, typename RequestFn, typename ArriveFn>
__device__ __forceinline__ void epilogue_resid_ln(uint32_t t_addr, uint8_t* xring, uint64_t* xbar, uint32_t& j,
                                                  RequestFn&& request, const CUtensorMap* tmC, const CUtensorMap* tmD,
                                                  int row0, int col0, uint32_t lane, const Args& a, ArriveFn&& release_tmem) {
  const uint32_t sw = lane & 7u;
  const int row = row0 + int(lane);
  [[maybe_unused]] uint8_t* Y = xring + XR * 4096;
  [[maybe_unused]] uint8_t* yrow = Y + lane * 128;
  float sum = 0.f, ssq = 0.f;
#pragma unroll 1
  for (int c = 0; c < NCH; ++c, ++j) {
    const int col = col0 + c * 32;
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(t_addr + c * 32, v);
    ptx::tmem_ld_wait();
    if (c == NCH - 1) release_tmem();
    float f[32];
    bias_act32<ACT_NONE>(v, f, a.bias, col);
    const uint32_t slot = j % uint32_t(XR);
    uint8_t* X = xring + slot * 4096;
    uint8_t* xrow = X + lane * 128;
    ptx::mbar_wait(&xbar[slot], (j / uint32_t(XR)) & 1u);   // x_old chunk j has landed (rows >= M: TMA zero fill)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float o0, o1, o2, o3;
      ptx::ld_shared_v4(xrow + ((uint32_t(q) ^ sw) << 4), o0, o1, o2, o3);
      f[4 * q + 0] += o0; f[4 * q + 1] += o1; f[4 * q + 2] += o2; f[4 * q + 3] += o3;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) { sum += f[i]; ssq = fmaf(f[i], f[i], ssq); }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      ptx::st_shared_v4(xrow + ((uint32_t(q) ^ sw) << 4), __float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]),
                        __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3]));
    uint32_t h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) h[i] = ptx::pack_bf16x2(f[2 * i], f[2 * i + 1]);
    if constexpr (YM == 2) {
      if ((c & 1) == 0) {   // the previous pair's Y store (in the group committed one chunk ago) must have read Y
        if (lane == 0) ptx::bulk_wait_read<0>();
        __syncwarp();
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        ptx::st_shared_v4(yrow + ((uint32_t((c & 1) * 4 + q) ^ sw) << 4), h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    } else if (row < a.M) {
      __nv_bfloat16* o = a.xb + size_t(row) * a.ldxb + col;
      if constexpr (YM == 1) {   // two full 32-byte sectors per lane
#pragma unroll
        for (int q = 0; q < 2; ++q)
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + 16 * q), "r"(h[8 * q]),
                       "r"(h[8 * q + 1]), "r"(h[8 * q + 2]), "r"(h[8 * q + 3]), "r"(h[8 * q + 4]), "r"(h[8 * q + 5]),
                       "r"(h[8 * q + 6]), "r"(h[8 * q + 7]) : "memory");
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<uint4*>(o)[q] = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      }
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (row0 < a.M) {   // rows past M are clipped by the TMA unit
        ptx::tma_store_2d(tmC, X, col, row0);
        if constexpr (YM == 2) { if (c & 1) ptx::tma_store_2d(tmD, Y, col - 32, row0); }
      }
      ptx::bulk_commit();
      if (j >= 1) ptx::bulk_wait_read<1>();   // the store of chunk j - 1 has finished reading slot (j + XR - 1) % XR
      request(j + uint32_t(XR) - 1u);
    }
    if ((c & 3) == 3) {
      if (row < a.M) a.part_out[size_t(row) * (a.N >> 7) + (col >> 7)] = make_float2(sum, ssq);
      sum = 0.f; ssq = 0.f;
    }
  }
}

// Round-1 form (RV 1), kept for A/B: 8 warps x 128 columns, ONE x_old chunk in flight per warp (X 4 KB + bf16 staging Y 4 KB).
template <typename ArriveFn>
__device__ __forceinline__ void epilogue_resid_ln_v0(uint32_t t_addr, uint8_t* stage, uint64_t* xbar, uint32_t& xphase,
                                                     const CUtensorMap* tmC, const CUtensorMap* tmD, int row0, int col0,
                                                     uint32_t lane, const Args& a, ArriveFn&& release_tmem) {
  uint8_t* X = stage;
  uint8_t* Y = stage + 4096;
  uint8_t* xrow = X + lane * 128;
  uint8_t* yrow = Y + lane * 128;
  const uint32_t sw = lane & 7u;
  const bool active = row0 < a.M;   // warp-uniform
  float sum = 0.f, ssq = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int col = col0 + c * 32;
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(t_addr + c * 32, v);
    ptx::tmem_ld_wait();
    if (c == 3) release_tmem();
    float f[32];
    bias_act32<ACT_NONE>(v, f, a.bias, col);
    if (active) { ptx::mbar_wait(xbar, xphase); xphase ^= 1u; }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float o0, o1, o2, o3;
      ptx::ld_shared_v4(xrow + ((uint32_t(q) ^ sw) << 4), o0, o1, o2, o3);
      f[4 * q + 0] += o0; f[4 * q + 1] += o1; f[4 * q + 2] += o2; f[4 * q + 3] += o3;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) { sum += f[i]; ssq = fmaf(f[i], f[i], ssq); }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      ptx::st_shared_v4(xrow + ((uint32_t(q) ^ sw) << 4), __float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]),
                        __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3]));
#pragma unroll
    for (int q = 0; q < 4; ++q)
      ptx::st_shared_v4(yrow + ((uint32_t((c & 1) * 4 + q) ^ sw) << 4), ptx::pack_bf16x2(f[8 * q], f[8 * q + 1]),
                        ptx::pack_bf16x2(f[8 * q + 2], f[8 * q + 3]), ptx::pack_bf16x2(f[8 * q + 4], f[8 * q + 5]),
                        ptx::pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (active && col < a.N) {
        ptx::tma_store_2d(tmC, X, col, row0);
        if (c & 1) ptx::tma_store_2d(tmD, Y, col0 + (c >> 1) * 64, row0);
      }
      ptx::bulk_commit();
      if (c < 3) {
        ptx::bulk_wait_read<0>();
        if (active) {
          ptx::mbar_arrive_expect_tx(xbar, 4096);
          ptx::tma_load_2d(X, tmC, xbar, col + 32, row0);
        }
      }
    }
    __syncwarp();
  }
  const int row = row0 + int(lane);
  if (row < a.M) a.part_out[size_t(row) * (a.N >> 7) + (col0 >> 7)] = make_float2(sum, ssq);
}
__device__ __forceinline__ void resid_ln_request_first(uint8_t* stage, uint64_t* xbar, const CUtensorMap* tmC, int row0,
                                                       int col0, uint32_t lane, const Args& a) {
  if (lane == 0) {
    ptx::bulk_wait_read<0>();
    if (row0 < a.M) {
      ptx::mbar_arrive_expect_tx(xbar, 4096);
      ptx::tma_load_2d(stage, tmC, xbar, col0, row0);
    }
  }
}

// OUT_DOTS epilogue of one warp: 32 rows x 128 accumulator columns -> three partial sums per row (lane == row).
template <int ACT, typename ArriveFn>
__device__ __forceinline__ void epilogue_dots(uint32_t t_addr, int row0, int col0, uint32_t lane, const Args& a,
                                              ArriveFn&& release_tmem) {
  float ss = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(t_addr + c * 32, v);
    ptx::tmem_ld_wait();
    if (c == 3) release_tmem();
    const int col = col0 + c * 32;
    float f[32];
    bias_act32<ACT>(v, f, a.bias, col);
    const float4* t4 = reinterpret_cast<const float4*>(a.anchors + size_t(col) * 2);   // same address in every lane
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 t = __ldg(t4 + j);   // T[col+2j][0..1], T[col+2j+1][0..1]
      ss = fmaf(f[2 * j], f[2 * j], fmaf(f[2 * j + 1], f[2 * j + 1], ss));
      d0 = fmaf(f[2 * j], t.x, fmaf(f[2 * j + 1], t.z, d0));
      d1 = fmaf(f[2 * j], t.y, fmaf(f[2 * j + 1], t.w, d1));
    }
  }
  const int row = row0 + int(lane);
  if (row < a.M) a.partials[size_t(row) * (a.dots_cols >> 7) + (col0 >> 7)] = make_float4(ss, d0, d1, 0.f);
}

template <int CG, int ACT, int OUT, int LNF = 0, int RV = 0, int BN = 256>
__global__ void __launch_bounds__(Cfg<CG>::THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, const Args args) {
  using C = Cfg<CG, OUT == OUT_F32_RESID_LN, RV, BN>;
  extern __shared__ uint8_t smem_raw[];
  // identical offset in both CTAs of a pair: the dynamic smem window starts at the same address
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;  // 1024-aligned: stage sizes are multiples of 1024
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + C::NUM_EPI_WARPS * C::STAGING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* xbar = tempty + 3;   // OUT_F32_RESID_LN: XRING x_old barriers per working epilogue warp
  static_assert((2 * C::STAGES + 4 + 1 + C::ACTIVE_EPI_WARPS * C::XRING) * 8 <= C::BAR_BYTES || OUT != OUT_F32_RESID_LN,
                "barrier block too small");

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = ptx::lane_id();
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);

  const int tiles_m = (args.M + C::BM * CG - 1) / (C::BM * CG);
  const int tiles_n = (args.N + C::BN - 1) / C::BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (args.K + C::BK - 1) / C::BK;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  // Warp roles: 0..7 epilogue, 8 TMA producer, 9 MMA issuer, 10 TMEM allocator, 11 idle.  The single-thread
  // control roles take the highest warp ids: the SMSP arbiter favours high warp ids, and a late TMA request or
  // MMA issue idles the tensor core, while these warps issue only a handful of instructions per k-block.
  if (warp == 8 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if constexpr (OUT != OUT_F32_PATCH) ptx::prefetch_tmap(&tmC);
    if constexpr (OUT == OUT_F32_RESID_LN && (RV == 1 || RV == 2)) ptx::prefetch_tmap(&tmD);
  }
  if (warp == 9 && ptx::elect_one()) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full[s], CG);   // leader's arrive.expect_tx (+ the peer producer's remote arrive)
      ptx::mbar_init(&empty[s], 1);   // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);                          // tcgen05.commit
      ptx::mbar_init(&tempty[a], CG * C::ACTIVE_EPI_WARPS);  // one arrive per working epilogue warp of the pair
    }
    if constexpr (OUT == OUT_F32_RESID_LN)
      for (int w = 0; w < C::ACTIVE_EPI_WARPS * C::XRING; ++w) ptx::mbar_init(&xbar[w], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 10) {
    ptx::tmem_alloc<CG>(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  ptx::grid_dep_sync();   // everything above overlapped the previous kernel's tail

  if (warp == 8) {
    // ===================================================== TMA producer
    if (ptx::elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
        const int m0 = m_blk * C::BM * CG + int(cta_rank) * C::BM;
        const int n0 = n_blk * C::BN + int(cta_rank) * C::BN_CTA;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          void* a_dst = sA + s * C::A_BYTES;
          void* b_dst = sB + s * C::B_BYTES;
          if constexpr (CG == 1) {
            ptx::mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
            ptx::tma_load_2d(a_dst, &tmA, &full[s], kb * C::BK, m0);
            ptx::tma_load_2d(b_dst, &tmB, &full[s], kb * C::BK, n0);
          } else {
            if (leader) ptx::mbar_arrive_expect_tx(&full[s], 2 * C::STAGE_BYTES);
            ptx::tma_load_2d_cg2(a_dst, &tmA, &full[s], kb * C::BK, m0);
            ptx::tma_load_2d_cg2(b_dst, &tmB, &full[s], kb * C::BK, n0);
            if (!leader) ptx::mbar_arrive_cluster(&full[s], 0);
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 9) {
    // ===================================================== MMA issuer (pair leader only)
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(C::BM * CG, C::BN, 0, 0);
      int s = 0; uint32_t ph = 0; uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
        ptx::mbar_wait(&tempty[acc], aph ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + s * C::A_BYTES);
          const uint32_t b_addr = ptx::smem_u32(sB + s * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < C::BK / C::UMMA_K; ++k) {
            const uint64_t ad = ptx::umma_desc_kmajor_sw128(a_addr + k * C::UMMA_K * 2);
            const uint64_t bd = ptx::umma_desc_kmajor_sw128(b_addr + k * C::UMMA_K * 2);
            ptx::mma_f16_ss<CG>(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 1) ptx::mma_commit(&empty[s]); else ptx::mma_commit_cg2(&empty[s], 0x3);
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
        if constexpr (CG == 1) ptx::mma_commit(&tfull[acc]); else ptx::mma_commit_cg2(&tfull[acc], 0x3);
      }
    }
  } else if (warp < uint32_t(C::ACTIVE_EPI_WARPS)) {
    // ===================================================== epilogue: TMEM -> regs -> (smem -> TMA | global)
    const uint32_t q = warp & 3u;            // TMEM lane quarter this warp may read
    const uint32_t half = warp >> 2;         // which 128 accumulator columns (RLN: 0 - the warp takes all 256)
    uint8_t* stage = staging + warp * (C::NUM_EPI_WARPS / C::ACTIVE_EPI_WARPS) * C::STAGING_BYTES;
    uint32_t it = 0;
    // RLN: x_old chunk jj of this warp's chunk stream (NCH per tile, tiles in this CTA's order) -> ring slot jj % XRING
    constexpr int NCH = C::BN / (8 * C::ACTIVE_EPI_WARPS);   // 32-column chunks per warp and tile (BN 256: 8 with 4 warps, 4 with 8)
    [[maybe_unused]] uint32_t xj = 0, xphase = 0;
    [[maybe_unused]] uint64_t* xb = &xbar[warp * C::XRING];
    [[maybe_unused]] auto request_x = [&](uint32_t jj) {
      const int t = cluster_id + int(jj / uint32_t(NCH)) * num_clusters;
      if (t >= num_tiles) return;
      const int mb = t / tiles_n, nb = t - mb * tiles_n;
      const uint32_t slot = jj % uint32_t(C::XRING);
      ptx::mbar_arrive_expect_tx(&xb[slot], 4096);
      ptx::tma_load_2d(stage + slot * 4096, &tmC, &xb[slot], nb * C::BN + int(half) * 128 + int(jj % uint32_t(NCH)) * 32,
                       mb * C::BM * CG + int(cta_rank) * C::BM + int(q * 32u));
    };
    if constexpr (OUT == OUT_F32_RESID_LN && RV != 1) {
      if (lane == 0) for (uint32_t jj = 0; jj + 1 < uint32_t(C::XRING); ++jj) request_x(jj);
    }
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
      const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
      const int row0 = m_blk * C::BM * CG + int(cta_rank) * C::BM + int(q * 32u);
      if constexpr (OUT == OUT_F32_RESID_LN && RV == 1)
        resid_ln_request_first(stage, xb, &tmC, row0, n_blk * C::BN + int(half) * 128, lane, args);
      ptx::mbar_wait(&tfull[acc], aph);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((q * 32u) << 16) + acc * C::BN + half * 128u;
      // every TMEM read of this accumulator buffer is done: hand it back to the MMA warp
      auto release_tmem = [&]() {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) ptx::mbar_arrive(&tempty[acc]); else ptx::mbar_arrive_cluster(&tempty[acc], 0);
        }
      };
      if constexpr (OUT == OUT_F32_PATCH) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_addr + c * 32, v);
          ptx::tmem_ld_wait();
          if (c == 3) release_tmem();
          epilogue_chunk<ACT, OUT>(v, row0 + int(lane), n_blk * C::BN + int(half) * 128 + c * 32, args);
        }
      } else if constexpr (OUT == OUT_DOTS) {
        const int col0 = n_blk * C::BN + int(half) * 128;
        if (col0 < args.dots_cols) epilogue_dots<ACT>(t_addr, row0, col0, lane, args, release_tmem);
        else epilogue_staged<ACT, OUT_F32>(t_addr, stage, &tmC, row0, col0, lane, args, release_tmem);
      } else if constexpr (OUT == OUT_F32_RESID_LN) {
        if constexpr (RV == 1)
          epilogue_resid_ln_v0(t_addr, stage, xb, xphase, &tmC, &tmD, row0, n_blk * C::BN + int(half) * 128, lane, args,
                               release_tmem);
        else
          epilogue_resid_ln<NCH, C::XRING, (RV == 2) ? 2 : (RV == 0) ? 0 : 1>(
              t_addr, stage, xb, xj, request_x, &tmC, &tmD, row0, n_blk * C::BN + int(half) * 128, lane, args, release_tmem);
      } else {
        epilogue_staged<ACT, OUT, LNF>(t_addr, stage, &tmC, row0, n_blk * C::BN + int(half) * 128, lane, args,
                                       release_tmem);
      }
    }
    if constexpr (OUT != OUT_F32_PATCH) {
      if (lane == 0) ptx::bulk_wait<0>();  // smem must outlive the last TMA store's read; writes are done too
    }
  }

  // ===================================================== teardown
  __syncwarp();
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 10) ptx::tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
}

}  // namespace gemm
