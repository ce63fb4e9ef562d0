// Multi-head self-attention core for head dim 64 on sm_100a (tcgen05 + TMEM + TMA), flash style: the
// [L, L] probability matrix the reference materialises (nn.MultiheadAttention slow path with
// need_weights=True, model/transformer.py:200,237) never leaves the SM.
//
// PERSISTENT kernel: 2 CTAs per SM, each walks work items (128-row query tile, head, image) round robin; barriers,
// the TMEM allocation and the K/V/P rings live for the whole launch and all pipelines run ACROSS item boundaries
// (the next item's Q/K/V loads and first S = Q K^T tiles are in flight while the previous item's tail is still
// being exponentiated; the previous item's output is normalised and stored after the next item's first tile).
// 6 warps; the two single-thread control roles sit in the HIGHEST warp ids (the SMSP arbiter favours high warp
// ids, a delayed MMA issue or TMA request stalls all four softmax warps, and the control warps issue very little):
//   warp 5      TMA producer: Q tile per item (double buffered), K_j / V_j tiles (64 keys) through 3-deep rings
//   warp 4      TMEM allocator + tcgen05.mma issuer:  S_g = Q K_j^T (M128 N64 K64) into a double-buffered TMEM
//               accumulator, always two tiles ahead of O += P_g V_j (M128 N64 K64; O double buffered per item)
//   warps 0..3  softmax, thread == query row: tcgen05.ld of the 64 scores of the tile, then ONE streaming pass:
//               exp2 on the SFU against a STALE stabiliser, row max / fp32 row sum / bf16 pack / swizzled smem store
//               of P all in its shadow.  Only when the row max has grown by more than 2^8 since the stabiliser was
//               adopted is O touched (tcgen05.ld / scale / tcgen05.st) and the tile redone; softmax is shift
//               invariant and P <= 256 keeps bf16's relative precision, so the result is unchanged.
// The kernel's bound is the SFU (one MUFU.EX2 per score, 16 lanes/SM, measured tools/micro/mufu_bench.cu);
// POLY of every 8 exponentials can be evaluated on the FMA pipe instead (ptx::ex2_poly3).
// Operands come straight out of the fused QKV GEMM output qkv[B*L, 3*heads*64] through 2-D tensor maps: rows past
// the image's last token are either the next image's tokens or TMA zero fill and are masked.  V tiles are consumed
// as an MN-major B operand exactly as TMA lands them (no transpose).  The ragged last key tile (577 = 9*64 + 1)
// only pays for the 32-key group(s) that hold valid keys.
//
// mbarrier protocol notes (each was a silent-wrong-result bug once):
//   * a parity wait can only tell the current phase from the previous one, so every wait below is on a barrier
//     that provably cannot be two phases ahead of (or behind) the waiter;
//   * p_full has four arrivers (one per softmax warp) that nothing else orders: two alternating instances, so
//     two arrivals of one warp never land in one phase;
//   * the tensor pipe retires one thread's MMAs in issue order and tcgen05.commit covers everything issued
//     before it: "S_g complete" therefore implies "PV_{g-2} complete", which frees P buffer g & 1 without a wait.
#include <limits.h>
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace attn {

constexpr int D = 64;          // head dim
constexpr int BQ = 128;        // query rows per work item
constexpr int BKV = 64;        // keys per tile
constexpr int NST = 3;         // K / V ring depth
constexpr int Q_BYTES = BQ * D * 2;      // 16 KB
constexpr int KV_BYTES = BKV * D * 2;    //  8 KB
constexpr int P_BYTES = BQ * BKV * 2;    // 16 KB: one 128B-swizzle atom column (64 keys) x 128 rows
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;           // S0 [0,64) S1 [64,128) O0 [128,192) O1 [192,256)
constexpr int OFF_Q = 0;                           // 2 buffers
constexpr int OFF_K = OFF_Q + 2 * Q_BYTES;
constexpr int OFF_V = OFF_K + NST * KV_BYTES;
constexpr int OFF_P = OFF_V + NST * KV_BYTES;      // 2 buffers: P_g is written while PV_{g-1} still reads
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;          // 114944 B: two CTAs per SM

struct Bars {
  uint64_t q_full[2], q_empty[2], k_full[NST], v_full[NST], k_empty[NST], v_empty[NST], s_full[2], p_free[2],
      p_full[2];
  uint32_t tmem_slot;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

// work item -> (query tile, head, image); consecutive items share K/V (same head and image) for L2 reuse
struct Item {
  int q0, h, row_base, kv_end, n_kv;
};
__device__ __forceinline__ Item decode_item(int w, int n_qt, int heads, int L, int causal) {
  Item it;
  const int qt = w % n_qt, r = w / n_qt;
  it.h = r % heads;
  it.row_base = (r / heads) * L;   // first token row of this image in qkv / out
  it.q0 = qt * BQ;
  it.kv_end = causal ? min(L, it.q0 + BQ) : L;   // keys [0, kv_end) can be visible to this query tile
  it.n_kv = (it.kv_end + BKV - 1) / BKV;
  return it;
}

template <bool TRACE_ON, int POLY>
__global__ void __launch_bounds__(THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 __nv_bfloat16* __restrict__ out, int L, int heads, int causal, int n_items, int n_qt,
                 long long* trace, int trace_cta, float rescale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-B alignment
  Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BAR);
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = ptx::lane_id();
  const int W = heads * D;
  const int first = blockIdx.x, stride = gridDim.x;
  // optional clock64 trace of one CTA's SECOND work item (steady state; diagnostics):
  // [0..7][tile] softmax warp 0, [16..20][tile] MMA thread
  const bool tracing = TRACE_ON && (trace != nullptr) && (int(blockIdx.x) == trace_cta);
#define TRACE(slot, tile) do { if (TRACE_ON && tracing && (tile) < 16) trace[(slot) * 16 + (tile)] = clock64(); } while (0)

  if (warp == 5 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmKV);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->q_full[i], 1);
      ptx::mbar_init(&bars->q_empty[i], 1);
      ptx::mbar_init(&bars->s_full[i], 1);
      ptx::mbar_init(&bars->p_free[i], 1);
      ptx::mbar_init(&bars->p_full[i], 4);   // one arrive per softmax warp
    }
    for (int i = 0; i < NST; ++i) {
      ptx::mbar_init(&bars->k_full[i], 1);
      ptx::mbar_init(&bars->v_full[i], 1);
      ptx::mbar_init(&bars->k_empty[i], 1);
      ptx::mbar_init(&bars->v_empty[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc<1>(&bars->tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_slot);

  if (warp == 5) {
    // ===================================================== TMA producer
    if (ptx::elect_one()) {
      int st = 0; uint32_t ph = 0;
      uint32_t n = 0;   // item ordinal of this CTA
      for (int w = first; w < n_items; w += stride, ++n) {
        const Item it = decode_item(w, n_qt, heads, L, causal);
        const uint32_t qb = n & 1u;
        ptx::mbar_wait(&bars->q_empty[qb], ((n >> 1) & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&bars->q_full[qb], Q_BYTES);
        ptx::tma_load_2d(smem + OFF_Q + qb * Q_BYTES, &tmQ, &bars->q_full[qb], it.h * D, it.row_base + it.q0);
        for (int j = 0; j < it.n_kv; ++j) {
          ptx::mbar_wait(&bars->k_empty[st], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&bars->k_full[st], KV_BYTES);
          ptx::tma_load_2d(smem + OFF_K + st * KV_BYTES, &tmKV, &bars->k_full[st], W + it.h * D, it.row_base + j * BKV);
          ptx::mbar_wait(&bars->v_empty[st], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&bars->v_full[st], KV_BYTES);
          ptx::tma_load_2d(smem + OFF_V + st * KV_BYTES, &tmKV, &bars->v_full[st], 2 * W + it.h * D,
                           it.row_base + j * BKV);
          if (++st == NST) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 4) {
    // ===================================================== MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(BQ, BKV, 0, 0);  // Q K-major, K K-major
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32(BQ, D, 0, 1);    // P K-major, V MN-major
      // smem matrix descriptors: constant high word (SBO 1024 B, version 1, SWIZZLE_128B), low word = addr >> 4
      // (| LBO field); stepping a tile / a 16-element K slice is an integer add on the low word.
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      auto desc = [](uint32_t lo) { return (uint64_t(DESC_HI) << 32) | lo; };
      const uint32_t q_lo = ((ptx::smem_u32(smem + OFF_Q) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t k_lo = ((ptx::smem_u32(smem + OFF_K) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t p_lo = ((ptx::smem_u32(smem + OFF_P) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t v_lo = ((ptx::smem_u32(smem + OFF_V) & 0x3FFFFu) >> 4) | ((1024u >> 4) << 16);
      // ---- S cursor: runs two tiles ahead of the PV cursor, across item boundaries
      int s_w = first, s_j = 0, s_st = 0; uint32_t s_ph = 0, s_n = 0, s_g = 0;
      int s_nkv = (s_w < n_items) ? decode_item(s_w, n_qt, heads, L, causal).n_kv : 0;
      auto issue_s = [&]() {   // S tile s_g -> TMEM S[s_g & 1]
        if (s_w >= n_items) return;
        const uint32_t qb = s_n & 1u;
        if (s_j == 0) ptx::mbar_wait(&bars->q_full[qb], (s_n >> 1) & 1u);
        ptx::mbar_wait(&bars->k_full[s_st], s_ph);
        ptx::tc_fence_after();
        const uint32_t qd = q_lo + qb * (Q_BYTES >> 4);
        const uint32_t kb = k_lo + uint32_t(s_st) * (KV_BYTES >> 4);
        const uint32_t t_s = tmem_base + (s_g & 1u) * 64u;
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::mma_f16_ss<1>(t_s, desc(qd + k * 2), desc(kb + k * 2), idesc_s, k != 0 ? 1u : 0u);
        ptx::mma_commit(&bars->s_full[s_g & 1u]);
        ptx::mma_commit(&bars->k_empty[s_st]);
        ++s_g;
        if (++s_st == NST) { s_st = 0; s_ph ^= 1u; }
        if (++s_j == s_nkv) {          // last S tile of the item: its Q buffer is free once these MMAs retire
          ptx::mma_commit(&bars->q_empty[qb]);
          s_w += stride; ++s_n; s_j = 0;
          if (s_w < n_items) s_nkv = decode_item(s_w, n_qt, heads, L, causal).n_kv;
        }
      };
      issue_s();
      issue_s();
      // ---- PV cursor
      int st = 0; uint32_t ph = 0, g = 0, n = 0;
      for (int w = first; w < n_items; w += stride, ++n) {
        const Item it = decode_item(w, n_qt, heads, L, causal);
        const uint32_t t_o = tmem_base + 128u + (n & 1u) * 64u;
        for (int j = 0; j < it.n_kv; ++j, ++g) {
          if (n == 1) TRACE(16, j);
          ptx::mbar_wait(&bars->p_full[g & 1u], (g >> 1) & 1u);  // P_g in smem, S[g&1] drained, O rescaled if needed
          if (n == 1) TRACE(17, j);
          ptx::mbar_wait(&bars->v_full[st], ph);
          ptx::tc_fence_after();
          if (n == 1) TRACE(18, j);
          const uint32_t pb = p_lo + (g & 1u) * (P_BYTES >> 4);
          const uint32_t vb = v_lo + uint32_t(st) * (KV_BYTES >> 4);
          const uint32_t acc0 = j > 0 ? 1u : 0u;   // O accumulates over the key tiles of one item
          // A: P tile = one 64-key swizzle atom column (128 rows x 128 B), 16 keys = +32 B;
          // B: V tile as TMA landed it (MN-major), 16 keys = 16 rows x 128 B = +2048 B
          if (j + 1 < it.n_kv || it.kv_end - j * BKV >= BKV) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)
              ptx::mma_f16_ss<1>(t_o, desc(pb + k * 2), desc(vb + k * 128), idesc_o, k != 0 ? 1u : acc0);
          } else {   // ragged last tile: skip 16-key groups that are fully hidden
            const int ksteps = (it.kv_end - j * BKV + 15) >> 4;
            for (int k = 0; k < ksteps; ++k)
              ptx::mma_f16_ss<1>(t_o, desc(pb + k * 2), desc(vb + k * 128), idesc_o, k != 0 ? 1u : acc0);
          }
          ptx::mma_commit(&bars->v_empty[st]);
          ptx::mma_commit(&bars->p_free[g & 1u]);   // PV_g has landed in O and has finished reading P[g & 1]
          if (++st == NST) { st = 0; ph ^= 1u; }
          if (n == 1) TRACE(19, j);
          issue_s();                                // S tile g + 2 (S[g & 1] was drained before p_full completed)
          if (n == 1) TRACE(20, j);
        }
      }
    }
  } else {
    // ===================================================== softmax warps: thread == query row
    const uint32_t quarter = warp & 3u;
    const int row = int(quarter * 32u + lane);
    const uint32_t t_lane = tmem_base + ((quarter * 32u) << 16);
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const float RESCALE_LOG2 = rescale_log2;       // lazy rescale: tolerate p up to 2^8 before touching O
    uint8_t* p_row = smem + OFF_P + row * 128;
    const uint32_t sw = uint32_t(row & 7);
    uint32_t sv[64];                               // raw scores of the current tile (this thread's row)
    uint32_t (&sv_lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[0]);
    uint32_t (&sv_hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[32]);

    // issue the TMEM loads of S tile g (no wait): 32 or 64 columns depending on how many keys the tile holds
    auto load_scores = [&](uint32_t g, int tile_keys) {
      const uint32_t t_s = t_lane + (g & 1u) * 64u;
      ptx::tmem_ld_32x32b_x32(t_s, sv_lo);
      if (tile_keys > 32) ptx::tmem_ld_32x32b_x32(t_s + 32, sv_hi);
    };
    // hidden keys -> -inf -> p = 0 (last key tile / causal diagonal only)
    auto mask_scores = [&](int kv0, int q0, int qi) {
      if ((kv0 + BKV > L) || (causal && kv0 + BKV > q0 + 1)) {
        int limit = L - kv0;                            // this row sees keys [0, limit) of the tile
        if (causal) limit = min(limit, qi - kv0 + 1);   // CLIP text mask (model/model.py:172): keys > qi hidden
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= limit) sv[i] = 0xff800000u;
      }
    };
    // One streaming pass over the tile: p = exp2((s - m) * c) -> bf16 -> swizzled smem (16 B per 8 keys, stored as
    // soon as packed), fp32 row sum and the tile's row max, all in one instruction stream so that the FMNMX /
    // FADD / F2FP / STS work hides under the SFU (MUFU.EX2) latency instead of forming serial phases.
    auto exp_pass = [&](uint8_t* pr, float mc, int n_groups, float& rowsum, float& rowmax) {
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) {
        if (g8 < n_groups) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x = fmaf(__uint_as_float(sv[8 * g8 + i]), c, -mc);
            e[i] = (i >= 8 - POLY) ? ptx::ex2_poly3(x) : ptx::ex2_approx(x);
          }
          mx4[g8 & 3] = fmaxf(mx4[g8 & 3],
                              fmaxf(fmaxf(__uint_as_float(sv[8 * g8 + 0]), __uint_as_float(sv[8 * g8 + 1])),
                                    fmaxf(__uint_as_float(sv[8 * g8 + 2]), __uint_as_float(sv[8 * g8 + 3]))));
          mx4[(g8 + 2) & 3] = fmaxf(mx4[(g8 + 2) & 3],
                                    fmaxf(fmaxf(__uint_as_float(sv[8 * g8 + 4]), __uint_as_float(sv[8 * g8 + 5])),
                                          fmaxf(__uint_as_float(sv[8 * g8 + 6]), __uint_as_float(sv[8 * g8 + 7]))));
          rs4[g8 & 3] += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
          ptx::st_shared_v4(pr + ((uint32_t(g8) ^ sw) << 4), ptx::pack_bf16x2(e[0], e[1]), ptx::pack_bf16x2(e[2], e[3]),
                            ptx::pack_bf16x2(e[4], e[5]), ptx::pack_bf16x2(e[6], e[7]));
        }
      }
      rowsum = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      rowmax = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    };
    // normalise O of a finished item by its row sum and store it (deferred: runs after the next item's first tile)
    auto store_output = [&](uint32_t t_o, uint32_t g_last, float l, __nv_bfloat16* orow, bool valid) {
      // PV_{g_last} complete.  The next completion of this barrier needs P_{g_last+2}, which this warp has not
      // produced yet, so the barrier is at most one phase ahead: the parity wait is unambiguous.
      ptx::mbar_wait(&bars->p_free[g_last & 1u], (g_last >> 1) & 1u);
      ptx::tc_fence_after();
      const float inv = 1.0f / l;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_o + hh * 32, v);
        ptx::tmem_ld_wait();
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(orow + hh * 32);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 wv;
            wv.x = ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 0]) * inv, __uint_as_float(v[q4 * 8 + 1]) * inv);
            wv.y = ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 2]) * inv, __uint_as_float(v[q4 * 8 + 3]) * inv);
            wv.z = ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 4]) * inv, __uint_as_float(v[q4 * 8 + 5]) * inv);
            wv.w = ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 6]) * inv, __uint_as_float(v[q4 * 8 + 7]) * inv);
            dst[q4] = wv;
          }
        }
      }
    };

    const bool tr = TRACE_ON && tracing && warp == 0 && lane == 0;
#define TRS(slot) do { if (TRACE_ON && tr && n == 1 && j < 16) trace[(slot) * 16 + j] = clock64(); } while (0)
    // pending output of the previous item
    bool pend = false, pend_valid = false;
    uint32_t pend_to = 0, pend_g = 0;
    float pend_l = 1.f;
    __nv_bfloat16* pend_row = nullptr;

    uint32_t g = 0, n = 0;
    if (first < n_items) {
      const Item it0 = decode_item(first, n_qt, heads, L, causal);
      ptx::mbar_wait(&bars->s_full[0], 0);
      ptx::tc_fence_after();
      load_scores(0, min(BKV, it0.kv_end));
      ptx::tmem_ld_wait();
    }
    for (int w = first; w < n_items; w += stride, ++n) {
      const Item it = decode_item(w, n_qt, heads, L, causal);
      const int qi = it.q0 + row;
      const uint32_t t_o = t_lane + 128u + (n & 1u) * 64u;
      float m_used = 0.f, l = 0.f;                   // stabiliser in use (<= true running max + 2^8), row sum
      // tile geometry of the NEXT item's first tile (for the prefetch across the item boundary)
      const int w_next = w + stride;
      const int next_first_keys = (w_next < n_items) ? min(BKV, decode_item(w_next, n_qt, heads, L, causal).kv_end) : 0;

      for (int j = 0; j < it.n_kv; ++j, ++g) {
        TRS(0);
        const int kv0 = j * BKV;
        const int n_groups = min(BKV, it.kv_end - kv0) > 32 ? 8 : 4;   // 8-key groups holding visible keys
        uint8_t* pr = p_row + (g & 1u) * P_BYTES;
        // P buffer g & 1 was last read by PV_{g-2}: free, because S_g (seen complete) was issued after it.
        mask_scores(kv0, it.q0, qi);
        if (j == 0) {   // first tile: adopt its true row max (a fully hidden row - rows >= L, never stored - uses 0)
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 64; i += 2)
            if (i < 8 * n_groups)
              mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
          const float mt0 = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          m_used = (mt0 == -INFINITY) ? 0.f : mt0;
        }
        TRS(1);
        float rs, mt;
        exp_pass(pr, m_used * c, n_groups, rs, mt);
        TRS(2);
        // ---- stale stabiliser check: only when some row's max grew by more than 2^RESCALE is O touched
        const bool grow = (mt - m_used) * c > RESCALE_LOG2;
        if (__any_sync(0xffffffffu, grow)) {          // warp-uniform: tcgen05.ld/st are warp-collective
          const float m_next = grow ? mt : m_used;
          const float alpha = ptx::ex2_approx((m_used - m_next) * c);   // 1 for rows that keep their max
          if (j > 0) {
            // PV_{g-1} must have landed in O.  p_free[b] completes with PV_b, PV_{b+2}, ...; S_g complete implies
            // PV_{g-3} complete (issue order) and PV_{g+1} needs this warp's P_{g+1}: the barrier is in the phase
            // of PV_{g-1} or one past it.
            ptx::mbar_wait(&bars->p_free[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
            ptx::tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[32];
              ptx::tmem_ld_32x32b_x32(t_o + hh * 32, v);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              ptx::tmem_st_32x32b_x32(t_o + hh * 32, v);
            }
            ptx::tmem_st_wait();
          }
          l *= alpha;
          m_used = m_next;
          // redo the tile against the new stabiliser (S_g is still in TMEM: it is released by the p_full arrive)
          load_scores(g, min(BKV, it.kv_end - kv0));
          ptx::tmem_ld_wait();
          mask_scores(kv0, it.q0, qi);
          exp_pass(pr, m_used * c, n_groups, rs, mt);
        }
        l += rs;
        TRS(3);
        // ---- prefetch the next S tile (possibly the next item's first) while the P hand-over is in flight
        const bool last_tile = (j + 1 == it.n_kv);
        const int next_keys = last_tile ? next_first_keys : min(BKV, it.kv_end - (kv0 + BKV));
        if (next_keys > 0) {
          ptx::mbar_wait(&bars->s_full[(g + 1) & 1u], ((g + 1) >> 1) & 1u);
          ptx::tc_fence_after();
          load_scores(g + 1, next_keys);
        }
        TRS(4);
        ptx::fence_proxy_async_smem();  // generic-proxy P stores -> visible to the tensor core (async proxy)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->p_full[g & 1u]);   // P_g in smem, S[g & 1] drained, O rescaled
        TRS(5);
        ptx::tmem_ld_wait();
        // ---- the previous item's output, one tile late: its last PV has long landed by now
        if (j == 0 && pend) {
          store_output(pend_to, pend_g, pend_l, pend_row, pend_valid);
          pend = false;
        }
        TRS(6);
      }
      pend = true; pend_to = t_o; pend_g = g - 1; pend_l = l; pend_valid = (qi < L);
      pend_row = out + (size_t)(it.row_base + qi) * W + it.h * D;
    }
    if (pend) store_output(pend_to, pend_g, pend_l, pend_row, pend_valid);
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

}  // namespace attn

constexpr int ATTN_POLY_DEFAULT = 0;
namespace {
long long* g_trace = nullptr;   // diagnostics only (aaclip_attention_trace)
int g_trace_cta = -1;
}

int k::launch_attention(const void* qkv, void* out, int B, int L, int heads, int causal, cudaStream_t stream) {
  if (B <= 0) return host::OK;
  if (L <= 0 || heads <= 0) return host::fail(host::ERR_INVALID, "attention: L=%d heads=%d", L, heads);
  const int W = heads * attn::D;
  CUtensorMap tmQ, tmKV;
  int rc = host::make_tmap_2d(&tmQ, qkv, (uint64_t)B * L, 3 * W, 3 * W, attn::BQ);
  if (rc) return rc;
  rc = host::make_tmap_2d(&tmKV, qkv, (uint64_t)B * L, 3 * W, 3 * W, attn::BKV);
  if (rc) return rc;
  // AACLIP_ATTN_POLY=<0..4> (diagnostics): exponentials per 8 evaluated on the FMA pipe instead of the SFU
  static int poly = getenv("AACLIP_ATTN_POLY") ? atoi(getenv("AACLIP_ATTN_POLY")) : ATTN_POLY_DEFAULT;
  static float rescale = getenv("AACLIP_ATTN_RESCALE") ? (float)atof(getenv("AACLIP_ATTN_RESCALE")) : 8.0f;
  typedef void (*Kern)(const CUtensorMap, const CUtensorMap, __nv_bfloat16*, int, int, int, int, int, long long*, int,
                       float);
  Kern kern = nullptr;
  if (g_trace) kern = attn::attention_kernel<true, 0>;
  else switch (poly) {
    case 0: kern = attn::attention_kernel<false, 0>; break;
    case 1: kern = attn::attention_kernel<false, 1>; break;
    case 2: kern = attn::attention_kernel<false, 2>; break;
    case 3: kern = attn::attention_kernel<false, 3>; break;
    case 4: kern = attn::attention_kernel<false, 4>; break;
    default: return host::fail(host::ERR_INVALID, "AACLIP_ATTN_POLY=%d out of range [0,4]", poly);
  }
  static Kern configured[8] = {nullptr};
  bool seen = false;
  for (Kern kk : configured) seen = seen || (kk == kern);
  if (!seen) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES));
    for (Kern& kk : configured) if (!kk) { kk = kern; break; }
  }
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  const int n_qt = (L + attn::BQ - 1) / attn::BQ;
  const long long items = (long long)n_qt * heads * B;
  if (items > INT_MAX) return host::fail(host::ERR_INVALID, "attention: %lld work items", items);
  const int sms = host::sm_count(dev);
  const int grid = (int)std::min<long long>(items, 2LL * (sms > 0 ? sms : 148));
  kern<<<grid, attn::THREADS, attn::SMEM_BYTES, stream>>>(tmQ, tmKV, static_cast<__nv_bfloat16*>(out), L, heads, causal,
                                                          (int)items, n_qt, g_trace, g_trace_cta, rescale);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

extern "C" int aaclip_attention(const void* qkv, void* out, int B, int L, int heads, int causal, void* stream) {
  return k::launch_attention(qkv, out, B, L, heads, causal, static_cast<cudaStream_t>(stream));
}

// Diagnostics: like aaclip_attention, but CTA number `cta` also records clock64() stamps of its softmax warp 0
// (slots 0..7) and of its MMA-issuing thread (slots 16..20) per key tile of its second work item into
// trace[slot * 16 + tile] (device memory, >= 24 * 16 int64).  Used to study the pipeline; not part of the hot path.
extern "C" int aaclip_attention_trace(const void* qkv, void* out, int B, int L, int heads, int causal, long long* trace,
                                      int cta, void* stream) {
  g_trace = trace; g_trace_cta = cta;
  int rc = k::launch_attention(qkv, out, B, L, heads, causal, static_cast<cudaStream_t>(stream));
  g_trace = nullptr; g_trace_cta = -1;
  return rc;
}
