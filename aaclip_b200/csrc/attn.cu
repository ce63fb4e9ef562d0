// Multi-head self-attention core for head dim 64 on sm_100a (tcgen05 + TMEM + TMA), flash style: the
// [L, L] probability matrix the reference materialises (nn.MultiheadAttention slow path with
// need_weights=True, model/transformer.py:200,237) never leaves the SM.
//
// The kernel's bound is the SFU: one MUFU.EX2 per score at 16 lanes/SM/clk (tools/micro/mufu_bench.cu: the
// packed f16x2 / bf16x2 forms run at the same per-element rate), the tensor pipe is ~20 % busy.  Everything is
// arranged to keep the four SFUs of an SM fed: FOUR small CTAs per SM (one softmax warp of each on every SMSP),
// 32-key tiles so that a CTA needs only 53.5 KB smem / 128 TMEM columns / 80 registers, and a softmax loop whose
// common case is a branch-free ~170-instruction stream per tile (instruction issue is the second bound).
//
// PERSISTENT: each CTA walks work items (128-row query tile, head, image) round robin; barriers, the TMEM
// allocation and the K/V/P rings live for the whole launch and the pipelines run ACROSS item boundaries.
// 6 warps; the two single-thread control roles sit in the HIGHEST warp ids (the SMSP arbiter favours high warp
// ids, a late MMA issue or TMA request stalls all four softmax warps, and the control warps issue very little):
//   warp 5      TMA producer: Q tile per item, K_j / V_j tiles (32 keys each) through 3- / 2-deep rings
//   warp 4      TMEM allocator + tcgen05.mma issuer:  S_g = Q K_j^T (M128 N32 K64) into a double-buffered TMEM
//               accumulator, always two tiles ahead of O += P_g V_j (M128 N64 K32)
//   warps 0..3  softmax, thread == query row: tcgen05.ld of the tile's 32 scores, then ONE streaming pass:
//               exp2 on the SFU against a STALE stabiliser, with row max / fp32 row sum / bf16 pack / swizzled smem
//               store of P in its shadow; the next S tile is loaded from TMEM before the P hand-over.  Only when a
//               row max has grown by more than 2^8 since the stabiliser was adopted is O touched (tcgen05.ld /
//               scale / tcgen05.st) and the tile redone: softmax is shift invariant and P <= 256 keeps bf16's
//               relative precision, so the result is unchanged.
//               Item end: O / l -> bf16 -> the warp's own (idle) rows of the P buffer -> one TMA tile store through
//               a rank-3 [B][L][W] tensor map (rows past the image's last token are clipped by the TMA unit).
// P is ONE 128-row x 128-B swizzled buffer whose two 64-B halves hold alternate tiles (double buffering at no
// extra smem); POLY of every 8 exponentials can be moved to the FMA pipe (ptx::ex2_poly3; measured slower here).
// Operands come straight out of the fused QKV GEMM output qkv[B*L, 3*heads*64] through 2-D tensor maps: rows past
// the image's last token are either the next image's tokens or TMA zero fill and are masked.  V tiles are consumed
// as an MN-major B operand exactly as TMA lands them (no transpose).  The ragged last key tile (577 = 18*32 + 1)
// only pays for one 16-key MMA step.
//
// mbarrier protocol notes (each was a silent-wrong-result bug once):
//   * a parity wait can only tell the current phase from the previous one, so every wait below is on a barrier
//     that provably cannot be two phases ahead of (or behind) the waiter;
//   * p_full has four arrivers (one per softmax warp) that nothing else orders: two alternating instances, so
//     two arrivals of one warp never land in one phase;
//   * the tensor pipe retires one thread's MMAs in issue order and tcgen05.commit covers everything issued
//     before it: "S_g complete" therefore implies "PV_{g-2} complete", which frees P half g & 1 without a wait.
#include <limits.h>
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include <type_traits>
#include "common.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace attn {

constexpr int D = 64;          // head dim
constexpr int BQ = 128;        // query rows per work item
constexpr int BKV = 32;        // keys per tile
constexpr int NSTK = 2;        // K ring depth (K_g is consumed as soon as it lands, two tiles ahead of V_g)
constexpr int NSTV = 3;        // V ring depth (V_g waits for P_g: its slot is held two tiles longer)
constexpr int Q_BYTES = BQ * D * 2;      // 16 KB
constexpr int KV_BYTES = BKV * D * 2;    //  4 KB
constexpr int P_BYTES = BQ * 128;        // 16 KB: 128 rows x 128 B = two 32-key tiles side by side (double buffer)
constexpr int THREADS = 192;           // 6 warps; the balanced variant (BAL) runs 8: 4 softmax + 4 control slots
constexpr int THREADS_BAL = 256;
constexpr int SOFTMAX_REGS = 96, CTL_REGS = 32;   // BAL: 128 * 96 + 128 * 32 = 256 * 64 (the launch allocation)
constexpr int TMEM_COLS = 128;           // S0 [0,32) S1 [32,64) O [64,128)

template <int NQ>
struct Lay {
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NQ * Q_BYTES;
  static constexpr int OFF_V = OFF_K + NSTK * KV_BYTES;
  static constexpr int OFF_P = OFF_V + NSTV * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;   // NQ=1: 53504 B (4 CTAs/SM), NQ=2: 69888 B (3 CTAs/SM)
};

struct Bars {
  uint64_t q_full[2], q_empty[2], k_full[NSTK], v_full[NSTV], k_empty[NSTK], v_empty[NSTV], s_full[2], p_free[2],
      p_full[2];
  uint32_t tmem_slot;
  uint32_t sm_slot;   // BAL: ordinal (mod 4) of this CTA among the CTAs that started on its SM
};
static_assert(sizeof(Bars) <= 256, "barrier block");

// work item -> (query tile, head, image); consecutive items share K/V (same head and image) for L2 reuse
struct Item {
  int q0, h, b, row_base, kv_end, n_kv;
};
__device__ __forceinline__ Item decode_item(int w, int n_qt, int heads, int L, int causal) {
  Item it;
  const int qt = w % n_qt, r = w / n_qt;
  it.h = r % heads;
  it.b = r / heads;
  it.row_base = it.b * L;          // first token row of this image in qkv
  it.q0 = qt * BQ;
  it.kv_end = causal ? min(L, it.q0 + BQ) : L;   // keys [0, kv_end) can be visible to this query tile
  it.n_kv = (it.kv_end + BKV - 1) / BKV;
  return it;
}

// BAL: per-SM arrival counters (monotonic; 4 CTAs start per SM per full-size launch, so `& 3` keeps rotating)
__device__ unsigned int g_sm_arrivals[1024];

// DEEP: S look-ahead of three tiles out of the same two TMEM buffers.  A softmax warp keeps S_g in REGISTERS while it
//   works on tile g, so buffer g & 1 is free as soon as every warp has loaded it - a whole tile earlier than "P_g is
//   ready".  The warps therefore complete the TMEM load of S_{g+1} BEFORE the p_full[g] arrive, which then also means
//   "S_{g+1} is in registers", and the MMA thread issues S_{g+3} (not S_{g+2}) behind P_g V_g.  Price: S_{g+2} is now
//   issued ahead of P_g V_g, so "S_{g+2} complete" no longer proves that P half g & 1 has been read - an explicit
//   p_free wait guards the rewrite - and the stale-stabiliser redo works from the register copy (no TMEM reload).
template <bool TRACE_ON, int POLY, int MINB, int NQ, int BAL, int DEEP>
__global__ void __launch_bounds__(BAL ? THREADS_BAL : THREADS, MINB)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmO, int L, int heads, int causal, int n_items, int n_qt,
                 long long* trace, int trace_cta, const float RESCALE_MAX) {
  using LY = Lay<NQ>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-B alignment
  Bars* bars = reinterpret_cast<Bars*>(smem + LY::OFF_BAR);
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = ptx::lane_id();
  const int W = heads * D;
  const int first = blockIdx.x, stride = gridDim.x;
  // optional clock64 trace of one CTA's SECOND work item (steady state; diagnostics):
  // [0..7][tile] softmax warp 0, [16..20][tile] MMA thread
  const bool tracing = TRACE_ON && (trace != nullptr) && (int(blockIdx.x) == trace_cta);
#define TRACE(slot, tile) do { if (TRACE_ON && tracing && (tile) < 32) trace[(slot) * 32 + (tile)] = clock64(); } while (0)

  if (warp == 5 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmKV);
    ptx::prefetch_tmap(&tmO);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->q_full[i], 1);
      ptx::mbar_init(&bars->q_empty[i], 1);
      ptx::mbar_init(&bars->s_full[i], 1);
      ptx::mbar_init(&bars->p_free[i], 1);
      ptx::mbar_init(&bars->p_full[i], 4);   // one arrive per softmax warp
    }
    for (int i = 0; i < NSTK; ++i) { ptx::mbar_init(&bars->k_full[i], 1); ptx::mbar_init(&bars->k_empty[i], 1); }
    for (int i = 0; i < NSTV; ++i) { ptx::mbar_init(&bars->v_full[i], 1); ptx::mbar_init(&bars->v_empty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc<1>(&bars->tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish<1>();
  }
  if (BAL && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    bars->sm_slot = atomicAdd(&g_sm_arrivals[smid & 1023u], 1u) & 3u;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_slot);
  // Control roles.  A CTA-relative warp w always runs on scheduler w % 4 (TMEM lane quarters are tied to it), so with
  // 6 warps the two single-thread roles of ALL FOUR resident CTAs land on schedulers 0 and 1 and make those two
  // issue bound while 2 and 3 idle.  BAL: 8 warps (4 softmax + 4 control slots, two of which only park at the final
  // barrier); the CTA's arrival ordinal on its SM picks which slots carry the MMA issuer and the TMA producer, so
  // every scheduler of the SM ends up with one of each.  Registers: 64 per thread at launch (256 x 4 CTAs), then
  // the softmax warpgroup grows to 96 out of what the control warpgroup gives back (32 each).
  uint32_t mma_warp = 4, tma_warp = 5;
  if (BAL) {
    const uint32_t slot = *reinterpret_cast<volatile uint32_t*>(&bars->sm_slot);
    const uint32_t pair = 4u + ((slot & 1u) << 1), flip = (slot >> 1) & 1u;
    mma_warp = pair + flip;
    tma_warp = pair + (flip ^ 1u);
  }
  ptx::grid_dep_sync();   // everything above overlapped the previous kernel's tail

  if (warp == tma_warp) {
    // ===================================================== TMA producer
    if (BAL) ptx::reg_dec<CTL_REGS>();
    if (ptx::elect_one()) {
      int kst = 0, vst = 0; uint32_t kph = 0, vph = 0;
      uint32_t n = 0;   // item ordinal of this CTA
      for (int w = first; w < n_items; w += stride, ++n) {
        const Item it = decode_item(w, n_qt, heads, L, causal);
        const uint32_t qb = n % NQ, qph = (n / NQ) & 1u;
        ptx::mbar_wait_spin(&bars->q_empty[qb], qph ^ 1u);
        ptx::mbar_arrive_expect_tx(&bars->q_full[qb], Q_BYTES);
        ptx::tma_load_2d(smem + LY::OFF_Q + qb * Q_BYTES, &tmQ, &bars->q_full[qb], it.h * D, it.row_base + it.q0);
        for (int j = 0; j < it.n_kv; ++j) {
          ptx::mbar_wait_spin(&bars->k_empty[kst], kph ^ 1u);
          ptx::mbar_arrive_expect_tx(&bars->k_full[kst], KV_BYTES);
          ptx::tma_load_2d(smem + LY::OFF_K + kst * KV_BYTES, &tmKV, &bars->k_full[kst], W + it.h * D,
                           it.row_base + j * BKV);
          if (++kst == NSTK) { kst = 0; kph ^= 1u; }
          ptx::mbar_wait_spin(&bars->v_empty[vst], vph ^ 1u);
          ptx::mbar_arrive_expect_tx(&bars->v_full[vst], KV_BYTES);
          ptx::tma_load_2d(smem + LY::OFF_V + vst * KV_BYTES, &tmKV, &bars->v_full[vst], 2 * W + it.h * D,
                           it.row_base + j * BKV);
          if (++vst == NSTV) { vst = 0; vph ^= 1u; }
        }
      }
    }
  } else if (warp == mma_warp) {
    // ===================================================== MMA issuer
    if (BAL) ptx::reg_dec<CTL_REGS>();
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(BQ, BKV, 0, 0);  // Q K-major, K K-major
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32(BQ, D, 0, 1);    // P K-major, V MN-major
      // smem matrix descriptors: constant high word (SBO 1024 B, version 1, SWIZZLE_128B), low word = addr >> 4
      // (| LBO field); stepping a tile / a 16-element K slice is an integer add on the low word.
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      auto desc = [](uint32_t lo) { return (uint64_t(DESC_HI) << 32) | lo; };
      const uint32_t q_lo = ((ptx::smem_u32(smem + LY::OFF_Q) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t k_lo = ((ptx::smem_u32(smem + LY::OFF_K) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t p_lo = ((ptx::smem_u32(smem + LY::OFF_P) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t v_lo = ((ptx::smem_u32(smem + LY::OFF_V) & 0x3FFFFu) >> 4) | ((1024u >> 4) << 16);
      // ---- S cursor: runs two tiles ahead of the PV cursor, across item boundaries
      int s_w = first, s_j = 0, s_st = 0; uint32_t s_ph = 0, s_n = 0, s_g = 0;
      int s_nkv = (s_w < n_items) ? decode_item(s_w, n_qt, heads, L, causal).n_kv : 0;
      auto issue_s = [&]() {   // S tile s_g -> TMEM S[s_g & 1]
        if (s_w >= n_items) return;
        const uint32_t qb = s_n % NQ;
        if (s_j == 0) ptx::mbar_wait_spin(&bars->q_full[qb], (s_n / NQ) & 1u);
        ptx::mbar_wait_spin(&bars->k_full[s_st], s_ph);
        ptx::tc_fence_after();
        const uint32_t qd = q_lo + qb * (Q_BYTES >> 4);
        const uint32_t kb = k_lo + uint32_t(s_st) * (KV_BYTES >> 4);
        const uint32_t t_s = tmem_base + (s_g & 1u) * uint32_t(BKV);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::mma_f16_ss<1>(t_s, desc(qd + k * 2), desc(kb + k * 2), idesc_s, k != 0 ? 1u : 0u);
        ptx::mma_commit(&bars->s_full[s_g & 1u]);
        ptx::mma_commit(&bars->k_empty[s_st]);
        ++s_g;
        if (++s_st == NSTK) { s_st = 0; s_ph ^= 1u; }
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
      const uint32_t t_o = tmem_base + 2u * BKV;
      for (int w = first; w < n_items; w += stride, ++n) {
        const Item it = decode_item(w, n_qt, heads, L, causal);
        for (int j = 0; j < it.n_kv; ++j, ++g) {
          if (n == 1) TRACE(16, j);
          ptx::mbar_wait_spin(&bars->p_full[g & 1u], (g >> 1) & 1u);  // P_g in smem, S[g&1] drained, O rescaled / drained
          if (n == 1) TRACE(17, j);
          ptx::mbar_wait_spin(&bars->v_full[st], ph);
          ptx::tc_fence_after();
          if (n == 1) TRACE(18, j);
          // A: P tile g = key columns [(g&1)*32, +32) of the 128-B swizzled P rows: +64 B per tile, +32 B per 16 keys
          // B: V tile as TMA landed it (MN-major), 16 keys = 16 rows x 128 B = +2048 B
          const uint32_t pb = p_lo + (g & 1u) * 4u;
          const uint32_t vb = v_lo + uint32_t(st) * (KV_BYTES >> 4);
          const uint32_t acc0 = j > 0 ? 1u : 0u;   // O accumulates over the key tiles of one item
          if (it.kv_end - j * BKV > 16) {
            ptx::mma_f16_ss<1>(t_o, desc(pb), desc(vb), idesc_o, acc0);
            ptx::mma_f16_ss<1>(t_o, desc(pb + 2), desc(vb + 128), idesc_o, 1u);
          } else {   // ragged last tile: the second 16-key step is fully hidden
            ptx::mma_f16_ss<1>(t_o, desc(pb), desc(vb), idesc_o, acc0);
          }
          ptx::mma_commit(&bars->v_empty[st]);
          ptx::mma_commit(&bars->p_free[g & 1u]);   // PV_g has landed in O and has finished reading P half g & 1
          if (++st == NSTV) { st = 0; ph ^= 1u; }
          if (n == 1) TRACE(19, j);
          // !DEEP: S tile g + 2 (S[g & 1] was drained before p_full[g] completed)
          //  DEEP: p_full[g] also says S_{g+1} sits in registers -> S tile g + 3 into buffer (g + 1) & 1; the very
          //        first hand-over frees both buffers (S_0 consumed, S_1 loaded): S_2 and S_3
          issue_s();
          if (DEEP && g == 0) issue_s();
          if (n == 1) TRACE(20, j);
        }
      }
    }
  } else if (warp < 4) {
    // ===================================================== softmax warps: thread == query row
    if (BAL) ptx::reg_inc<SOFTMAX_REGS>();
    const uint32_t quarter = warp & 3u;
    const int row = int(quarter * 32u + lane);
    const uint32_t t_lane = tmem_base + ((quarter * 32u) << 16);
    const uint32_t t_o = t_lane + 2u * BKV;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    // RESCALE_MAX (a kernel parameter: read from the constant bank, no register): lazy rescale - p and tile row sums may
    // reach 2^8 before O is touched
    uint8_t* p_row = smem + LY::OFF_P + row * 128;
    uint8_t* p_warp = smem + LY::OFF_P + quarter * 32u * 128u;   // this warp's 32 rows x 128 B (output staging)
    const uint32_t sw = uint32_t(row & 7);
    // shared-space address of the 16-B chunk that holds keys [8c, 8c+8) of this row's P line (128B swizzle)
    // Chunks 0..3 of the row hold the tile's keys when g is even, 4..7 when it is odd: (ch + 4) ^ sw == (ch ^ sw) ^ 4, so
    // the odd tile's addresses are the even tile's with bit 6 flipped (the row is 128-byte aligned) - four registers
    // and one XOR per store instead of eight addresses the compiler rebuilt every tile (20 instructions of 147).
    uint32_t p_addr[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) p_addr[ch] = ptx::smem_u32(p_row) + ((uint32_t(ch) ^ sw) << 4);
    uint32_t sv[BKV];                              // raw scores of the current tile (this thread's row)

    auto load_scores = [&](uint32_t g) { ptx::tmem_ld_32x32b_x32(t_lane + (g & 1u) * uint32_t(BKV), sv); };
    // hidden keys -> -inf -> p = 0 (last key tile / causal diagonal only)
    auto mask_scores = [&](int kv0, int q0, int qi) {
      int limit = L - kv0;                            // this row sees keys [0, limit) of the tile
      if (causal) limit = min(limit, qi - kv0 + 1);   // CLIP text mask (model/model.py:172): keys > qi hidden
#pragma unroll
      for (int i = 0; i < BKV; ++i)
        if (i >= limit) sv[i] = 0xff800000u;
    };
    // One streaming pass over NG 8-key groups of the tile: p = exp2((s - m) * c) -> bf16 -> swizzled smem (16 B per
    // group, stored as soon as packed) and the fp32 row sum, in one branch-free instruction stream so that the
    // FFMA2 / FADD2 / F2FP / STS work hides under the SFU (MUFU.EX2) latency.  No row max here: the row sum bounds
    // every p of the tile, which is all the stale-stabiliser check needs.
    // `half` selects the tile's 64-B half of the 128-B P row (chunks half*4 .. half*4+3).
    auto exp_pass = [&](auto ng_tag, uint32_t half, float mc, float& rowsum) {
      constexpr int NG = decltype(ng_tag)::value;
      float r0 = 0.f, r1 = 0.f;
      const float nmc = -mc;
#pragma unroll
      for (int g8 = 0; g8 < NG; ++g8) {
        float x[8], e[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2)
          ptx::ffma2(x[i], x[i + 1], __uint_as_float(sv[8 * g8 + i]), __uint_as_float(sv[8 * g8 + i + 1]), c, nmc);
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = (i >= 8 - POLY) ? ptx::ex2_poly3(x[i]) : ptx::ex2_approx(x[i]);
        float a0, a1, b0, b1;
        ptx::fadd2(a0, a1, e[0], e[1], e[2], e[3]);
        ptx::fadd2(b0, b1, e[4], e[5], e[6], e[7]);
        ptx::fadd2(a0, a1, a0, a1, b0, b1);
        ptx::fadd2(r0, r1, r0, r1, a0, a1);
        const uint32_t a = p_addr[g8] ^ (half << 6);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(ptx::pack_bf16x2(e[0], e[1])),
                     "r"(ptx::pack_bf16x2(e[2], e[3])), "r"(ptx::pack_bf16x2(e[4], e[5])),
                     "r"(ptx::pack_bf16x2(e[6], e[7])) : "memory");
      }
      rowsum = r0 + r1;
    };
    // row max over the first NG 8-key groups of the tile in registers
    auto tile_max = [&](bool four) {
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < BKV; i += 2)
        if (i < 16 || four)
          mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
      return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    };
    using NG4 = std::integral_constant<int, 4>;
    using NG2 = std::integral_constant<int, 2>;

    const bool tr = TRACE_ON && tracing && warp == 0 && lane == 0;
#define TRS(slot) do { if (TRACE_ON && tr && n == 1 && j < 32) trace[(slot) * 32 + j] = clock64(); } while (0)
    bool store_pending = false;   // a TMA store is still reading this warp's P rows (lane 0 owns the bulk group)

    uint32_t g = 0, n = 0;
    if (first < n_items) {
      ptx::mbar_wait(&bars->s_full[0], 0);
      ptx::tc_fence_after();
      load_scores(0);
      ptx::tmem_ld_wait();
    }
    for (int w = first; w < n_items; w += stride, ++n) {
      const Item it = decode_item(w, n_qt, heads, L, causal);
      const int qi = it.q0 + row;
      float m_used = 0.f, l = 0.f;                   // stabiliser in use (<= true running max + 2^8), row sum
      const bool more_items = (w + stride < n_items);
      // tiles [1, j_fast) are complete and unmasked for every row of the item and are not its last tile (so a
      // next S tile always exists): they take the lean path
      const int j_fast = min(it.n_kv - 1, (causal ? min(L, it.q0 + 1) : L) / BKV);

      // One key tile.  FAST: 32 visible keys for every row, not the item's first tile.  Everything else (first tile:
      // adopt the true row max and wait for the previous output store; ragged / causal tiles: mask, fewer groups)
      // goes through the general form.
      auto tile = [&](auto fast_tag, int j) {
        constexpr bool FAST = decltype(fast_tag)::value;
        TRS(0);
        const int kv0 = j * BKV;
        const uint32_t half = g & 1u;
        // P half g & 1 was last read by PV_{g-2}: free, because S_g (seen complete) was issued after it.
        bool four = true;   // 8-key groups to exponentiate: whole 16-key MMA k-steps that hold visible keys
        if constexpr (!FAST) {
          four = (it.kv_end - kv0) > 16;
          if ((kv0 + BKV > L) || (causal && kv0 + BKV > it.q0 + 1)) mask_scores(kv0, it.q0, qi);
          if (j == 0) {   // adopt the true row max (a fully hidden row - rows >= L, never stored - uses 0)
            const float mt0 = tile_max(four);
            m_used = (mt0 == -INFINITY) ? 0.f : mt0;
            if (store_pending) {   // the previous item's output store must have finished reading these P rows
              if (lane == 0) ptx::bulk_wait_read<0>();
              __syncwarp();
              store_pending = false;
            }
          }
        }
        if (DEEP && g >= 2) {
          // PV_{g-2} (issued AFTER S_g in this ordering) must have finished reading P half g & 1.  p_free[g & 1]'s next
          // completion (PV_g) needs this warp's P_g, so the barrier is in the phase of PV_{g-2} or one before it.
          ptx::mbar_wait(&bars->p_free[half], ((g - 2) >> 1) & 1u);
        }
        TRS(1);
        float rs;
        if (FAST || four) exp_pass(NG4{}, half, m_used * c, rs);
        else exp_pass(NG2{}, half, m_used * c, rs);
        TRS(2);
        // ---- stale stabiliser check: p <= row sum, so a row sum within 2^RESCALE proves every p of the tile is; only
        //      when some row exceeds it (its max has grown a lot since the stabiliser was adopted) is O touched
        if (__any_sync(0xffffffffu, !(rs <= RESCALE_MAX))) {   // warp-uniform: tcgen05.ld/st are warp-collective
          // S_g is still in TMEM (released by the p_full arrive): reload it, find the true tile max of the rows that
          // tripped the check and restabilise them; every other row keeps its stabiliser (alpha = 1)
          const bool trip = !(rs <= RESCALE_MAX);
          const bool masked = !FAST && ((kv0 + BKV > L) || (causal && kv0 + BKV > it.q0 + 1));
          if (!DEEP) {   // DEEP: S_g's buffer may already hold S_{g+2}; sv (masked in place) is still intact
            load_scores(g);
            ptx::tmem_ld_wait();
            if (masked) mask_scores(kv0, it.q0, qi);
          }
          const float mt = tile_max(FAST || four);
          const float m_next = trip ? fmaxf(m_used, mt) : m_used;
          const float alpha = ptx::ex2_approx((m_used - m_next) * c);
          if (j > 0) {
            // PV_{g-1} must have landed in O.  p_free[b] completes with PV_b, PV_{b+2}, ...; S_g complete implies
            // PV_{g-3} complete (issue order) and PV_{g+1} needs this warp's P_{g+1}: the barrier is in the phase
            // of PV_{g-1} or one past it.
            ptx::mbar_wait(&bars->p_free[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[32];
              ptx::tmem_ld_32x32b_x32(t_o + hh * 32, v);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              ptx::tmem_st_32x32b_x32(t_o + hh * 32, v);
            }
            ptx::tmem_st_wait();
            if (!DEEP) {   // the O loads above overwrote nothing of sv; !DEEP redoes the tile from a fresh copy anyway
              load_scores(g);
              ptx::tmem_ld_wait();
              if (masked) mask_scores(kv0, it.q0, qi);
            }
          }
          l *= alpha;
          m_used = m_next;
          if (FAST || four) exp_pass(NG4{}, half, m_used * c, rs);
          else exp_pass(NG2{}, half, m_used * c, rs);
        }
        l += rs;
        TRS(3);
        // ---- prefetch the next S tile (possibly the next item's first) while the P hand-over is in flight
        if (FAST || j + 1 < it.n_kv || more_items) {
          ptx::mbar_wait(&bars->s_full[half ^ 1u], ((g + 1) >> 1) & 1u);
          ptx::tc_fence_after();
          load_scores(g + 1);
          if (DEEP) ptx::tmem_ld_wait();   // the arrive below also releases S buffer (g + 1) & 1
        }
        TRS(4);
        ptx::fence_proxy_async_smem();  // generic-proxy P stores -> visible to the tensor core (async proxy)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->p_full[half]);   // P_g in smem, S[g & 1] drained, O rescaled
        TRS(5);
        ptx::tmem_ld_wait();
        TRS(6);
        ++g;
      };

      // A warp whose 32 query rows all lie past the sequence end (L = 577: the fourth warp of every image's fifth query
      // tile, 5 % of the warp-tiles; a third of the text tower's) computes nothing that is ever stored: it keeps the
      // barrier protocol (wait for the next S tile, take it from TMEM so that a following live item finds its first
      // scores in registers, hand the P slot over) and skips the exponentials, the P stores and the output tile.  Its P
      // rows keep stale finite values: the rows of O they feed are never stored either.
      if (it.q0 + int(quarter) * 32 >= L) {
#pragma unroll 1
        for (int j = 0; j < it.n_kv; ++j) {
          const uint32_t half = g & 1u;
          if (j + 1 < it.n_kv || more_items) {
            ptx::mbar_wait(&bars->s_full[half ^ 1u], ((g + 1) >> 1) & 1u);
            ptx::tc_fence_after();
            load_scores(g + 1);
            ptx::tmem_ld_wait();
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars->p_full[half]);
          ++g;
        }
        continue;
      }

      int j = 0;
      tile(std::false_type{}, j++);
#pragma unroll 1
      for (; j < j_fast; ++j) tile(std::true_type{}, j);
#pragma unroll 1
      for (; j < it.n_kv; ++j) tile(std::false_type{}, j);

      // ---- item done: O / l -> bf16 -> this warp's (now idle) P rows -> one TMA tile store (rows >= L are clipped
      //      by the per-image tensor map).  PV_{g-1} complete means every PV of the item is (issue order).
      {
        const uint32_t g_last = g - 1;
        // the next completion of this barrier needs P_{g_last+2}, which this warp has not produced: unambiguous
        ptx::mbar_wait(&bars->p_free[g_last & 1u], (g_last >> 1) & 1u);
        ptx::tc_fence_after();
        const float inv = 1.0f / l;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_o + hh * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            ptx::st_shared_v4(p_row + ((uint32_t(hh * 4 + q4) ^ sw) << 4),
                              ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 0]) * inv, __uint_as_float(v[q4 * 8 + 1]) * inv),
                              ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 2]) * inv, __uint_as_float(v[q4 * 8 + 3]) * inv),
                              ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 4]) * inv, __uint_as_float(v[q4 * 8 + 5]) * inv),
                              ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 6]) * inv, __uint_as_float(v[q4 * 8 + 7]) * inv));
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();     // O has been read: orders the TMEM loads before the next p_full arrive
        __syncwarp();
        const int r0 = it.q0 + int(quarter) * 32;
        if (lane == 0 && r0 < L) {
          ptx::tma_store_3d(&tmO, p_warp, it.h * D, r0, it.b);
          ptx::bulk_commit();
        }
        store_pending = true;
      }
    }
    if (store_pending && lane == 0) ptx::bulk_wait<0>();   // smem must outlive the last store's read
  } else {
    if (BAL) ptx::reg_dec<CTL_REGS>();   // parked control slots: setmaxnreg is warpgroup-collective
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

}  // namespace attn

namespace {
long long* g_trace = nullptr;   // diagnostics only (aaclip_attention_trace)
int g_trace_cta = -1;

typedef void (*AttnKern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, int, int, int, int, int, long long*,
                         int, float);
struct AttnVariant { AttnKern fn; int smem; int ctas_per_sm; int threads; };

template <int POLY, int DEEP>
AttnVariant pick2(int minb, bool bal) {
  if (minb == 4 && bal) return {attn::attention_kernel<false, POLY, 4, 1, 1, DEEP>, attn::Lay<1>::SMEM_BYTES, 4, attn::THREADS_BAL};
  if (minb == 4) return {attn::attention_kernel<false, POLY, 4, 1, 0, DEEP>, attn::Lay<1>::SMEM_BYTES, 4, attn::THREADS};
  return {attn::attention_kernel<false, POLY, 3, 2, 0, DEEP>, attn::Lay<2>::SMEM_BYTES, 3, attn::THREADS};
}
template <int POLY>
AttnVariant pick(int minb, bool trace, bool bal, int deep) {
  if (trace) return {attn::attention_kernel<true, 0, 3, 2, 0, 0>, attn::Lay<2>::SMEM_BYTES, 3, attn::THREADS};
  return deep ? pick2<POLY, 1>(minb, bal) : pick2<POLY, 0>(minb, bal);
}
}  // namespace

int k::launch_attention(const void* qkv, void* out, int B, int L, int heads, int causal, cudaStream_t stream) {
  if (B <= 0) return host::OK;
  if (L <= 0 || heads <= 0) return host::fail(host::ERR_INVALID, "attention: L=%d heads=%d", L, heads);
  const int W = heads * attn::D;
  CUtensorMap tmQ, tmKV, tmO;
  int rc = host::make_tmap_2d(&tmQ, qkv, (uint64_t)B * L, 3 * W, 3 * W, attn::BQ);
  if (rc) return rc;
  rc = host::make_tmap_2d(&tmKV, qkv, (uint64_t)B * L, 3 * W, 3 * W, attn::BKV);
  if (rc) return rc;
  {   // output [B][L][W] bf16 as a rank-3 map: a 32-row store box is clipped at each image's last token
    const uint64_t dims[3] = {(uint64_t)W, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)W * 2, (uint64_t)L * W * 2};
    const uint32_t box[3] = {64, 32, 1};
    rc = host::make_tmap_bf16(&tmO, out, 3, dims, strides, box);
    if (rc) return rc;
  }
  // diagnostics: AACLIP_ATTN_POLY=<0..2> exponentials per 8 evaluated on the FMA pipe instead of the SFU;
  //              AACLIP_ATTN_CTAS=<3|4> resident CTAs per SM (4: single Q buffer, 80 registers)
  static int poly = getenv("AACLIP_ATTN_POLY") ? atoi(getenv("AACLIP_ATTN_POLY")) : 0;
  static int minb = getenv("AACLIP_ATTN_CTAS") ? atoi(getenv("AACLIP_ATTN_CTAS")) : 4;
  static float rescale = getenv("AACLIP_ATTN_RESCALE") ? (float)atof(getenv("AACLIP_ATTN_RESCALE")) : 8.0f;
  //              AACLIP_ATTN_BAL=<0|1> 8-warp CTAs with the control roles spread over all four schedulers
  static bool bal = getenv("AACLIP_ATTN_BAL") ? atoi(getenv("AACLIP_ATTN_BAL")) != 0 : false;
  //              AACLIP_ATTN_DEEP=<0|1> three-tile S look-ahead (see the DEEP template parameter)
  static int ord = getenv("AACLIP_ATTN_DEEP") ? atoi(getenv("AACLIP_ATTN_DEEP")) : 0;
  AttnVariant v;
  switch (poly) {
    case 0: v = pick<0>(minb, g_trace != nullptr, bal, ord); break;
    case 1: v = pick<1>(minb, g_trace != nullptr, bal, ord); break;
    case 2: v = pick<2>(minb, g_trace != nullptr, bal, ord); break;
    default: return host::fail(host::ERR_INVALID, "AACLIP_ATTN_POLY=%d out of range [0,2]", poly);
  }
  // the > 48 KB dynamic-smem opt-in is per (kernel, device context): remembered per device
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  static AttnKern configured[64][4] = {{nullptr}};
  bool seen = false;
  if (dev >= 0 && dev < 64)
    for (AttnKern kk : configured[dev]) seen = seen || (kk == v.fn);
  if (!seen) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem));
    if (dev >= 0 && dev < 64) {
      bool stored = false;
      for (AttnKern& kk : configured[dev]) if (!kk && !stored) { kk = v.fn; stored = true; }
    }
  }
  const int n_qt = (L + attn::BQ - 1) / attn::BQ;
  const long long items = (long long)n_qt * heads * B;
  if (items > INT_MAX) return host::fail(host::ERR_INVALID, "attention: %lld work items", items);
  const int sms = host::sm_count(dev);
  const int grid = (int)std::min<long long>(items, (long long)v.ctas_per_sm * (sms > 0 ? sms : 148));
  AACLIP_CUDA_CHECK(host::launch(v.fn, dim3(grid), dim3(v.threads), (size_t)v.smem, stream, tmQ, tmKV, tmO, L, heads,
                                 causal, (int)items, n_qt, g_trace, g_trace_cta, exp2f(rescale)));
  return host::OK;
}

extern "C" int aaclip_attention(const void* qkv, void* out, int B, int L, int heads, int causal, void* stream) {
  host::PointerDeviceGuard dev_guard(qkv);
  return k::launch_attention(qkv, out, B, L, heads, causal, static_cast<cudaStream_t>(stream));
}

// Diagnostics: like aaclip_attention, but CTA number `cta` also records clock64() stamps of its softmax warp 0
// (slots 0..7) and of its MMA-issuing thread (slots 16..20) per key tile of its second work item into
// trace[slot * 32 + tile] (device memory, >= 24 * 32 int64).  Used to study the pipeline; not part of the hot path.
extern "C" int aaclip_attention_trace(const void* qkv, void* out, int B, int L, int heads, int causal, long long* trace,
                                      int cta, void* stream) {
  host::PointerDeviceGuard dev_guard(qkv);
  g_trace = trace; g_trace_cta = cta;
  int rc = k::launch_attention(qkv, out, B, L, heads, causal, static_cast<cudaStream_t>(stream));
  g_trace = nullptr; g_trace_cta = -1;
  return rc;
}
