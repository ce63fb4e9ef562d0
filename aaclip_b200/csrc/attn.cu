// Multi-head self-attention core for head dim 64 on sm_100a (tcgen05 + TMEM + TMA), flash style: the
// [L, L] probability matrix the reference materialises (nn.MultiheadAttention slow path with
// need_weights=True, model/transformer.py:200,237) never leaves the SM.
//
// One CTA per (128-row query tile, head, image), two CTAs per SM; 6 warps.  The two single-thread control
// roles sit in the HIGHEST warp ids: the SMSP arbiter favours high warp ids, and a delayed MMA issue or TMA
// request stalls all four softmax warps, while the control warps themselves issue very few instructions.
//   warp 5      TMA producer: Q tile once, then K_j / V_j tiles (64 keys each) through 3-deep rings
//   warp 4      TMEM allocator + tcgen05.mma issuer:  S_j = Q K_j^T (M128 N64 K64) into a double-buffered
//               TMEM accumulator, O_j = P_j V_j (M128 N64 K64) into a second double-buffered accumulator
//   warps 0..3  softmax, thread == query row: ONE tcgen05.ld of the 64 scores of the tile into registers,
//               row max (FMNMX3 chains), exp2 on the SFU, fp32 row sum, P_j written as bf16 into
//               128B-swizzled, double-buffered smem (the A operand of the PV MMA).
// O never leaves TMEM during the key loop: PV_j accumulates into it (tcgen05.mma accumulate), and the
// online-softmax rescale is LAZY - the row keeps exponentiating against a stale max until the true max has
// grown by more than 2^8, only then is O read (tcgen05.ld), scaled and written back (tcgen05.st).  Softmax
// is shift invariant, so the result is unchanged; P entries are bounded by 256, which bf16 holds with the
// same relative precision.  The softmax warps are therefore pure SFU streams (the kernel's real bound:
// one exp2 per score on 16 SFU lanes/SM) and never wait on the tensor core in steady state
// (tools/attn_trace.py records the per-tile timeline).
// S_{j+1} is always computed while the softmax warps work on S_j, so they never wait for the tensor core
// in steady state.  Operands come straight out of the fused QKV GEMM output qkv[B*L, 3*heads*64] through
// 2-D tensor maps: rows past the image's last token are either the next image's tokens or TMA zero fill
// and are masked.  V tiles are consumed as an MN-major B operand exactly as TMA lands them (no transpose).
// The ragged last key tile (577 = 9*64 + 1) only pays for the 32-key group(s) that hold valid keys: the
// softmax skips fully masked 32-column groups and the PV MMA shortens its K extent.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace attn {

constexpr int D = 64;          // head dim
constexpr int BQ = 128;        // query rows per CTA
constexpr int BKV = 64;        // keys per tile
constexpr int NST = 3;         // K / V ring depth
constexpr int Q_BYTES = BQ * D * 2;      // 16 KB
constexpr int KV_BYTES = BKV * D * 2;    //  8 KB
constexpr int P_BYTES = BQ * BKV * 2;    // 16 KB: one 128B-swizzle atom column (64 keys) x 128 rows
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;           // S0 [0,64) S1 [64,128) O [128,192)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;
constexpr int OFF_V = OFF_K + NST * KV_BYTES;
constexpr int OFF_P = OFF_V + NST * KV_BYTES;
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;       // P is double buffered: P_j is written while PV_{j-1} still reads
constexpr int SMEM_BYTES = OFF_BAR + 256;

struct Bars {
  uint64_t q_full, k_full[NST], v_full[NST], k_empty[NST], v_empty[NST], s_full[2], p_free[2], p_full[2];
  uint32_t tmem_slot;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

template <bool TRACE_ON, int POLY>
__global__ void __launch_bounds__(THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 __nv_bfloat16* __restrict__ out, int L, int heads, int causal, long long* trace, int trace_cta,
                 float rescale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-B alignment
  Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BAR);
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = ptx::lane_id();
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int W = heads * D;
  const int q0 = qt * BQ;
  int kv_end = L;                                   // keys [0, kv_end) can be visible to this query tile
  if (causal) kv_end = min(L, q0 + BQ);
  const int n_kv = (kv_end + BKV - 1) / BKV;
  const int row_base = b * L;  // first token row of this image in qkv / out
  // optional clock64 trace of one CTA (diagnostics): [0..15][tile] softmax warp 0, [16..23][tile] MMA thread
  const int cta_linear = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const bool tracing = TRACE_ON && (trace != nullptr) && (cta_linear == trace_cta);
#define TRACE(slot, tile) do { if (TRACE_ON && tracing) trace[(slot) * 16 + (tile)] = clock64(); } while (0)

  if (warp == 5 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmKV);
    ptx::mbar_init(&bars->q_full, 1);
    for (int i = 0; i < NST; ++i) {
      ptx::mbar_init(&bars->k_full[i], 1);
      ptx::mbar_init(&bars->v_full[i], 1);
      ptx::mbar_init(&bars->k_empty[i], 1);
      ptx::mbar_init(&bars->v_empty[i], 1);
    }
    ptx::mbar_init(&bars->s_full[0], 1);
    ptx::mbar_init(&bars->s_full[1], 1);
    ptx::mbar_init(&bars->p_free[0], 1);
    ptx::mbar_init(&bars->p_free[1], 1);
    // one arrive per softmax warp.  Two alternating barriers: a fast warp may hand in P_{j+1} before a slow
    // warp has handed in P_j (nothing else orders them), and two arrivals must never land in one phase.
    ptx::mbar_init(&bars->p_full[0], 4);
    ptx::mbar_init(&bars->p_full[1], 4);
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
      ptx::mbar_arrive_expect_tx(&bars->q_full, Q_BYTES);
      ptx::tma_load_2d(smem + OFF_Q, &tmQ, &bars->q_full, h * D, row_base + q0);
      int st = 0; uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        ptx::mbar_wait(&bars->k_empty[st], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&bars->k_full[st], KV_BYTES);
        ptx::tma_load_2d(smem + OFF_K + st * KV_BYTES, &tmKV, &bars->k_full[st], W + h * D, row_base + j * BKV);
        ptx::mbar_wait(&bars->v_empty[st], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&bars->v_full[st], KV_BYTES);
        ptx::tma_load_2d(smem + OFF_V + st * KV_BYTES, &tmKV, &bars->v_full[st], 2 * W + h * D, row_base + j * BKV);
        if (++st == NST) { st = 0; ph ^= 1u; }
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
      int s_j = 0, s_st = 0; uint32_t s_ph = 0;     // next S tile to issue and its K ring slot / phase
      auto issue_s = [&]() {   // S_{s_j} -> TMEM S[s_j & 1]
        ptx::mbar_wait(&bars->k_full[s_st], s_ph);
        ptx::tc_fence_after();
        const uint32_t kb = k_lo + uint32_t(s_st) * (KV_BYTES >> 4);
        const uint32_t t_s = tmem_base + uint32_t(s_j & 1) * 64u;
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::mma_f16_ss<1>(t_s, desc(q_lo + k * 2), desc(kb + k * 2), idesc_s, k != 0 ? 1u : 0u);
        ptx::mma_commit(&bars->s_full[s_j & 1]);
        ptx::mma_commit(&bars->k_empty[s_st]);
        ++s_j;
        if (++s_st == NST) { s_st = 0; s_ph ^= 1u; }
      };
      ptx::mbar_wait(&bars->q_full, 0);
      issue_s();
      if (n_kv > 1) issue_s();
      int st = 0; uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        TRACE(16, j);
        ptx::mbar_wait(&bars->p_full[j & 1], (j >> 1) & 1);  // P_j in smem, S[j&1] drained, O rescaled if needed
        TRACE(17, j);
        ptx::mbar_wait(&bars->v_full[st], ph);
        ptx::tc_fence_after();
        TRACE(18, j);
        const uint32_t pb = p_lo + uint32_t(j & 1) * (P_BYTES >> 4);
        const uint32_t vb = v_lo + uint32_t(st) * (KV_BYTES >> 4);
        const uint32_t t_o = tmem_base + 128u;
        const uint32_t acc0 = j > 0 ? 1u : 0u;   // O accumulates over key tiles
        // A: P tile = one 64-key swizzle atom column (128 rows x 128 B), 16 keys = +32 B;
        // B: V tile as TMA landed it (MN-major), 16 keys = 16 rows x 128 B = +2048 B
        if (j + 1 < n_kv || kv_end - j * BKV >= BKV) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)
            ptx::mma_f16_ss<1>(t_o, desc(pb + k * 2), desc(vb + k * 128), idesc_o, k != 0 ? 1u : acc0);
        } else {   // ragged last tile: skip 16-key groups that are fully hidden
          const int ksteps = (kv_end - j * BKV + 15) >> 4;
          for (int k = 0; k < ksteps; ++k)
            ptx::mma_f16_ss<1>(t_o, desc(pb + k * 2), desc(vb + k * 128), idesc_o, k != 0 ? 1u : acc0);
        }
        ptx::mma_commit(&bars->v_empty[st]);
        ptx::mma_commit(&bars->p_free[j & 1]);   // PV_j has landed in O and has finished reading P[j & 1]
        if (++st == NST) { st = 0; ph ^= 1u; }
        TRACE(19, j);
        if (s_j < n_kv) issue_s();
        TRACE(20, j);
      }
    }
  } else {
    // ===================================================== softmax warps: thread == query row
    const uint32_t quarter = warp & 3u;
    const int row = int(quarter * 32u + lane);
    const uint32_t t_lane = tmem_base + ((quarter * 32u) << 16);
    const uint32_t t_o = t_lane + 128u;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const float RESCALE_LOG2 = rescale_log2;       // lazy rescale: tolerate p up to 2^8 before touching O
    uint8_t* p_row = smem + OFF_P + row * 128;
    const uint32_t sw = uint32_t(row & 7);
    float m_used = 0.f, l = 0.f;                   // stabiliser in use (<= true running max + 2^8), row sum
    const int qi = q0 + row;
    uint32_t sv[64];                               // raw scores of the current tile (this thread's row)
    uint32_t (&sv_lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[0]);
    uint32_t (&sv_hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[32]);

    // tile geometry (warp-uniform except `limit`)
    auto tile_keys_of = [&](int j) { return min(BKV, kv_end - j * BKV); };   // keys any row of the tile may see
    // issue the TMEM loads of S_j (no wait): 32 or 64 columns depending on how many keys the tile holds
    auto load_scores = [&](int j) {
      const uint32_t t_s = t_lane + uint32_t(j & 1) * 64u;
      ptx::tmem_ld_32x32b_x32(t_s, sv_lo);
      if (tile_keys_of(j) > 32) ptx::tmem_ld_32x32b_x32(t_s + 32, sv_hi);
    };
    // hidden keys -> -inf -> p = 0 (last key tile / causal diagonal only)
    auto mask_scores = [&](int j) {
      const int kv0 = j * BKV;
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
    // POLY of every 8 exponentials are evaluated on the FMA pipe (Cody-Waite + cubic) to unload the SFU.
    auto exp_pass = [&](uint8_t* pr, float mc, int n_groups, float& rowsum, float& rowmax) {
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (g < n_groups) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x = fmaf(__uint_as_float(sv[8 * g + i]), c, -mc);
            e[i] = (i >= 8 - POLY) ? ptx::ex2_poly3(x) : ptx::ex2_approx(x);
          }
          mx4[g & 3] = fmaxf(mx4[g & 3], fmaxf(fmaxf(__uint_as_float(sv[8 * g + 0]), __uint_as_float(sv[8 * g + 1])),
                                               fmaxf(__uint_as_float(sv[8 * g + 2]), __uint_as_float(sv[8 * g + 3]))));
          mx4[(g + 2) & 3] = fmaxf(mx4[(g + 2) & 3],
                                   fmaxf(fmaxf(__uint_as_float(sv[8 * g + 4]), __uint_as_float(sv[8 * g + 5])),
                                         fmaxf(__uint_as_float(sv[8 * g + 6]), __uint_as_float(sv[8 * g + 7]))));
          rs4[g & 3] += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
          ptx::st_shared_v4(pr + ((uint32_t(g) ^ sw) << 4), ptx::pack_bf16x2(e[0], e[1]), ptx::pack_bf16x2(e[2], e[3]),
                            ptx::pack_bf16x2(e[4], e[5]), ptx::pack_bf16x2(e[6], e[7]));
        }
      }
      rowsum = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      rowmax = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    };

    const bool tr = TRACE_ON && tracing && warp == 0 && lane == 0;
#define TRS(slot) do { if (TRACE_ON && tr) trace[(slot) * 16 + j] = clock64(); } while (0)
    ptx::mbar_wait(&bars->s_full[0], 0);
    ptx::tc_fence_after();
    load_scores(0);
    ptx::tmem_ld_wait();

    for (int j = 0; j < n_kv; ++j) {
      TRS(0);
      const int n_groups = tile_keys_of(j) > 32 ? 8 : 4;   // 8-key groups holding visible keys (warp-uniform)
      uint8_t* pr = p_row + (j & 1) * P_BYTES;
      // P buffer j & 1 was last read by PV_{j-2}.  The tensor pipe retires one thread's MMAs in issue order and
      // S_j was issued after PV_{j-2}, so having seen s_full for S_j implies that buffer is free: no wait here.
      mask_scores(j);
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
          // PV_{j-1} must have landed in O.  p_free[b] phase k completes with PV_{b+2k}; S_j complete implies
          // PV_{j-3} complete (issue order), so the barrier is in phase (j-1)>>1 or one past it: the parity
          // wait is unambiguous.
          ptx::mbar_wait(&bars->p_free[(j - 1) & 1], ((j - 1) >> 1) & 1);
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
        // redo the tile against the new stabiliser (S_j is still in TMEM: it is released by the p_full arrive)
        load_scores(j);
        ptx::tmem_ld_wait();
        mask_scores(j);
        exp_pass(pr, m_used * c, n_groups, rs, mt);
      }
      l += rs;
      TRS(3);
      // ---- prefetch S_{j+1} into registers while the P hand-over is in flight
      if (j + 1 < n_kv) {
        ptx::mbar_wait(&bars->s_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        ptx::tc_fence_after();
        load_scores(j + 1);
      }
      TRS(4);
      ptx::fence_proxy_async_smem();  // generic-proxy P stores -> visible to the tensor core (async proxy)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->p_full[j & 1]);   // P_j in smem, S[j & 1] drained, O rescaled if needed
      TRS(5);
      ptx::tmem_ld_wait();
      TRS(6);
    }
    // ---- O is complete once the last PV has landed: normalise by the row sum and store
    ptx::mbar_wait(&bars->p_free[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
    ptx::tc_fence_after();
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = out + (size_t)(row_base + qi) * W + h * D;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(t_o + hh * 32, v);
      ptx::tmem_ld_wait();
      if (qi < L) {
        uint4* dst = reinterpret_cast<uint4*>(orow + hh * 32);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 w;
          w.x = ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 0]) * inv, __uint_as_float(v[q4 * 8 + 1]) * inv);
          w.y = ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 2]) * inv, __uint_as_float(v[q4 * 8 + 3]) * inv);
          w.z = ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 4]) * inv, __uint_as_float(v[q4 * 8 + 5]) * inv);
          w.w = ptx::pack_bf16x2(__uint_as_float(v[q4 * 8 + 6]) * inv, __uint_as_float(v[q4 * 8 + 7]) * inv);
          dst[q4] = w;
        }
      }
    }
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
  typedef void (*Kern)(const CUtensorMap, const CUtensorMap, __nv_bfloat16*, int, int, int, long long*, int, float);
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
  dim3 grid((L + attn::BQ - 1) / attn::BQ, heads, B);
  kern<<<grid, attn::THREADS, attn::SMEM_BYTES, stream>>>(tmQ, tmKV, static_cast<__nv_bfloat16*>(out), L, heads, causal,
                                                          g_trace, g_trace_cta, rescale);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

extern "C" int aaclip_attention(const void* qkv, void* out, int B, int L, int heads, int causal, void* stream) {
  return k::launch_attention(qkv, out, B, L, heads, causal, static_cast<cudaStream_t>(stream));
}

// Diagnostics: like aaclip_attention, but CTA number `cta` also records clock64() stamps of its softmax warp 0
// (slots 0..7) and of its MMA-issuing thread (slots 16..20) per key tile into trace[slot * 16 + tile] (device
// memory, >= 24 * 16 int64).  Used to study the pipeline; not part of the hot path.
extern "C" int aaclip_attention_trace(const void* qkv, void* out, int B, int L, int heads, int causal, long long* trace,
                                      int cta, void* stream) {
  g_trace = trace; g_trace_cta = cta;
  int rc = k::launch_attention(qkv, out, B, L, heads, causal, static_cast<cudaStream_t>(stream));
  g_trace = nullptr; g_trace_cta = -1;
  return rc;
}
