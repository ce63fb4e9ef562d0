// Anomaly-map head, test modes (the A7 contract of SURVEY 8(a)):
//
//   calculate_similarity_map(test=True) x n_levels + cat + sum(1) + image score + map extrema
//   (forward_utils.py:196-216, test.py:83-93, forward_utils.py:241-252)
//
// as a persistent HBM-streaming kernel (patch tokens -> one scalar per patch) chained by programmatic dependent launch
// to the per-image tail (head.cu: maps_from_dots_kernel - blur, upsample, map rows, extrema).  The pair is bound by
// reading the normalised patch tokens (3.5 MB / image as bf16, 7.1 MB as fp32) and writing the fp32 map (0.45 MB).
//
//   * one CTA per SM; ONE producer thread streams a whole (unit, level) slab - 64 bf16 / 32 fp32 consecutive patch rows
//     of one level, 96 KB, contiguous in HBM - per cp.async.bulk (TMA, L2 evict-first) into a two-stage shared-memory
//     ring with mbarrier transaction counts: ~28 MB in flight on the device, independent of the consumer warps;
//   * 8 consumer warps take the dot of every token with (T1 - T0) straight from shared memory (conflict-free 16-byte
//     reads, the lane's 24 anchor differences in registers, one transposed shuffle reduction per slab); warp w owns rows
//     [R w, R w + R) of every slab (R = 8 / 4) and every warp passes every stage in order, so no parity wait can alias;
//   * units are dealt round robin (image-major: CTA i takes units i, i + grid, ...); a unit leaves its level-summed
//     scalars in a small global array with plain stores - no fences, atomics or CTA barriers on the streaming path.
//
// History (round 2, ncu r2g): a first version ran the image tail inside this kernel on the CTA that finished an
// image's last unit (atomic counter).  Every unit then paid a device-scope fence + barrier + atomic round trip (25 %
// of the consumer warps' time), one 12 KB copy per (warp, level) made the producer thread the bottleneck (24 % of the
// time waiting for slabs at 28 % DRAM utilisation), and the last images' tails ran alone: 98 us at batch 64.  Before
// that, per-warp slots shared between warps that skipped each other's stages aliased mbarrier phases (a parity wait
// only tells the current phase from the previous one): garbage metadata, illegal addresses.
#include <limits.h>
#include <stdarg.h>
#include <algorithm>
#include "common.cuh"
#include "head_epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace {

constexpr int HS_E = 768;
constexpr int HS_WARPS = 8;                       // consumer warps
constexpr int HS_WARP_BYTES = 12288;              // one warp's share of a slab: 8 bf16 / 4 fp32 token rows
constexpr int HS_STAGE_BYTES = HS_WARPS * HS_WARP_BYTES;   // one (unit, level) slab: 96 KB
constexpr int HS_STAGES = 2;
constexpr int HS_THREADS = (HS_WARPS + 1) * 32;   // + the producer warp
constexpr int HS_MAX_LEVELS = 8;
constexpr int HS_SMEM_BYTES = HS_STAGES * HS_STAGE_BYTES + 2 * HS_STAGES * 8 + HS_STAGES * 16 + 64;

struct HsArgs {
  const void* seg[HS_MAX_LEVELS];
  int n_levels;
  const float* anchors;     // [E, 2]
  const float* det;         // [B, E] or null
  int B, P;
  float* scores;            // [B] or null
  float* m_glob;            // [B, P] level-summed per-patch scalars (workspace)
  int units_per_image, total_units;
};

struct __align__(16) HsMeta { int b, pbase, nrows, level; };

// sums v[0..R) over the warp with a transposed reduction (9 shuffles for R = 8, 6 for R = 4, instead of 5 R):
// afterwards every lane holds the total of v[lane / (32 / R)]
template <int R>
__device__ __forceinline__ float reduce_rows(float (&v)[R], int lane) {
  static_assert(R == 8 || R == 4, "slab rows");
  float c[2];
  if constexpr (R == 8) {
    const bool b4 = lane & 16, b3 = lane & 8;
    float a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float send = b4 ? v[j] : v[j + 4], keep = b4 ? v[j + 4] : v[j];
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float send = b3 ? a[j] : a[j + 2], keep = b3 ? a[j + 2] : a[j];
      c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const bool b2 = lane & 4;
    const float send = b2 ? c[0] : c[1], keep = b2 ? c[1] : c[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);   // row = bit2 + 2 bit3 + 4 bit4 = lane >> 2
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
  } else {
    const bool b4 = lane & 16, b3 = lane & 8;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float send = b4 ? v[j] : v[j + 2], keep = b4 ? v[j + 2] : v[j];
      c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const float send = b3 ? c[0] : c[1], keep = b3 ? c[1] : c[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 8);   // row = bit3 + 2 bit4 = lane >> 3
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
  }
}

template <bool BF16>
__global__ void __launch_bounds__(HS_THREADS, 1) head_stream_kernel(const HsArgs a) {
  constexpr int ROW_BYTES = HS_E * (BF16 ? 2 : 4);
  constexpr int ROWS = HS_WARP_BYTES / ROW_BYTES;   // patch rows per warp and slab: 8 (bf16) / 4 (fp32)
  constexpr int UNIT = ROWS * HS_WARPS;             // patches per unit: 64 / 32
  constexpr int CH = ROW_BYTES / 512;               // 16-byte reads per lane and row: 3 (bf16) / 6 (fp32)
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;                             // HS_STAGES x HS_STAGE_BYTES
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + HS_STAGES * HS_STAGE_BYTES);
  uint64_t* empty = full + HS_STAGES;
  HsMeta* meta = reinterpret_cast<HsMeta*>(empty + HS_STAGES);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < HS_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], HS_WARPS); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  ptx::grid_dep_sync();

  if (warp == HS_WARPS) {
    // ===================================================== producer: units -> (unit, level) slabs -> the ring
    if (ptx::elect_one()) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      int slot = 0; uint32_t ph = 0;
      for (unsigned int u = blockIdx.x;; u += gridDim.x) {
        const bool valid = u < (unsigned int)a.total_units;
        const int b = valid ? int(u / (unsigned int)a.units_per_image) : -1;
        const int pbase = valid ? int(u % (unsigned int)a.units_per_image) * UNIT : 0;
        const int nrows = valid ? min(UNIT, a.P - pbase) : 0;
        const int n_slabs = valid ? a.n_levels : 1;     // the terminator is one empty slab
        for (int l = 0; l < n_slabs; ++l) {
          ptx::mbar_wait_spin(&empty[slot], ph ^ 1u);
          meta[slot] = HsMeta{b, pbase, nrows, l};
          if (nrows > 0) {
            const uint32_t bytes = uint32_t(nrows) * ROW_BYTES;
            const uint8_t* src = static_cast<const uint8_t*>(a.seg[l]) + ((size_t)b * a.P + pbase) * ROW_BYTES;
            ptx::mbar_arrive_expect_tx(&full[slot], bytes);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                ::"r"(ptx::smem_u32(ring + (size_t)slot * HS_STAGE_BYTES)), "l"(src), "r"(bytes),
                  "r"(ptx::smem_u32(&full[slot])), "l"(policy) : "memory");
          } else {
            ptx::mbar_arrive(&full[slot]);   // terminator: metadata only
          }
          if (++slot == HS_STAGES) { slot = 0; ph ^= 1u; }
        }
        if (!valid) break;
      }
    }
    __syncwarp();
  } else {
    // ===================================================== consumers: every warp passes every stage, in order
    float td[CH * (BF16 ? 8 : 4)];   // T[c][1] - T[c][0] of this lane's columns: the test-mode map needs only s1 - s0
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
      for (int j = 0; j < (BF16 ? 8 : 4); ++j) {
        const int c = BF16 ? (i * 256 + lane * 8 + j) : (i * 128 + lane * 4 + j);
        const float2 t2 = __ldg(reinterpret_cast<const float2*>(a.anchors) + c);
        td[i * (BF16 ? 8 : 4) + j] = t2.y - t2.x;
      }
    constexpr int LPR = 32 / ROWS;     // lanes that end up with the same row's total
    float acc = 0.f;                   // level sum of patch row (lane / LPR) of this warp's share of the unit
    int slot = 0; uint32_t ph = 0;
    for (;;) {
      ptx::mbar_wait(&full[slot], ph);
      const int4 mt = *reinterpret_cast<const int4*>(&meta[slot]);   // ordered after the wait (acquire + "memory" clobber)
      const int b = mt.x, pbase = mt.y, level = mt.w;
      if (b < 0) break;                // terminator
      const int nrows = min(ROWS, mt.z - warp * ROWS);   // this warp's rows of the slab (<= 0: none)
      // the image score is taken by whoever streams the image's first patches (test.py:83-84)
      if (level == 0 && pbase == 0 && warp == 0 && a.scores != nullptr)
        headepi::image_score_unrolled<HS_E / 32>(a.det, a.anchors, b, lane, a.scores);
      if (level == 0) acc = 0.f;
      if (nrows > 0) {
        const uint8_t* rows = ring + (size_t)slot * HS_STAGE_BYTES + (size_t)warp * HS_WARP_BYTES + lane * 16;
        float d[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const uint4 raw = *reinterpret_cast<const uint4*>(rows + r * ROW_BYTES + c * 512);
            if constexpr (BF16) {
              const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                s0 = fmaf(__uint_as_float(w4[k] << 16), td[c * 8 + 2 * k], s0);
                s1 = fmaf(__uint_as_float(w4[k] & 0xffff0000u), td[c * 8 + 2 * k + 1], s1);
              }
            } else {
              s0 = fmaf(__uint_as_float(raw.x), td[c * 4 + 0], s0);
              s1 = fmaf(__uint_as_float(raw.y), td[c * 4 + 1], s1);
              s0 = fmaf(__uint_as_float(raw.z), td[c * 4 + 2], s0);
              s1 = fmaf(__uint_as_float(raw.w), td[c * 4 + 3], s1);
            }
          }
          d[r] = s0 + s1;   // rows >= nrows hold stale bytes of an earlier slab: computed and dropped below
        }
        const float dsum = reduce_rows<ROWS>(d, lane);
        acc += fmaf(50.0f, dsum, 0.5f);   // (s1 + 1 - s0) / 2 of this level with s = 100 <f, T>   (test.py:85)
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty[slot]);
      if (level == a.n_levels - 1) {   // unit done: its level-summed scalars (read by the tail kernel, next launch)
        const int r = lane / LPR;
        if ((lane % LPR) == 0 && r < nrows) a.m_glob[(size_t)b * a.P + pbase + warp * ROWS + r] = acc;
      }
      if (++slot == HS_STAGES) { slot = 0; ph ^= 1u; }
    }
  }
  __syncthreads();
}

}  // namespace

long long k::head_stream_workspace_bytes(int B, int P) {
  return (long long)B * P * 4;
}

bool k::head_stream_supported(int n_levels, int seg_is_bf16, int anchors_batched, int E, int P, int G, int S, int mode,
                              const void* const* seg) {
  (void)seg_is_bf16; (void)P; (void)S;
  if (mode != AACLIP_HEAD_TEST_INDUSTRIAL && mode != AACLIP_HEAD_TEST_MEDICAL) return false;
  if (anchors_batched || E != HS_E || n_levels < 1 || n_levels > HS_MAX_LEVELS) return false;
  const int pad = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 3 : 4;
  if (G <= pad) return false;
  for (int l = 0; l < n_levels; ++l)
    if (seg[l] == nullptr || (reinterpret_cast<uintptr_t>(seg[l]) & 15u) != 0) return false;   // cp.async.bulk alignment
  return true;
}

// tokens -> m_glob [B, P] (level-summed per-patch scalars, at the start of `workspace`) and scores [B]; the caller
// chains k::launch_maps_from_dots(..., pdl = true) for the image tails.
int k::launch_head_stream(const void* const* seg, int n_levels, int seg_is_bf16, const float* anchors, const float* det,
                          int B, int P, float* scores, void* workspace, cudaStream_t stream) {
  if (B <= 0) return host::OK;
  if (!workspace) return host::fail(host::ERR_INVALID, "head: workspace missing");
  HsArgs a;
  memset(&a, 0, sizeof a);
  for (int l = 0; l < n_levels; ++l) a.seg[l] = seg[l];
  a.n_levels = n_levels; a.anchors = anchors; a.det = det; a.B = B; a.P = P;
  a.scores = (det != nullptr) ? scores : nullptr;
  a.m_glob = static_cast<float*>(workspace);
  const int unit = HS_WARPS * (HS_WARP_BYTES / (HS_E * (seg_is_bf16 ? 2 : 4)));   // 64 (bf16) / 32 (fp32) patches
  a.units_per_image = (P + unit - 1) / unit;
  const long long total = (long long)B * a.units_per_image;
  if (total > INT_MAX) return host::fail(host::ERR_INVALID, "head: %lld units", total);
  a.total_units = (int)total;
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  const int sms = host::sm_count(dev) > 0 ? host::sm_count(dev) : 148;
  const int grid = (int)std::min<long long>(sms, total);
  // the > 48 KB dynamic-smem opt-in belongs to the current device's context: once per (kernel, device)
  static bool configured[2][64] = {{false}};
  const bool known = dev >= 0 && dev < 64 && configured[seg_is_bf16 ? 1 : 0][dev];
  if (seg_is_bf16) {
    if (!known) AACLIP_CUDA_CHECK(cudaFuncSetAttribute(head_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HS_SMEM_BYTES));
    AACLIP_CUDA_CHECK(host::launch(head_stream_kernel<true>, dim3(grid), dim3(HS_THREADS), HS_SMEM_BYTES, stream, a));
  } else {
    if (!known) AACLIP_CUDA_CHECK(cudaFuncSetAttribute(head_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HS_SMEM_BYTES));
    AACLIP_CUDA_CHECK(host::launch(head_stream_kernel<false>, dim3(grid), dim3(HS_THREADS), HS_SMEM_BYTES, stream, a));
  }
  if (dev >= 0 && dev < 64) configured[seg_is_bf16 ? 1 : 0][dev] = true;
  return host::OK;
}
