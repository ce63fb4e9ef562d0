// Anomaly-map head, test modes, as ONE persistent HBM-streaming kernel (the A7 contract of SURVEY 8(a)):
//
//   calculate_similarity_map(test=True) x n_levels + cat + sum(1) + image score + map extrema
//   (forward_utils.py:196-216, test.py:83-93, forward_utils.py:241-252)
//
// The kernel is bound by reading the normalised patch tokens (3.5 MB / image as bf16, 7.1 MB as fp32) and writing
// the fp32 map (0.45 MB / image).  Its predecessor (a cluster of 8 CTAs per image, loads issued from registers)
// stopped at 62 % of HBM peak at batch 64: every warp alternated between waiting for its loads and 400 instructions
// of arithmetic, and the load and store phases of all images ran in lockstep.  Here
//
//   * one CTA per SM; ONE producer thread streams 8-patch slabs (12 / 24 KB, contiguous in HBM) through a 192 KB
//     shared-memory ring with cp.async.bulk (TMA, L2 evict-first) + mbarrier transaction counts: ~28 MB in flight
//     on the device, independent of what the consumer warps are doing;
//   * 8 consumer warps; warp w owns patches [8w, 8w+8) of the CTA's current 64-patch unit and takes the dot of every
//     token with (T1 - T0) straight from shared memory (conflict-free 16-byte reads, the lane's 24 anchor
//     differences in registers, one 9-shuffle transposed reduction per slab);
//   * units are drawn dynamically (image-major) by the producer; a unit leaves 64 scalars in a small global array;
//     the CTA that completes an image's LAST unit (atomic counter) runs that image's tail - blur, upsample, map
//     rows, extrema, score (head_epilogue.cuh) - while its producer keeps prefetching and every other CTA keeps
//     streaming, so map stores overlap token loads even when the whole batch is a single wave.
#include <stdarg.h>
#include <algorithm>
#include "common.cuh"
#include "head_epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "../../include/aaclip_b200.h"

namespace {

constexpr int HS_E = 768;
constexpr int HS_ROWS = 8;                        // patches per slab (one consumer warp's share of a unit)
constexpr int HS_WARPS = headepi::WARPS;          // consumer warps
constexpr int HS_UNIT = HS_ROWS * HS_WARPS;       // patches per unit
constexpr int HS_THREADS = (HS_WARPS + 1) * 32;   // + the producer warp
constexpr int HS_MAX_STAGES = 16;
constexpr int HS_MAX_LEVELS = 8;
constexpr int HS_SMEM_BUDGET = 227 * 1024;

struct HsArgs {
  const void* seg[HS_MAX_LEVELS];
  int n_levels;
  const float* anchors;     // [E, 2]
  const float* det;         // [B, E] or null
  int B, P, G, S, ksize;
  float sigma;
  float* maps;              // [B, S, S]
  float* scores;            // [B] or null
  float* minmax;            // [B, 2] or null
  float* m_glob;            // [B, P] level-summed per-patch scalars (workspace)
  unsigned int* done;       // [B] units finished per image (workspace, zeroed before the launch)
  unsigned int* next_unit;  // [1] dynamic unit counter (workspace, zeroed before the launch)
  int units_per_image, total_units, n_stages;
};

struct __align__(16) HsMeta { int b, p0, nrows, level; };

__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// sums v[0..7] over the warp: afterwards every lane holds the total of v[(lane >> 2) & 7]  (9 shuffles instead of 40)
__device__ __forceinline__ float reduce8(float (&v)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float a[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = b4 ? v[j] : v[j + 4], keep = b4 ? v[j + 4] : v[j];
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float c[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = b3 ? a[j] : a[j + 2], keep = b3 ? a[j + 2] : a[j];
    c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float send = b2 ? c[0] : c[1], keep = b2 ? c[1] : c[0];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

template <bool BF16>
__global__ void __launch_bounds__(HS_THREADS, 1) head_stream_kernel(const HsArgs a) {
  constexpr int ROW_BYTES = HS_E * (BF16 ? 2 : 4);
  constexpr int STAGE_BYTES = HS_ROWS * ROW_BYTES;
  constexpr int CH = ROW_BYTES / 512;             // 16-byte reads per lane and row: 3 (bf16) / 6 (fp32)
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.n_stages * STAGE_BYTES);
  uint64_t* empty = full + HS_MAX_STAGES;
  HsMeta* meta = reinterpret_cast<HsMeta*>(empty + HS_MAX_STAGES);
  int* s_last = reinterpret_cast<int*>(meta + HS_MAX_STAGES);
  const headepi::Smem epi = headepi::carve(reinterpret_cast<float*>(s_last + 4), a.P, a.G, a.S);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < a.n_stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp < HS_WARPS) headepi::setup(epi, tid, headepi::THREADS, a.G, a.S, a.ksize, a.sigma);
  __syncthreads();
  ptx::grid_dep_sync();

  if (warp == HS_WARPS) {
    // ===================================================== producer: units -> slabs -> the ring
    if (ptx::elect_one()) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      unsigned int next = atomicAdd(a.next_unit, 1u);
      uint32_t i = 0;
      int slot = 0; uint32_t ph = 0;
      for (;;) {
        const unsigned int u = next;
        const bool valid = u < (unsigned int)a.total_units;
        if (valid) next = atomicAdd(a.next_unit, 1u);   // drawn one unit ahead: its latency hides behind this unit
        const int b = valid ? int(u / (unsigned int)a.units_per_image) : -1;
        const int pbase = valid ? int(u % (unsigned int)a.units_per_image) * HS_UNIT : 0;
        for (int w = 0; w < HS_WARPS; ++w) {
          const int p0 = pbase + w * HS_ROWS;
          const int nrows = valid ? max(0, min(HS_ROWS, a.P - p0)) : 0;
          for (int l = 0; l < a.n_levels; ++l, ++i) {
            ptx::mbar_wait_spin(&empty[slot], ph ^ 1u);
            meta[slot] = HsMeta{b, p0, nrows, l};
            if (nrows > 0) {
              const uint32_t bytes = uint32_t(nrows) * ROW_BYTES;
              const uint8_t* src = static_cast<const uint8_t*>(a.seg[l]) + ((size_t)b * a.P + p0) * ROW_BYTES;
              ptx::mbar_arrive_expect_tx(&full[slot], bytes);
              asm volatile(
                  "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                  ::"r"(ptx::smem_u32(ring + (size_t)slot * STAGE_BYTES)), "l"(src), "r"(bytes),
                    "r"(ptx::smem_u32(&full[slot])), "l"(policy) : "memory");
            } else {
              ptx::mbar_arrive(&full[slot]);   // empty slab (ragged last unit / terminator round): metadata only
            }
            if (++slot == a.n_stages) { slot = 0; ph ^= 1u; }
          }
        }
        if (!valid) break;   // that was the terminator round (b = -1 for every consumer warp)
      }
    }
    __syncwarp();
  } else {
    // ===================================================== consumers: dots, unit hand-over, image tails
    float td[CH * (BF16 ? 8 : 4)];   // T[c][1] - T[c][0] of this lane's columns: the test-mode map needs only s1 - s0
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
      for (int j = 0; j < (BF16 ? 8 : 4); ++j) {
        const int c = BF16 ? (i * 256 + lane * 8 + j) : (i * 128 + lane * 4 + j);
        const float2 t2 = __ldg(reinterpret_cast<const float2*>(a.anchors) + c);
        td[i * (BF16 ? 8 : 4) + j] = t2.y - t2.x;
      }
    const int NL = a.n_levels;
    for (uint32_t n = 0;; ++n) {
      float acc = 0.f;                 // level sum of patch (lane >> 2) & 7 of this warp's slab
      int b = -1, p0 = 0, nrows = 0;
      uint32_t i = (n * HS_WARPS + uint32_t(warp)) * uint32_t(NL);
      for (int l = 0; l < NL; ++l, ++i) {
        const uint32_t slot = i % uint32_t(a.n_stages), ph = (i / uint32_t(a.n_stages)) & 1u;
        ptx::mbar_wait(&full[slot], ph);
        const int4 mt = *reinterpret_cast<const int4*>(&meta[slot]);   // ordered after the wait (acquire + "memory" clobber)
        b = mt.x; p0 = mt.y; nrows = mt.z;
        if (nrows > 0) {
          const uint8_t* slab = ring + (size_t)slot * STAGE_BYTES + lane * 16;
          float d[HS_ROWS];
#pragma unroll
          for (int r = 0; r < HS_ROWS; ++r) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const uint4 raw = *reinterpret_cast<const uint4*>(slab + r * ROW_BYTES + c * 512);
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
          const float dsum = reduce8(d, lane);
          acc += fmaf(50.0f, dsum, 0.5f);   // (s1 + 1 - s0) / 2 of this level with s = 100 <f, T>   (test.py:85)
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[slot]);
      }
      if (b < 0) break;   // terminator round
      const int r = lane >> 2;
      if ((lane & 3) == 0 && r < nrows) a.m_glob[(size_t)b * a.P + p0 + r] = acc;
      __threadfence();
      named_bar_sync(1, headepi::THREADS);
      if (tid == 0) {
        __threadfence();
        *s_last = (atomicAdd(&a.done[b], 1u) == (unsigned int)(a.units_per_image - 1)) ? 1 : 0;
      }
      named_bar_sync(1, headepi::THREADS);
      if (*reinterpret_cast<volatile int*>(s_last)) {
        // every unit of image b has landed in m_glob (ours and the other CTAs'): this CTA runs the image's tail
        __threadfence();
        for (int k = tid; k < a.P; k += headepi::THREADS) epi.m[k] = __ldcg(a.m_glob + (size_t)b * a.P + k);
        if (warp == 0 && a.scores != nullptr) headepi::image_score(a.det, a.anchors, HS_E, b, lane, a.scores);
        named_bar_sync(1, headepi::THREADS);
        headepi::image(epi, tid, b, a.G, a.S, a.ksize, a.maps, a.minmax, [] { named_bar_sync(1, headepi::THREADS); });
        named_bar_sync(1, headepi::THREADS);   // s_last / epilogue smem are rewritten by the next unit
      }
    }
  }
  __syncthreads();
}

size_t hs_dyn_smem(int n_stages, int stage_bytes, int P, int G, int S) {
  return (size_t)n_stages * stage_bytes + 2 * HS_MAX_STAGES * sizeof(uint64_t) + HS_MAX_STAGES * sizeof(HsMeta) + 16 +
         headepi::smem_floats(P, G, S) * sizeof(float);
}

}  // namespace

long long k::head_stream_workspace_bytes(int B, int P) {
  return ((long long)B * P + B + 4) * 4;
}

bool k::head_stream_supported(int n_levels, int seg_is_bf16, int anchors_batched, int E, int P, int G, int S, int mode,
                              const void* const* seg) {
  if (mode != AACLIP_HEAD_TEST_INDUSTRIAL && mode != AACLIP_HEAD_TEST_MEDICAL) return false;
  if (anchors_batched || E != HS_E || n_levels < 1 || n_levels > HS_MAX_LEVELS) return false;
  const int pad = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 3 : 4;
  if (G <= pad || S < 1) return false;
  for (int l = 0; l < n_levels; ++l)
    if (seg[l] == nullptr || (reinterpret_cast<uintptr_t>(seg[l]) & 15u) != 0) return false;   // cp.async.bulk alignment
  const int stage = HS_ROWS * HS_E * (seg_is_bf16 ? 2 : 4);
  return hs_dyn_smem(2, stage, P, G, S) <= (size_t)HS_SMEM_BUDGET;
}

int k::launch_head_stream(const void* const* seg, int n_levels, int seg_is_bf16, const float* anchors, const float* det,
                          int B, int P, int G, int S, int mode, float* maps, float* scores, float* minmax, void* workspace,
                          cudaStream_t stream) {
  if (B <= 0) return host::OK;
  if (!workspace) return host::fail(host::ERR_INVALID, "head: workspace missing");
  HsArgs a;
  memset(&a, 0, sizeof a);
  for (int l = 0; l < n_levels; ++l) a.seg[l] = seg[l];
  a.n_levels = n_levels; a.anchors = anchors; a.det = det; a.B = B; a.P = P; a.G = G; a.S = S;
  a.ksize = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 7 : 9;
  a.sigma = (mode == AACLIP_HEAD_TEST_INDUSTRIAL) ? 1.0f : 1.5f;
  a.maps = maps; a.scores = (det != nullptr) ? scores : nullptr; a.minmax = minmax;
  a.m_glob = static_cast<float*>(workspace);
  a.done = reinterpret_cast<unsigned int*>(a.m_glob + (size_t)B * P);
  a.next_unit = a.done + B;
  a.units_per_image = (P + HS_UNIT - 1) / HS_UNIT;
  const long long total = (long long)B * a.units_per_image;
  if (total > INT_MAX) return host::fail(host::ERR_INVALID, "head: %lld units", total);
  a.total_units = (int)total;
  const int stage = HS_ROWS * HS_E * (seg_is_bf16 ? 2 : 4);
  int n_stages = HS_MAX_STAGES;
  while (n_stages > 2 && hs_dyn_smem(n_stages, stage, P, G, S) > (size_t)HS_SMEM_BUDGET) --n_stages;
  if (seg_is_bf16 == 0) n_stages = std::min(n_stages, 8);   // 24 KB slabs: 8 x 24 KB = the same 192 KB in flight
  a.n_stages = n_stages;
  const size_t smem = hs_dyn_smem(n_stages, stage, P, G, S);
  if (smem > (size_t)HS_SMEM_BUDGET) return host::fail(host::ERR_INVALID, "head: grid %d / size %d needs %zu B smem", G, S, smem);
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  const int sms = host::sm_count(dev) > 0 ? host::sm_count(dev) : 148;
  AACLIP_CUDA_CHECK(cudaMemsetAsync(a.done, 0, ((size_t)B + 4) * sizeof(unsigned int), stream));
  const int grid = (int)std::min<long long>(sms, total);
  if (seg_is_bf16) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(head_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AACLIP_CUDA_CHECK(host::launch(head_stream_kernel<true>, dim3(grid), dim3(HS_THREADS), smem, stream, a));
  } else {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(head_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AACLIP_CUDA_CHECK(host::launch(head_stream_kernel<false>, dim3(grid), dim3(HS_THREADS), smem, stream, a));
  }
  return host::OK;
}
