// Microbenchmark: the read-modify-write pattern of the residual GEMM epilogue WITHOUT the GEMM.
// x fp32 [M, N] is updated in place (x += 1) and its bf16 copy written, by 148 persistent CTAs of 4 warps; a warp owns
// 32 rows x 256 columns of a 128 x 256 tile and walks them in 32-column chunks (4 KB TMA boxes, 128B swizzle) through a
// ring of RING chunks, exactly as epilogue_resid_ln (csrc/gemm_sm100.cuh) does.  Question: what HBM rate does this
// access pattern reach on its own, and how does it depend on the ring depth / on the chunk order?
//   MODE 0: chunk-major walk of a tile (the epilogue's order)      MODE 1: ring primed with a whole tile at once
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../aaclip_b200/csrc -o xrmw_bench xrmw_bench.cu
#include <stdarg.h>
#include <cstdio>
#include <vector>
#include <cuda_bf16.h>
#include "common.cuh"
#include "ptx.cuh"

template <int RING, int WITH_Y, int PROD>
__global__ void __launch_bounds__(160) rmw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                                                  const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + warp * (RING + 1) * 4096;
  uint8_t* Y = ring + RING * 4096;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * (RING + 1) * 4096) + warp * RING;
  uint8_t* opring = smem + 4 * (RING + 1) * 4096 + 1024;      // PROD: 4 stages x 32 KB operand ring
  uint64_t* pbar = reinterpret_cast<uint64_t*>(smem + 4 * (RING + 1) * 4096) + 4 * RING;
  if (warp < 4 && lane == 0) for (int s = 0; s < RING; ++s) ptx::mbar_init(&bars[s], 1);
  if (warp == 4 && lane == 0) for (int s = 0; s < 4; ++s) ptx::mbar_init(&pbar[s], 1);
  ptx::fence_barrier_init();
  __syncthreads();
  const int tiles_n = N / 256, tiles = ((M + 127) / 128) * tiles_n;
  if (warp == 4) {
    // operand-like traffic of a 128 x 256 x K tile: per k-block A [128 x 64] and B [128 x 64] bf16 boxes, waited and dropped
    if (PROD && lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int mb = t / tiles_n, nb = t - mb * tiles_n;
        for (int kb = 0; kb < K / 64; ++kb, ++it) {
          const uint32_t s = it & 3u;
          if (it >= 4) ptx::mbar_wait(&pbar[s], ((it >> 2) - 1) & 1u);
          ptx::mbar_arrive_expect_tx(&pbar[s], 32768);
          ptx::tma_load_2d(opring + s * 32768, &tmA, &pbar[s], kb * 64, mb * 128);
          ptx::tma_load_2d(opring + s * 32768 + 16384, &tmB, &pbar[s], kb * 64, nb * 256);
        }
      }
      for (uint32_t d = (it > 4 ? it - 4 : 0); d < it; ++d) ptx::mbar_wait(&pbar[d & 3u], (d >> 2) & 1u);
    }
    return;
  }
  const uint32_t sw = lane & 7u;
  constexpr int NCH = 8;
  auto request = [&](uint32_t jj) {
    const int t = blockIdx.x + int(jj / NCH) * gridDim.x;
    if (t >= tiles) return;
    const int mb = t / tiles_n, nb = t - mb * tiles_n;
    const uint32_t slot = jj % RING;
    ptx::mbar_arrive_expect_tx(&bars[slot], 4096);
    ptx::tma_load_2d(ring + slot * 4096, &tmX, &bars[slot], nb * 256 + int(jj % NCH) * 32, mb * 128 + int(warp) * 32);
  };
  if (lane == 0) for (uint32_t jj = 0; jj + 1 < RING; ++jj) request(jj);
  uint32_t j = 0;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int mb = t / tiles_n, nb = t - mb * tiles_n;
    const int row0 = mb * 128 + int(warp) * 32;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c, ++j) {
      const int col = nb * 256 + c * 32;
      const uint32_t slot = j % RING;
      uint8_t* xrow = ring + slot * 4096 + lane * 128;
      uint8_t* yrow = Y + lane * 128;
      ptx::mbar_wait(&bars[slot], (j / RING) & 1u);
      float f[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        ptx::ld_shared_v4(xrow + ((uint32_t(q) ^ sw) << 4), f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        f[4 * q] += 1.f; f[4 * q + 1] += 1.f; f[4 * q + 2] += 1.f; f[4 * q + 3] += 1.f;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        ptx::st_shared_v4(xrow + ((uint32_t(q) ^ sw) << 4), __float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]),
                          __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3]));
      if (WITH_Y) {
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = ptx::pack_bf16x2(f[2 * i], f[2 * i + 1]);
        if ((c & 1) == 0) { if (lane == 0) ptx::bulk_wait_read<0>(); __syncwarp(); }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          ptx::st_shared_v4(yrow + ((uint32_t((c & 1) * 4 + q) ^ sw) << 4), h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (row0 < M) {
          ptx::tma_store_2d(&tmX, ring + slot * 4096, col, row0);
          if (WITH_Y && (c & 1)) ptx::tma_store_2d(&tmY, Y, col - 32, row0);
        }
        ptx::bulk_commit();
        if (j >= 1) ptx::bulk_wait_read<1>();
        request(j + RING - 1);
      }
    }
  }
  if (lane == 0) ptx::bulk_wait<0>();
}

template <int RING, int WITH_Y, int PROD = 0>
void run(float* x, __nv_bfloat16* y, int M, int N, const __nv_bfloat16* a = nullptr, const __nv_bfloat16* w = nullptr, int K = 1024) {
  CUtensorMap tmX, tmY, tmA, tmB;
  if (host::make_tmap_out(&tmX, x, M, N, N, false) || host::make_tmap_out(&tmY, y, M, N, N, true)) { printf("tmap failed: %s\n", host::last_error().c_str()); return; }
  memset(&tmA, 0, sizeof tmA); memset(&tmB, 0, sizeof tmB);
  if (PROD && (host::make_tmap_2d(&tmA, a, M, K, K, 128) || host::make_tmap_2d(&tmB, w, N, K, K, 128))) { printf("tmap failed: %s\n", host::last_error().c_str()); return; }
  const size_t smem = 4 * (RING + 1) * 4096 + 1024 + (PROD ? 4 * 32768 : 0) + 1024;
  auto k = rmw_kernel<RING, WITH_Y, PROD>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int grid : {148, 296}) {
    if (grid == 296 && smem > 110 * 1024) continue;
    for (int i = 0; i < 3; ++i) k<<<grid, 160, smem>>>(tmX, tmY, tmA, tmB, M, N, K);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k<<<grid, 160, smem>>>(tmX, tmY, tmA, tmB, M, N, K);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = double(M) * N * (8 + (WITH_Y ? 2 : 0));
    printf("ring %d  bf16 copy %d  operand stream K=%4d  grid %3d (%d warps/SM): %7.1f us  %6.0f GB/s (epilogue bytes only)   %s\n", RING, WITH_Y, PROD ? K : 0, grid, grid / 148 * 4, ms / reps * 1e3,
           bytes / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError())); fflush(stdout);
  }
}

int main() {
  const int M = 64 * 577, N = 1024;
  float* x; __nv_bfloat16* y;
  cudaMalloc(&x, size_t(M) * N * 4); cudaMalloc(&y, size_t(M) * N * 2);
  cudaMemset(x, 0, size_t(M) * N * 4);
  run<2, 1>(x, y, M, N); run<3, 1>(x, y, M, N); run<4, 1>(x, y, M, N); run<6, 1>(x, y, M, N); run<8, 1>(x, y, M, N);
  run<3, 0>(x, y, M, N); run<8, 0>(x, y, M, N);
  __nv_bfloat16 *a, *w;
  cudaMalloc(&a, size_t(M) * 4096 * 2); cudaMalloc(&w, size_t(N) * 4096 * 2);
  cudaMemset(a, 0, size_t(M) * 4096 * 2); cudaMemset(w, 0, size_t(N) * 4096 * 2);
  run<3, 1, 1>(x, y, M, N, a, w, 1024); run<4, 1, 1>(x, y, M, N, a, w, 1024); run<3, 1, 1>(x, y, M, N, a, w, 4096);
  // correctness of the walk: every element was incremented the same number of times
  std::vector<float> h(1024);
  cudaMemcpy(h.data(), x + size_t(M - 1) * N, 4096, cudaMemcpyDeviceToHost);
  printf("x[last row][0] = %.0f  x[last row][1023] = %.0f\n", h[0], h[1023]);
  return 0;
}
