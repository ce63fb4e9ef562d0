// Microbenchmark: per-SM throughput of MUFU.EX2 in f32, f16x2 and bf16x2 form, and of the FMA-pipe cubic exp2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, int iters, float seed) {
  float a[8];
  unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + 0.001f * (threadIdx.x + i); h[i] = 0x38003800u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*reinterpret_cast<unsigned short*>(&h[i])));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int per_instr) {
  float* d; cudaMalloc(&d, 4);
  const int iters = 4096, blocks = 148 * 8;
  kern<MODE><<<blocks, 256>>>(d, iters, 0.3f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<MODE><<<blocks, 256>>>(d, iters, 0.3f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double instr = double(blocks) * 256 * iters * 8;
  printf("%-28s %8.3f ms  %7.2f G thread-instr/s  %7.2f G elements/s  (%.1f elem/clk/SM at %d MHz nominal)\n", name, ms,
         instr / ms / 1e6, instr * per_instr / ms / 1e6, instr * per_instr / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  cudaFree(d);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("ex2.approx.f16", 1);
  return 0;
}
