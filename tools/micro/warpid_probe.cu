// Which hardware warp slot (%warpid; scheduler = slot % 4) does warp w of a 6-warp CTA get when 4 CTAs share an SM?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(192, 4) probe(int* out) {
  extern __shared__ char pad[];
  unsigned wid, smid;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if ((threadIdx.x & 31) == 0) {
    int w = threadIdx.x >> 5;
    out[(blockIdx.x * 6 + w) * 2] = smid;
    out[(blockIdx.x * 6 + w) * 2 + 1] = wid;
  }
  // keep the CTA resident long enough for all 4 per SM to coexist
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) { }
  if (pad[0] == 77) out[0] = 0;
}
int main() {
  const int blocks = 148 * 4;
  int* d; cudaMalloc(&d, blocks * 6 * 2 * sizeof(int));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 53504);
  probe<<<blocks, 192, 53504>>>(d);
  cudaDeviceSynchronize();
  static int h[148 * 4 * 6 * 2];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  for (int sm = 0; sm < 2; ++sm) {
    printf("SM %d:\n", sm);
    for (int b = 0; b < blocks; ++b) {
      if (h[b * 12] != sm) continue;
      printf("  block %3d: warp slots", b);
      for (int w = 0; w < 6; ++w) printf(" %2d(q%d)", h[(b * 6 + w) * 2 + 1], h[(b * 6 + w) * 2 + 1] % 4);
      printf("\n");
    }
  }
  return 0;
}
