// Where do the warps of two co-resident 256-thread CTAs live?  Prints (%smid, %warpid) per logical warp for the CTAs of SM 0.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out, int regs_pad) {
  __shared__ double big[4096];
  unsigned smid, warpid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
  big[threadIdx.x] = threadIdx.x;
  __syncthreads();
  // keep both CTAs of an SM resident for a while
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
  if ((threadIdx.x & 31) == 0) {
    out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 2] = smid;
    out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + 1] = warpid;
  }
  if (regs_pad == 12345) out[0] = (int)big[5];
}
int main() {
  int* d; cudaMalloc(&d, 296 * 8 * 2 * 4);
  k<<<296, 256>>>(d, 0);
  cudaDeviceSynchronize();
  static int h[296 * 8 * 2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int sm = 0; sm < 3; ++sm)
    for (int b = 0; b < 296; ++b)
      if (h[b * 16] == sm) {
        printf("sm %d block %d warpids:", sm, b);
        for (int w = 0; w < 8; ++w) printf(" %d", h[(b * 8 + w) * 2 + 1]);
        printf("\n");
      }
  return 0;
}
