// In which order does DMMA.8x8x4 (mma.sync.m8n8k4.f64) accumulate its four products?  Compares the tensor-core result bit for
// bit against candidate orderings of IEEE fma chains on random operands with wide exponent spread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ uint64_t mix(uint64_t z) { z += 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
__device__ double rnd(uint64_t k) { uint64_t r = mix(k); double m = double(r >> 11) / 9007199254740992.0 * 2 - 1; int e = int(mix(r) % 41) - 20; return ldexp(m, e); }
__global__ void k(int trials, unsigned long long* hits) {
  __shared__ double A[8][4], B[4][8], C[8][8];
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  unsigned long long h[6] = {0, 0, 0, 0, 0, 0};
  for (int tr = 0; tr < trials; ++tr) {
    const uint64_t base = (uint64_t(blockIdx.x) * trials + tr) * 4096;
    A[g][t] = rnd(base + lane);
    B[t][g] = rnd(base + 100 + lane);
    C[g][2 * t] = rnd(base + 200 + lane); C[g][2 * t + 1] = rnd(base + 300 + lane);
    __syncwarp();
    double c0 = C[g][2 * t], c1 = C[g][2 * t + 1];
    dmma884(c0, c1, A[g][t], B[t][g]);
    for (int e = 0; e < 2; ++e) {
      const int col = 2 * t + e;
      const double got = e ? c1 : c0, c = C[g][col];
      const double p0a = A[g][0], p1a = A[g][1], p2a = A[g][2], p3a = A[g][3];
      const double b0 = B[0][col], b1 = B[1][col], b2 = B[2][col], b3 = B[3][col];
      const double fwd = fma(p3a, b3, fma(p2a, b2, fma(p1a, b1, fma(p0a, b0, c))));
      const double rev = fma(p0a, b0, fma(p1a, b1, fma(p2a, b2, fma(p3a, b3, c))));
      const double pair = (fma(p0a, b0, p1a * b1) + fma(p2a, b2, p3a * b3)) + c;
      const double dotfirst = fma(p3a, b3, fma(p2a, b2, fma(p1a, b1, p0a * b0))) + c;
      const double mulsum = ((p0a * b0 + p1a * b1) + (p2a * b2 + p3a * b3)) + c;
      h[0] += (got == fwd); h[1] += (got == rev); h[2] += (got == pair); h[3] += (got == dotfirst); h[4] += (got == mulsum); h[5] += 1;
    }
    __syncwarp();
  }
  for (int i = 0; i < 6; ++i) atomicAdd(&hits[i], h[i]);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 48); cudaMemset(d, 0, 48);
  k<<<64, 32>>>(2000, d);
  cudaDeviceSynchronize();
  unsigned long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("{\"trials\": %llu, \"fma_chain_k0_to_k3\": %llu, \"fma_chain_k3_to_k0\": %llu, \"pairwise\": %llu, \"dot_then_add_c\": %llu, \"products_rounded_then_summed\": %llu, \"status\": \"%s\"}\n",
         h[5], h[0], h[1], h[2], h[3], h[4], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
