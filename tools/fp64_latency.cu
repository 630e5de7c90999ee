// FP64 latency micro-benchmarks on B200 (sm_100a): what the serial pivot chain of the batched Cholesky kernels (K2/K3) pays
// per dependent instruction, alone and while the other warps of the SM stream DMMA.8x8x4 through the shared FP64 pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu && tools/fp64_latency
// Prints one JSON line (cycles per dependent instruction).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int ITERS = 4096;

// mode 0: dependent DFMA chain; 1: dependent DMUL->DFMA pairs with a shuffle in between (the diagonal-tile column step);
// 2: dependent DMMA chain (same accumulators); 3: LDS -> DFMA chain (panel row solve)
// Warp `probe` of every CTA runs the chain; the other warps run `bg` background work: 0 nothing (they exit), 1 independent DMMA
// streams (4 accumulator pairs), 2 independent DFMA streams.
__global__ void lat_kernel(int mode, int bg, int probe, double seed, long long* cycles, double* sink, int idle_mask = 0) {
  __shared__ double sm[64];
  __shared__ volatile int stop;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) sm[threadIdx.x] = 1.0 + 1e-9 * threadIdx.x;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (w == probe) {
    double x = seed, y = 1.0 + 1e-12, c0 = 0.0, c1 = 0.0;
    const long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 16
      for (int i = 0; i < ITERS; ++i) x = fma(x, y, 1e-30);
    } else if (mode == 1) {
#pragma unroll 8
      for (int i = 0; i < ITERS; ++i) {
        const double p = __shfl_sync(0xffffffffu, x, (lane + 4) & 31);
        const double dt = p * y;
        x = fma(-dt, 1e-30, dt * x);
      }
    } else if (mode == 2) {
#pragma unroll 16
      for (int i = 0; i < ITERS; ++i) dmma884(c0, c1, y, 1e-30);
      x = c0 + c1;
    } else {
#pragma unroll 8
      for (int i = 0; i < ITERS; ++i) x = fma(-sm[(i + lane) & 63], x, x * 1e-30);
    }
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = x;
    stop = 1;
  } else if (bg == 0 || ((idle_mask >> w) & 1)) {
    return;
  } else {
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    double p = seed, q = seed + 1, r = seed + 2, s = seed + 3;
    while (!stop) {
      if (bg == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { dmma884(a0, a1, seed, 1e-30); dmma884(b0, b1, seed, 1e-30); dmma884(e0, e1, seed, 1e-30); dmma884(f0, f1, seed, 1e-30); }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { p = fma(p, 1.0000001, 1e-30); q = fma(q, 1.0000001, 1e-30); r = fma(r, 1.0000001, 1e-30); s = fma(s, 1.0000001, 1e-30); }
      }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + b0 + b1 + e0 + e1 + f0 + f1 + p + q + r + s;
  }
}

int main() {
  long long* cyc; double* sink;
  cudaMalloc(&cyc, 8); cudaMalloc(&sink, sizeof(double) * 1024 * 512);
  const char* names[4] = {"dfma_chain", "shfl_dmul_dfma_step", "dmma_chain", "lds_dfma_chain"};
  const int per_iter[4] = {1, 1, 1, 1};
  printf("{");
  bool first = true;
  for (int mode = 0; mode < 4; ++mode)
    for (int cfg = 0; cfg < 9; ++cfg) {
      // cfg 0: probe alone (1 CTA of 1 warp per SM); 1: 8 warps, 7 DMMA background, probe = warp 7; 2: same, probe = warp 0;
      // 3: two such CTAs per SM (grid 296); 4: 7 DFMA background warps
      if (cfg == 3) continue;
      const int bg = cfg == 0 ? 0 : (cfg == 4 ? 2 : 1);
      const int probe = cfg == 2 ? 0 : (cfg == 0 ? 0 : 7);
      const int threads = cfg == 0 ? 32 : 256;
      const int grid = cfg >= 7 ? 296 : 148;      // cfg 7/8: TWO CTAs per SM (as the Cholesky kernels run)
      // cfg 5: the warp sharing the probe's scheduler (warp 3, same id mod 4) stays idle; cfg 6: warps 3 AND 2 idle
      const int idle_mask = (cfg == 5 || cfg == 8) ? (1 << 3) : (cfg == 6 ? (1 << 3) | (1 << 2) : 0);
      for (int rep = 0; rep < 2; ++rep) lat_kernel<<<grid, threads>>>(mode, bg, probe, 1.0, cyc, sink, idle_mask);
      cudaDeviceSynchronize();
      long long h = 0;
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const char* cn[9] = {"alone", "with_7_dmma_warps_probe_last", "with_7_dmma_warps_probe_first", "unused", "with_7_dfma_warps", "with_6_dmma_warps_same_scheduler_idle", "with_5_dmma_warps_two_idle", "two_ctas_per_sm_7_dmma_warps_each", "two_ctas_per_sm_same_scheduler_idle"};
      printf("%s\"%s_%s\": %.1f", first ? "" : ", ", names[mode], cn[cfg], double(h) / ITERS / per_iter[mode]);
      first = false;
    }
  cudaError_t e = cudaGetLastError();
  printf(", \"status\": \"%s\", \"unit\": \"cycles per dependent step\"}\n", cudaGetErrorString(e));
  return 0;
}
