// FP64 peak micro-benchmarks for the roofline denominator (B200, sm_100a).
// Measures, with CUDA events on the launching stream:
//   dfma        : dependent-chain-free DFMA throughput (vector FP64 pipe)
//   dmma884     : mma.sync.aligned.m8n8k4.f64 throughput (FP64 tensor path)
//   dmma16816   : mma.sync.aligned.m16n8k16.f64 throughput
//   mixed       : DFMA and DMMA issued from the same warps (do the pipes overlap?)
//   exp_libm    : libdevice exp() evaluations / s
//   lds_dfma    : broadcast LDS.128 + 2 DFMA (the naive quadratic-form inner loop)
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void k_dmma884(double* out, double a, double b) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, double a0, double b0) {
  double c[4][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = b0 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 DMMA.884 (8*256 FMA/warp) + NV DFMA per lane per iteration
template <int NV>
__global__ void k_mixed(double* out, double a, double b) {
  double c[16];
  double v[NV > 0 ? NV : 1];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma884(c[2 * i], c[2 * i + 1], a, b);
#pragma unroll
      for (int j = 0; j < NV / 8; ++j) v[i * (NV / 8) + j] = fma(v[i * (NV / 8) + j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, double x0) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -(threadIdx.x * 1e-3 + i) - x0;
  double s = 0;
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += exp(x[i]); x[i] = x[i] * 0.999; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// broadcast LDS.128 feeding 2*C DFMA (C candidates per thread share the broadcast matrix entry)
template <int C>
__global__ void k_lds_dfma(double* out, double a) {
  __shared__ double2 A[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) A[i] = make_double2(a + i, a - i);
  __syncthreads();
  double acc[C][2];
#pragma unroll
  for (int c = 0; c < C; ++c) { acc[c][0] = threadIdx.x + c; acc[c][1] = threadIdx.x - c; }
  double kv[C];
#pragma unroll
  for (int c = 0; c < C; ++c) kv[c] = 1.0 + 1e-9 * (threadIdx.x + c);
  for (int it = 0; it < ITERS / 64; ++it) {
#pragma unroll 16
    for (int j = 0; j < 1024; ++j) {
      double2 m = A[j];
#pragma unroll
      for (int c = 0; c < C; ++c) { acc[c][0] = fma(m.x, kv[c], acc[c][0]); acc[c][1] = fma(m.y, kv[c], acc[c][1]); }
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c][0] + acc[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  int blocks = sms * 8, threads = 256;
  double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  double nthreads = double(blocks) * threads, nwarps = nthreads / 32;

  float t_dfma = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  double dfma_tf = nthreads * ITERS * 16 * 2 / (t_dfma * 1e-3) / 1e12;
  float t_884 = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  double d884_tf = nwarps * ITERS * 8 * 256.0 * 2 / (t_884 * 1e-3) / 1e12;
  float t_16816 = time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  double d16816_tf = nwarps * (ITERS / 8) * 4 * (16.0 * 8 * 16) * 2 / (t_16816 * 1e-3) / 1e12;
  float t_m8 = time_ms([&] { k_mixed<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  double m8_tf = (nwarps * ITERS * 8 * 256.0 + nthreads * ITERS * 8) * 2 / (t_m8 * 1e-3) / 1e12;
  float t_m32 = time_ms([&] { k_mixed<32><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  double m32_tf = (nwarps * ITERS * 8 * 256.0 + nthreads * ITERS * 32) * 2 / (t_m32 * 1e-3) / 1e12;
  float t_m64 = time_ms([&] { k_mixed<64><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  double m64_tf = (nwarps * ITERS * 8 * 256.0 + nthreads * ITERS * 64) * 2 / (t_m64 * 1e-3) / 1e12;
  float t_exp = time_ms([&] { k_exp<<<blocks, threads>>>(out, 0.5); });
  double exp_g = nthreads * (ITERS / 8) * 8 / (t_exp * 1e-3) / 1e9;
  float t_l1 = time_ms([&] { k_lds_dfma<1><<<blocks, threads>>>(out, 1e-3); });
  double l1_tf = nthreads * (ITERS / 64) * 1024.0 * 2 * 1 * 2 / (t_l1 * 1e-3) / 1e12;
  float t_l2 = time_ms([&] { k_lds_dfma<2><<<blocks, threads>>>(out, 1e-3); });
  double l2_tf = nthreads * (ITERS / 64) * 1024.0 * 2 * 2 * 2 / (t_l2 * 1e-3) / 1e12;
  float t_l4 = time_ms([&] { k_lds_dfma<4><<<blocks, threads>>>(out, 1e-3); });
  double l4_tf = nthreads * (ITERS / 64) * 1024.0 * 2 * 4 * 2 / (t_l4 * 1e-3) / 1e12;

  // sustained DFMA: ~3 s back to back
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int nl = std::max(1, int(3000.0f / t_dfma));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < nl; ++i) k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double dfma_sus = nthreads * ITERS * 16 * 2 * nl / (ms * 1e-3) / 1e12;
  nl = std::max(1, int(3000.0f / t_884));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < nl; ++i) k_dmma884<<<blocks, threads>>>(out, 1.0000001, 1e-9);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double d884_sus = nwarps * ITERS * 8 * 256.0 * 2 * nl / (ms * 1e-3) / 1e12;

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"dfma_tflops\": %.2f, \"dfma_tflops_sustained\": %.2f, "
         "\"dmma884_tflops\": %.2f, \"dmma884_tflops_sustained\": %.2f, \"dmma16816_tflops\": %.2f, "
         "\"mixed_dmma_plus_8dfma_tflops\": %.2f, \"mixed_dmma_plus_32dfma_tflops\": %.2f, \"mixed_dmma_plus_64dfma_tflops\": %.2f, "
         "\"exp_libdevice_gevals\": %.2f, \"lds128_bcast_dfma_c1_tflops\": %.2f, \"lds128_bcast_dfma_c2_tflops\": %.2f, \"lds128_bcast_dfma_c4_tflops\": %.2f}\n",
         p.name, sms, p.clockRate, dfma_tf, dfma_sus, d884_tf, d884_sus, d16816_tf, m8_tf, m32_tf, m64_tf, exp_g, l1_tf, l2_tf, l4_tf);
  return 0;
}
