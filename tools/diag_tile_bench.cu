// Cycles of ONE call of the diagonal-tile elimination (k3_diag_tile_lean) and of one panel-row solve pass, alone on an SM.
#include <cstdio>
#include "../gp_rto_ei_b200/csrc/k3_nlml_la.cuh"
using namespace gprto;
__global__ void bench(long long* out, double* sink, int reps, int nwarps_active) {
  __shared__ double tile[4][8 * K3R_LDP];
  __shared__ double rinv[4][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w >= nwarps_active) return;
  double pm = 1.0; int pe = 0, info = 0;
  long long total = 0;
  for (int r = 0; r < reps; ++r) {
    // SPD 8x8 tile: 2 I + 0.1 * ones-ish
    const int rr = lane >> 2, q = lane & 3;
    tile[w][rr * K3R_LDP + 2 * q] = (rr == 2 * q ? 2.0 : 0.1) + 1e-3 * r;
    tile[w][rr * K3R_LDP + 2 * q + 1] = (rr == 2 * q + 1 ? 2.0 : 0.1) + 1e-3 * r;
    __syncwarp();
    const long long t0 = clock64();
    k3_diag_tile_lean(tile[w], K3R_LDP, rinv[w], lane, 0, pm, pe, info);
    __syncwarp();
    const long long t1 = clock64();
    total += t1 - t0;
  }
  if (lane == 0 && blockIdx.x == 0 && w == 0) out[0] = total / reps;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = pm + pe + info + tile[w][lane];
}
int main() {
  long long* d; double* s; cudaMalloc(&d, 8); cudaMalloc(&s, 8 * 148 * 128);
  for (int nw = 1; nw <= 4; nw *= 2) {
    bench<<<148, 128>>>(d, s, 200, nw);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("k3_diag_tile_lean: %lld cycles per call with %d warps (one per scheduler) active per SM\n", h, nw);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
