// Launchers defined in separate translation units (k3_inst.cu, k4_inst_*.cu) so that the template instantiations
// compile in parallel (gp_rto_ei_b200/build.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gprto.h"

namespace gprto {
// One-time per-DEVICE set-up of a kernel (cudaFuncSetAttribute applies to the current device only; a process may hold
// handles on several devices).
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};
struct K4Params;
constexpr int K3R_MAXN = 128;
constexpr int K3R_MAXD = 8;
// K4 DMMA kernel, grouped by NB = ceil(n / 8): a: 1-4, b: 5-8, c: 9-12, d: 13-16.  GPRTO_EUNSUPPORTED if (nb, d) has no instance.
int gprto_k4_launch_a(int nb, const K4Params& p, cudaStream_t s);
int gprto_k4_launch_b(int nb, const K4Params& p, cudaStream_t s);
int gprto_k4_launch_c(int nb, const K4Params& p, cudaStream_t s);
int gprto_k4_launch_d(int nb, const K4Params& p, cudaStream_t s);
// K3 register-resident kernels (GPRTO_EUNSUPPORTED outside their size range)
int launch_nlml_la(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* nll,
                   int32_t* status, cudaStream_t s, const int32_t* gidx);
int launch_nlml_reg(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* nll,
                    int32_t* status, cudaStream_t s, const int32_t* gidx);
int launch_nlml_lb(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* nll,
                   int32_t* status, cudaStream_t s, const int32_t* gidx);
int launch_nlml_lc(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* nll,
                   int32_t* status, cudaStream_t s, const int32_t* gidx);
// K2 register-tile / DMMA factor kernel (32 < n <= 128; GPRTO_EUNSUPPORTED otherwise)
int launch_factor_la(int64_t B, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* invK,
                     double* alpha, double* logdet, int32_t* status, cudaStream_t s);
}  // namespace gprto
