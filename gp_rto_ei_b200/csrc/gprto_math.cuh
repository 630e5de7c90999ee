// Device math shared by the GP kernels (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "gprto_det_math.h"

namespace gprto {

// exp / log / rsqrt / erfc / acquisition live in gprto_det_math.h: IEEE operations and integer bit manipulation only, so that
// the ordered CPU twin (oracle/twin/) reproduces them bit for bit.  The library is compiled with -fmad=false: a*b+c is two
// roundings unless fma() is written, on the device exactly as on the host.
__device__ __forceinline__ double exp_fast(double a) { return exp_det(a); }

// numpy.argmin ordering on (value, index): a NaN beats any number, ties go to the lower index.
__device__ __forceinline__ bool better(double va, int64_t ia, double vb, int64_t ib) {
  const bool na = va != va, nb = vb != vb;
  if (na || nb) return na && (!nb || ia < ib);
  return va < vb || (va == vb && ia < ib);
}

__device__ __forceinline__ double acquisition(int acq, double mean, double var, double prior, double f_best) {
  return acquisition_det(acq, mean, var, prior, f_best);
}

__device__ __forceinline__ double rsqrt_fast(double x) { return rsqrt_det(x); }

// Counter-based trust-region sampler shared by K4's fused mode and gprto_sample_candidates: candidate c of GP b is a
// pure function of (seed, b, c).  shape 0: uniform in the box [-1,1]^D; shape 1: uniform in the unit ball (Gaussian
// direction x U^(1/D)); the direction arithmetic is FP32 (SFU log/sin/cos) - it only has to be deterministic, the
// resulting point is then used in full FP64 everywhere.  The caller maps u to centre + radius * u.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {      // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

template <int D>
__device__ __forceinline__ void tr_sample(uint64_t seed, int64_t b, int64_t c, int shape, double (&u)[D]) {
  const uint64_t key = mix64(seed ^ mix64(uint64_t(b) * 0xD1342543DE82EF95ULL + uint64_t(c)));
  if (shape == 0) {
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      const uint64_t r = mix64(key + uint64_t(dd + 1));
      u[dd] = double(r >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
  } else {
    float z[D];
    float n2 = 0.f;
#pragma unroll
    for (int dd = 0; dd < D; dd += 2) {
      const uint64_t r = mix64(key + uint64_t(dd + 1));
      const float u1 = (float(uint32_t(r >> 40)) + 0.5f) * (1.0f / 16777216.0f);      // (0,1)
      const float u2 = (float(uint32_t(r) >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * __logf(u1));
      float sn, cs;
      __sincosf(6.283185307179586f * u2, &sn, &cs);
      z[dd] = rad * cs;
      n2 = fmaf(z[dd], z[dd], n2);
      if (dd + 1 < D) { z[dd + 1] = rad * sn; n2 = fmaf(z[dd + 1], z[dd + 1], n2); }
    }
    const uint64_t r = mix64(key + uint64_t(D + 7));
    const float ur = (float(uint32_t(r >> 40)) + 0.5f) * (1.0f / 16777216.0f);
    const float scale = exp2f(__log2f(ur) * (1.0f / D)) * rsqrtf(n2);
#pragma unroll
    for (int dd = 0; dd < D; ++dd) u[dd] = double(z[dd] * scale);
  }
}

// Candidate c of GP b as a point: centre + radius * u for the box (shape 0) and ball (shape 1) samplers (rad[0] = the
// radius).  Shape 2 is the reference's start-point sampler for scaled inputs (ITR_GP_RTO.Ellipse_sampling,
// sub_uts/utilities_2.py:693-730): u uniform in the unit ball, stretched by the per-dimension semi-axes rad[0..D-1]
// (the reference's sqrt(0.99) * Delta * sqrt(range_k)), and - as the reference's TR_con test at :725-728 does - accepted
// only inside the trust-region ball of radius rad[D] (rad[D] <= 0: no ball test).  The reference scans 200 pre-drawn
// samples for the first accepted one; here attempt a of candidate c is the pure function tr_sample(seed, b, c + a*2^40),
// up to 64 attempts (the acceptance rate is the area ratio ball / ellipse, 11 % for the Williams-Otto scaling), after
// which the last sample is pulled radially into the ball.
template <int D>
__device__ __forceinline__ void tr_point(uint64_t seed, int64_t b, int64_t c, int shape, const double* __restrict__ centre,
                                         const double* __restrict__ rad, double (&x)[D]) {
  double u[D];
  if (shape != 2) {
    tr_sample<D>(seed, b, c, shape, u);
    const double r = rad[0];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) x[dd] = fma(r, u[dd], centre[dd]);
    return;
  }
  const double ball = rad[D];
  double off[D];
  double n2 = 0.0;
  for (int a = 0; a < 64; ++a) {
    tr_sample<D>(seed, b, c + (int64_t(a) << 40), 1, u);
    n2 = 0.0;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) { off[dd] = rad[dd] * u[dd]; n2 = fma(off[dd], off[dd], n2); }
    if (!(ball > 0.0) || n2 <= ball * ball) break;
  }
  if (ball > 0.0 && n2 > ball * ball) {
    const double sc = 0.999 * ball * rsqrt_det(n2);
#pragma unroll
    for (int dd = 0; dd < D; ++dd) off[dd] *= sc;
  }
#pragma unroll
  for (int dd = 0; dd < D; ++dd) x[dd] = centre[dd] + off[dd];
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

}  // namespace gprto
