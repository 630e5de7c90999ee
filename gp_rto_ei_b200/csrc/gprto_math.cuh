// Device math shared by the GP kernels (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace gprto {

// exp(a) for finite a <= ~700.  Cody-Waite reduction a = n*ln2 + r, |r| <= ln2/2, degree-11 polynomial
// (Chebyshev-node fit; max relative error 1.7e-17 in exact arithmetic, measured < 1 ulp in FP64),
// scaling by exponent-field addition.  15 FP64-pipe instructions against ~21 for libdevice exp(): the
// cross-covariance needs n exponentials per candidate, which is ~1/3 of the FP64 work of K4.
// a < -700 returns 0 (the reference would return a value below 1e-304 there).
__device__ __forceinline__ double exp_fast(double a) {
  const double LOG2E = 1.4426950408889634074;
  const double LN2_HI = 6.93147180369123816490e-01;   // high part of ln2 (low 21 bits zero)
  const double LN2_LO = 1.90821492927058770002e-10;
  const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: round-to-nearest-integer shifter
  double t = fma(a, LOG2E, MAGIC);
  int n = __double2loint(t);
  double nf = t - MAGIC;
  double r = fma(nf, -LN2_HI, a);
  r = fma(nf, -LN2_LO, r);
  double p = 2.5110049204818658e-08;
  p = fma(p, r, 2.763265472252779e-07);
  p = fma(p, r, 2.755724088722987e-06);
  p = fma(p, r, 2.4801485441561313e-05);
  p = fma(p, r, 0.00019841269890076403);
  p = fma(p, r, 0.0013888888952352863);
  p = fma(p, r, 0.008333333333319589);
  p = fma(p, r, 0.04166666666648795);
  p = fma(p, r, 0.1666666666666668);
  p = fma(p, r, 0.5000000000000019);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int hi = __double2hiint(p) + (int)((unsigned)n << 20);
  double res = __hiloint2double(hi, __double2loint(p));
  return a < -700.0 ? 0.0 : res;
}

// numpy.argmin ordering on (value, index): a NaN beats any number, ties go to the lower index.
__device__ __forceinline__ bool better(double va, int64_t ia, double vb, int64_t ib) {
  const bool na = va != va, nb = vb != vb;
  if (na || nb) return na && (!nb || ia < ib);
  return va < vb || (va == vb && ia < ib);
}

// The three acquisition scores of ITR_GP_RTO.GP_obj_f (reference: sub_uts/utilities_2.py:485-504).
// mean/var are the de-normalised GP outputs, prior = obj_model(x).  EI keeps the reference's quirk
// Z := 0 when sigma == 0 and scipy's normal pdf/cdf definitions (pdf = exp(-Z^2/2)/sqrt(2 pi),
// cdf = ndtr(Z) = 0.5*erfc(-Z/sqrt 2)).
__device__ __forceinline__ double acquisition(int acq, double mean, double var, double prior, double f_best) {
  const double mu = prior + mean;
  if (acq == 1) return mu;
  const double sigma = sqrt(var);
  if (acq == 2) return mu - 3.0 * sigma;
  const double delta = f_best - mu;
  const double z = (sigma == 0.0) ? 0.0 : delta / sigma;
  const double pdf = exp(-0.5 * z * z) * 0.39894228040143267794;
  const double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
  return -(sigma * pdf + delta * cdf);
}

// 1/sqrt(x) for normal positive x: hardware seed (rsqrt.approx.f64, ~2^-22) + two Newton steps (6 dependent FP64
// instructions, relative error ~1e-16).  Non-positive or NaN x yields NaN/inf, which the callers flag beforehand.
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  double u = h * y;
  double t = fma(-u, y, 0.5);
  y = fma(y, t, y);
  u = h * y;
  t = fma(-u, y, 0.5);
  y = fma(y, t, y);
  return y;
}

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
    const double sc = 0.999 * ball * rsqrt(n2);
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
