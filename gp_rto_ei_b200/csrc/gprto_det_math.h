// Deterministic FP64 math shared by the CUDA kernels (device) and their ordered CPU twin (host, oracle/twin/).
//
// SURVEY.md par. 7.3-1a: bit-identity of RTO trajectories against NumPy is not attainable by any independent implementation,
// but the GPU arithmetic can be made SPECIFIED - fixed summation orders, explicit fma, own transcendental functions - so that
// the same header compiled for the host reproduces every kernel result bit for bit.  Everything here is built from IEEE-754
// operations that are correctly rounded on both sides (+, *, fma, /, sqrt) plus integer bit manipulation: no libdevice
// function and no MUFU approximation instruction (their results cannot be reproduced on a CPU).
//
//   exp_det    Cody-Waite reduction a = n ln2 + r, |r| <= ln2/2, degree-11 polynomial (Chebyshev-node fit, < 1 ulp measured),
//              scaling by exponent-field addition.  15 FP64 instructions (libdevice exp: ~21).
//   log_det    x = 2^k m, m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1), log m = 2 atanh(s) by its Taylor series to s^25.
//   rsqrt_det  integer magic-constant seed (3.4 % error) + four Newton steps in fma form (relative error ~1e-16).
//   erfc_det   erfc(x) = exp(-x^2) p(t) / (1 + 2x), t = (x-4)/(x+4), p = degree-26 polynomial fitted here (mpmath, Chebyshev
//              nodes) to erfcx(x)(1+2x) on x in [0, inf): max relative error 2e-16; erfc(-x) = 2 - erfc(x).
//   acquisition_det   the three scores of ITR_GP_RTO.GP_obj_f (reference: sub_uts/utilities_2.py:485-504).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GPRTO_HD __host__ __device__ __forceinline__
#else
#define GPRTO_HD static inline
#endif

namespace gprto {

GPRTO_HD int det_hi(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
GPRTO_HD int det_lo(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)u;
#endif
}
GPRTO_HD double det_make(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
#endif
}

// exp(a).  a < -700 returns 0 (the true value is below 1e-304), a > 709 returns +inf, NaN propagates.
GPRTO_HD double exp_det(double a) {
  const double LOG2E = 1.4426950408889634074;
  const double LN2_HI = 6.93147180369123816490e-01;   // high part of ln2 (low 21 bits zero)
  const double LN2_LO = 1.90821492927058770002e-10;
  const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: round-to-nearest-integer shifter
  double t = fma(a, LOG2E, MAGIC);
  int n = det_lo(t);
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
  int hi = det_hi(p) + (int)((unsigned)n << 20);
  double res = det_make(hi, det_lo(p));
  if (a < -700.0) res = 0.0;
  if (a > 709.0) res = INFINITY;
  return res;
}

// log(x) for positive normal x (x <= 0 or NaN: NaN; subnormals are not expected on this path and give a coarse result).
GPRTO_HD double log_det(double x) {
  if (!(x > 0.0)) return NAN;
  const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  int hx = det_hi(x);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  // mantissa in [sqrt(1/2), sqrt(2)): above sqrt(2) (high word 0x6a09e) halve it and bump the exponent
  const int up = hx >= 0x6a09e ? 1 : 0;
  k += up;
  const double m = det_make(hx | ((1023 - up) << 20), det_lo(x));
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  double p = 2.0 / 25.0;
  p = fma(p, z, 2.0 / 23.0);
  p = fma(p, z, 2.0 / 21.0);
  p = fma(p, z, 2.0 / 19.0);
  p = fma(p, z, 2.0 / 17.0);
  p = fma(p, z, 2.0 / 15.0);
  p = fma(p, z, 2.0 / 13.0);
  p = fma(p, z, 2.0 / 11.0);
  p = fma(p, z, 2.0 / 9.0);
  p = fma(p, z, 2.0 / 7.0);
  p = fma(p, z, 2.0 / 5.0);
  p = fma(p, z, 2.0 / 3.0);
  const double dk = (double)k;
  // log m = 2s + s z p ; assemble from small to large
  const double tail = fma(s * z, p, dk * LN2_LO);
  return fma(dk, LN2_HI, fma(2.0, s, tail));
}

// 1/sqrt(x) for positive normal x.  Non-positive or NaN x: the callers flag them beforehand.
GPRTO_HD double rsqrt_det(double x) {
  uint64_t u;
#if defined(__CUDA_ARCH__)
  u = (uint64_t)__double_as_longlong(x);
#else
  memcpy(&u, &x, 8);
#endif
  u = 0x5FE6EB50C7B537A9ULL - (u >> 1);
  double y;
#if defined(__CUDA_ARCH__)
  y = __longlong_as_double((long long)u);
#else
  memcpy(&y, &u, 8);
#endif
  const double h = 0.5 * x;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int it = 0; it < 4; ++it) {
    const double v = h * y;
    const double t = fma(-v, y, 0.5);
    y = fma(y, t, y);
  }
  return y;
}

// erfc(x), all finite x (|x| > 27: 0 or 2).
GPRTO_HD double erfc_det(double x) {
  const double ax = fabs(x);
  const double t = (ax - 4.0) / (ax + 4.0);
  double p = 4.33574469688012796e-11;
  p = fma(p, t, -1.37212432109557996e-10);
  p = fma(p, t, -5.47796315403793008e-10);
  p = fma(p, t, 1.78874921890775359e-09);
  p = fma(p, t, 3.97990820993533376e-09);
  p = fma(p, t, -1.47051875523876298e-08);
  p = fma(p, t, -2.19171286847767334e-08);
  p = fma(p, t, 1.10090415607283036e-07);
  p = fma(p, t, 7.89343963476141997e-08);
  p = fma(p, t, -8.24912743496457224e-07);
  p = fma(p, t, 2.90906296837142686e-07);
  p = fma(p, t, 5.70920121222635183e-06);
  p = fma(p, t, -1.12207534527254248e-05);
  p = fma(p, t, -2.43989423090698060e-05);
  p = fma(p, t, 1.50621307463917777e-04);
  p = fma(p, t, -1.99256776457919693e-04);
  p = fma(p, t, -7.57773219535229294e-04);
  p = fma(p, t, 5.03196999281092534e-03);
  p = fma(p, t, -1.61977340659525300e-02);
  p = fma(p, t, 3.71675155359114107e-02);
  p = fma(p, t, -6.63303658116485007e-02);
  p = fma(p, t, 9.37328349984395126e-02);
  p = fma(p, t, -1.01039066036325037e-01);
  p = fma(p, t, 6.80970542546903562e-02);
  p = fma(p, t, 1.53796521026200415e-02);
  p = fma(p, t, -1.39621116840562498e-01);
  p = fma(p, t, 1.23299511862555256e+00);
  const double e = exp_det(-ax * ax);
  const double r = e * p / fma(2.0, ax, 1.0);
  return x < 0.0 ? 2.0 - r : r;
}

// The three acquisition scores of ITR_GP_RTO.GP_obj_f (reference: sub_uts/utilities_2.py:485-504).  mean/var are the
// de-normalised GP outputs, prior = obj_model(x).  EI keeps the reference's quirk Z := 0 when sigma == 0 and scipy's normal
// pdf/cdf definitions (pdf = exp(-Z^2/2)/sqrt(2 pi), cdf = ndtr(Z) = 0.5*erfc(-Z/sqrt 2)).
GPRTO_HD double acquisition_det(int acq, double mean, double var, double prior, double f_best) {
  const double mu = prior + mean;
  if (acq == 1) return mu;
  const double sigma = sqrt(var);
  if (acq == 2) return mu - 3.0 * sigma;
  const double delta = f_best - mu;
  const double z = (sigma == 0.0) ? 0.0 : delta / sigma;
  const double pdf = exp_det(-0.5 * z * z) * 0.39894228040143267794;
  const double cdf = 0.5 * erfc_det(-z * 0.70710678118654752440);
  return -(sigma * pdf + delta * cdf);
}

}  // namespace gprto
