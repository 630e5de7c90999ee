// Ordered CPU twin of the CUDA kernels.  TEST INFRASTRUCTURE ONLY (SURVEY.md par. 7.3-1a).
//
// The kernels' arithmetic is SPECIFIED: every floating-point operation is an IEEE-754 correctly rounded +, *, fma, /, sqrt
// in a fixed order, the transcendental functions are the library's own (gp_rto_ei_b200/csrc/gprto_det_math.h, built here for
// the host from the very same header), and the device code is compiled with -fmad=false.  This file restates, element by
// element and in the kernels' own order, what three kernels compute, so that the results are BIT-IDENTICAL to the GPU's:
//
//   twin_nlml       k3_nlml_tile_kernel  (k3_nlml_reg.cuh): register-tile blocked Cholesky NLML, n <= 32 is what the RTO
//                   loops of the two benchmark problems use (n = 4 .. 25); valid for every even tile count
//   twin_factor     factor_generic_kernel (k23_chol.cuh): Cholesky, explicit inverse, alpha; n <= 32 (above that the kernel's
//                   log-det uses an atomicAdd across warps and the register-tile kernel takes over)
//   twin_posterior  k4_dmma_kernel (k4_posterior.cuh): cross-covariance, folded quadratic form through DMMA.8x8x4,
//                   mean / variance, acquisition score
//
// What needed to be established about the hardware: DMMA.8x8x4 accumulates its four products as the fma chain
// c <- fma(a_k, b_k, c), k = 0..3 (tools/dmma_order.cu compares the tensor-core result bit for bit against candidate orderings).
// Warp-shuffle reductions are replayed as the same xor trees.  Who may use this: tests/ only (tests/twin_handle.py).
//
//   g++ -O2 -ffp-contract=off -shared -fPIC -o libgprto_twin.so gprto_twin.cpp
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../gp_rto_ei_b200/csrc/gprto_det_math.h"

using namespace gprto;

namespace {

// one DMMA.8x8x4 contribution to one accumulator element: four products, k = 0..3
inline double dmma4(double c, const double* a, const double* b) {
  c = fma(a[0], b[0], c);
  c = fma(a[1], b[1], c);
  c = fma(a[2], b[2], c);
  c = fma(a[3], b[3], c);
  return c;
}

// __shfl_xor butterfly over `width` lanes (offsets in the given order), value of lane 0
inline double xor_tree(double* v, int width, const int* offs, int noffs) {
  std::vector<double> t(width);
  for (int s = 0; s < noffs; ++s) {
    for (int l = 0; l < width; ++l) t[l] = v[l] + v[l ^ offs[s]];
    for (int l = 0; l < width; ++l) v[l] = t[l];
  }
  return v[0];
}

inline double k_entry(const double* xi, const double* xj, const double* nhw, int d, double c0) {
  double a = c0;
  for (int dd = 0; dd < d; ++dd) {
    const double df = xi[dd] - xj[dd];
    a = fma(df * df, nhw[dd], a);
  }
  return exp_det(a);
}

}  // namespace

extern "C" {

// k3_nlml_tile_kernel<NT>: nll = y'(K + (sn2 + jitter) I)^-1 y + log det, status = dpotrf code
int twin_nlml(int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* nll, int* status) {
  const int NT = (n + 15) / 16 * 2, NP = 8 * NT;
  std::vector<double> Xs(size_t(NP) * d, 0.0), ys(NP, 0.0), nhw(d);
  for (int e = 0; e < n * d; ++e) Xs[e] = Xn[e];
  for (int i = 0; i < n; ++i) ys[i] = Yn[i];
  for (int dd = 0; dd < d; ++dd) nhw[dd] = -0.5 * exp_det(-2.0 * theta[dd]);
  const double c0 = 2.0 * theta[d];
  const double diag_add = exp_det(2.0 * theta[d + 1]) + jitter;
  // full NP x NP storage; only tiles (I, J) with J <= I are ever touched (diagonal tiles hold both halves)
  std::vector<double> A(size_t(NP) * NP, 0.0);
  auto at = [&](int i, int j) -> double& { return A[size_t(i) * NP + j]; };
  for (int I = 0; I < NT; ++I)
    for (int J = 0; J <= I; ++J)
      for (int g = 0; g < 8; ++g)
        for (int c = 0; c < 8; ++c) {
          const int i = 8 * I + g, j = 8 * J + c;
          const double v = k_entry(&Xs[size_t(i) * d], &Xs[size_t(j) * d], nhw.data(), d, c0) + (i == j ? diag_add : 0.0);
          at(i, j) = (i < n && j < n) ? v : (i == j ? 1.0 : 0.0);
        }
  double logdet = 0.0, quad = 0.0;
  int info = 0;
  for (int kb = 0; kb < NT; ++kb) {
    // (2) diagonal tile: scaled, division- and sqrt-free elimination; the right-hand side block rides along
    double a[8][8], yv[8], myS[8], mydp[8];
    for (int r = 0; r < 8; ++r) {
      for (int k = 0; k < 8; ++k) a[r][k] = at(8 * kb + r, 8 * kb + k);
      yv[r] = ys[8 * kb + r];
    }
    double S = 1.0;
    for (int c = 0; c < 8; ++c) {
      const double dpiv = a[c][c];
      if (!(dpiv > 0.0) && info == 0) info = 8 * kb + c + 1;
      myS[c] = S;
      mydp[c] = dpiv;
      const int e = ((det_hi(dpiv) >> 20) & 0x7ff) - 1023;
      const double f = det_make((1023 - e) << 20, 0);
      const double dt = dpiv * f;
      double colc[8];
      for (int r = 0; r < 8; ++r) colc[r] = a[r][c];
      const double yc = yv[c];
      for (int r = 0; r < 8; ++r) {
        const double ac = colc[r] * f;
        for (int k = c + 1; k < 8; ++k) a[r][k] = fma(-ac, colc[k], dt * a[r][k]);
      }
      const double ycf = yc * f;
      for (int k8 = c + 1; k8 < 8; ++k8) yv[k8] = fma(-ycf, colc[k8], dt * yv[k8]);
      S *= dt;
    }
    double tm[8], lg[8], zq[8], z[8], rinv[8];
    for (int c = 0; c < 8; ++c) tm[c] = rsqrt_det(mydp[c] * myS[c]);
    for (int r = 0; r < 8; ++r)
      for (int k = 0; k < 8; ++k) at(8 * kb + r, 8 * kb + k) = a[r][k] * tm[k];
    for (int c = 0; c < 8; ++c) {
      z[c] = yv[c] * tm[c];
      zq[c] = z[c] * z[c];
      rinv[c] = tm[c] * myS[c];
      lg[c] = 2.0 * log_det(mydp[c] * tm[c]);
    }
    const int offs3[3] = {1, 2, 4};
    const double lgs = xor_tree(lg, 8, offs3, 3), zqs = xor_tree(zq, 8, offs3, 3);
    logdet += lgs;
    if (info != 0) break;
    // (3) panel rows L21 = A21 L11^-T
    for (int row = 8 * (kb + 1); row < NP; ++row) {
      double x[8];
      for (int c = 0; c < 8; ++c) x[c] = at(row, 8 * kb + c);
      for (int c = 0; c < 8; ++c) {
        x[c] *= rinv[c];
        for (int c2 = c + 1; c2 < 8; ++c2) x[c2] = fma(-at(8 * kb + c2, 8 * kb + c), x[c], x[c2]);
      }
      for (int c = 0; c < 8; ++c) at(row, 8 * kb + c) = x[c];
    }
    quad += zqs;
    // (4) trailing update, two DMMA.8x8x4 per tile (k = 0..3, then 4..7), B operand negated
    for (int J = kb + 1; J < NT; ++J) {
      for (int I = J; I < NT; ++I)
        for (int g = 0; g < 8; ++g)
          for (int c = 0; c < 8; ++c) {
            double bn[8];
            for (int k = 0; k < 8; ++k) bn[k] = -at(8 * J + c, 8 * kb + k);
            const double* ar = &at(8 * I + g, 8 * kb);
            double v = at(8 * I + g, 8 * J + c);
            v = dmma4(v, ar, bn);
            v = dmma4(v, ar + 4, bn + 4);
            at(8 * I + g, 8 * J + c) = v;
          }
      // the warp holding diagonal tile (J, J) also applies y'_J -= L(J, kb) z_kb
      for (int g = 0; g < 8; ++g) {
        double part[4];
        for (int t = 0; t < 4; ++t) {
          const double b0 = -at(8 * J + g, 8 * kb + t), b1 = -at(8 * J + g, 8 * kb + 4 + t);
          part[t] = fma(b0, z[t], b1 * z[4 + t]);
        }
        const int offs2[2] = {1, 2};
        ys[8 * J + g] += xor_tree(part, 4, offs2, 2);
      }
    }
  }
  *status = info;
  *nll = info == 0 ? quad + logdet : NAN;
  return 0;
}

// factor_generic_kernel: invK[n,n], alpha[n], logdet, status (n <= 32 for a deterministic log det)
int twin_factor(int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* invK, double* alpha,
                double* logdet, int* status) {
  const int ld = n | 1;
  std::vector<double> A(size_t(n) * ld, 0.0), nhw(d), rinv(n), z(n);
  for (int dd = 0; dd < d; ++dd) nhw[dd] = -0.5 * exp_det(-2.0 * theta[dd]);
  const double c0 = 2.0 * theta[d];
  const double diag_add = exp_det(2.0 * theta[d + 1]) + jitter;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = k_entry(&Xn[size_t(i) * d], &Xn[size_t(j) * d], nhw.data(), d, c0);
      if (i == j) v += diag_add;
      A[i * ld + j] = v;
      A[j * ld + i] = v;
    }
  int info = 0;
  for (int k = 0; k < n; ++k) {
    const double piv = A[k * ld + k];
    if (!(piv > 0.0)) { info = k + 1; break; }
    const double r = rsqrt_det(piv);
    rinv[k] = r;
    const double ip = r * r;
    for (int i = k + 1; i < n; ++i) {
      const double lik = A[i * ld + k] * ip;
      for (int j = k + 1; j <= i; ++j) A[i * ld + j] = fma(-lik, A[j * ld + k], A[i * ld + j]);
    }
  }
  if (info != 0) {
    for (int e = 0; e < n * n; ++e) invK[e] = NAN;
    for (int i = 0; i < n; ++i) alpha[i] = NAN;
    *status = info;
    if (logdet) *logdet = NAN;
    return 0;
  }
  {  // log det: one term per lane of warp 0 (n <= 32), butterfly 16, 8, 4, 2, 1
    double lg[32];
    for (int l = 0; l < 32; ++l) lg[l] = l < n ? 0.0 + log_det(A[l * ld + l]) : 0.0;
    const int offs[5] = {16, 8, 4, 2, 1};
    const double s = xor_tree(lg, 32, offs, 5);
    if (logdet) *logdet = 0.0 + s;
  }
  for (int i = 0; i < n; ++i)
    for (int k = 0; k <= i; ++k) A[i * ld + k] *= (k < i ? rinv[k] : rinv[i]);
  // inverse of the lower factor into the upper triangle: U[c][i] = X[i][c]
  for (int c = 0; c < n; ++c) {
    const double xcc = rinv[c];
    for (int i = c + 1; i < n; ++i) {
      double a0 = A[i * ld + c] * xcc, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int k = c + 1;
      for (; k + 3 < i; k += 4) {
        a0 = fma(A[i * ld + k], A[c * ld + k], a0);
        a1 = fma(A[i * ld + k + 1], A[c * ld + k + 1], a1);
        a2 = fma(A[i * ld + k + 2], A[c * ld + k + 2], a2);
        a3 = fma(A[i * ld + k + 3], A[c * ld + k + 3], a3);
      }
      for (; k < i; ++k) a0 = fma(A[i * ld + k], A[c * ld + k], a0);
      A[c * ld + i] = -((a0 + a1) + (a2 + a3)) * rinv[i];
    }
  }
  for (int c = 0; c < n; ++c) A[c * ld + c] = rinv[c];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double a0 = 0.0, a1 = 0.0;
      int k = i;
      for (; k + 1 < n; k += 2) {
        a0 = fma(A[i * ld + k], A[j * ld + k], a0);
        a1 = fma(A[i * ld + k + 1], A[j * ld + k + 1], a1);
      }
      if (k < n) a0 = fma(A[i * ld + k], A[j * ld + k], a0);
      const double acc = a0 + a1;
      invK[size_t(i) * n + j] = acc;
      invK[size_t(j) * n + i] = acc;
    }
  for (int k = 0; k < n; ++k) {
    double acc = 0.0;
    for (int i = 0; i <= k; ++i) acc = fma(A[i * ld + k], Yn[i], acc);
    z[k] = acc;
  }
  for (int i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int k = i; k < n; ++k) acc = fma(A[i * ld + k], z[k], acc);
    alpha[i] = acc;
  }
  *status = 0;
  return 0;
}

// k4_dmma_kernel<NB, D>: mean / var / score of m candidates of one GP (any of the outputs may be NULL)
int twin_posterior(int n, int d, int m, const double* Xn, const double* invK, const double* alpha, const double* theta,
                   const double* x_mean, const double* x_std, double y_mean, double y_std, const double* cand, const double* prior,
                   double f_best, int acq, double* mean, double* var, double* score) {
  const int NB = (n + 7) / 8, NP = 8 * NB, NS = 2 * NB;
  std::vector<double> M(size_t(NP) * NP, 0.0), Xs(size_t(NP) * d, 0.0), As(NP, 0.0), a_d(d), nhw(d);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) M[size_t(i) * NP + j] = j < i ? invK[size_t(i) * n + j] + invK[size_t(j) * n + i] : invK[size_t(i) * n + i];
  for (int e = 0; e < n * d; ++e) Xs[e] = Xn[e];
  for (int i = 0; i < n; ++i) As[i] = alpha[i];
  for (int dd = 0; dd < d; ++dd) {
    a_d[dd] = 1.0 / x_std[dd];
    nhw[dd] = -0.5 * exp_det(-2.0 * theta[dd]);
  }
  const double c0 = 2.0 * theta[d], sf2 = exp_det(2.0 * theta[d]);
  auto pi = [](int s, int t) { return 8 * (s >> 1) + 2 * t + (s & 1); };
  std::vector<double> x(d), kp(NP), Mk(NP);
  for (int c = 0; c < m; ++c) {
    for (int dd = 0; dd < d; ++dd) x[dd] = (cand[size_t(c) * d + dd] - x_mean[dd]) * a_d[dd];
    for (int j = 0; j < NP; ++j) kp[j] = k_entry(x.data(), &Xs[size_t(j) * d], nhw.data(), d, c0);
    // (M k)_i for row i of block bb = i / 8: DMMA k-steps s = 0 .. 2 bb + 1, four data points pi(s, 0..3) each
    for (int i = 0; i < NP; ++i) {
      const int bb = i >> 3;
      double acc = 0.0;
      for (int s = 0; s <= 2 * bb + 1 && s < NS; ++s) {
        double a4[4], b4[4];
        for (int t = 0; t < 4; ++t) { a4[t] = kp[pi(s, t)]; b4[t] = M[size_t(i) * NP + pi(s, t)]; }
        acc = dmma4(acc, a4, b4);
      }
      Mk[i] = acc;
    }
    double qp[4], mp[4];
    for (int t = 0; t < 4; ++t) {
      double q = 0.0, mm = 0.0;
      for (int bb = 0; bb < NB; ++bb) {
        const int i0 = 8 * bb + 2 * t;
        q = fma(Mk[i0], kp[i0], q);
        q = fma(Mk[i0 + 1], kp[i0 + 1], q);
        mm = fma(kp[i0], As[i0], mm);
        mm = fma(kp[i0 + 1], As[i0 + 1], mm);
      }
      qp[t] = q; mp[t] = mm;
    }
    const int offs2[2] = {1, 2};
    const double q = xor_tree(qp, 4, offs2, 2), mn = xor_tree(mp, 4, offs2, 2);
    const double mu = fma(mn, y_std, y_mean);
    const double vv = fmax(0.0, sf2 - q) * (y_std * y_std);
    if (mean) mean[c] = mu;
    if (var) var[c] = vv;
    if (score) score[c] = acquisition_det(acq, mu, vv, prior ? prior[c] : 0.0, f_best);
  }
  return 0;
}

// the deterministic scalar functions themselves (unit tests against mpmath / libm)
double twin_exp(double a) { return exp_det(a); }
double twin_log(double a) { return log_det(a); }
double twin_rsqrt(double a) { return rsqrt_det(a); }
double twin_erfc(double a) { return erfc_det(a); }

}  // extern "C"
