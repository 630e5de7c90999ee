// K1/K2/K3 (generic path): SE-ARD Gram matrix, Cholesky, log-det, solves, explicit inverse.
// One CTA per matrix, matrix in shared memory (row stride n|1 to keep column walks conflict-free).
// Replaces GP_model.Cov_mat / negative_loglikelihood / the tail of determine_hyperparameters
// (reference: sub_uts/utilities_2.py:50-59, 92-114, 168-175).
#pragma once
#include "gprto_math.cuh"

namespace gprto {

constexpr int CH_THREADS = 256;
constexpr int CH_MAXD = 16;

__host__ __device__ inline int ch_ld(int n) { return n | 1; }
inline size_t ch_smem_bytes(int n, int d) {
  return sizeof(double) * (size_t(n) * ch_ld(n) + size_t(n) * d + 3 * size_t(n) + 2 * CH_MAXD + 8);
}

// K_ij = exp_det(2 theta_sf + sum_dd (x_i - x_j)^2 * nhw_dd),  nhw_dd = -1/(2 W_dd) = -0.5 exp_det(-2 theta_dd).
// The difference is formed BEFORE scaling so its rounding error stays relative to |x_i - x_j| (as in the
// reference's cdist); scaling the points first loses ~length-scale^-1 digits on near-duplicate points, which
// cond(K) then amplifies in the NLML.
__device__ __forceinline__ void ch_stage_inputs(const double* __restrict__ Xn, const double* __restrict__ theta,
                                                int n, int d, double* Xs, double* nhw) {
  for (int e = threadIdx.x; e < n * d; e += blockDim.x) Xs[e] = Xn[e];
  if (threadIdx.x < d) nhw[threadIdx.x] = -0.5 * exp_det(-2.0 * theta[threadIdx.x]);
}

__device__ __forceinline__ void ch_build_K(const double* Xs, const double* nhw, int n, int d, double c0,
                                           double diag_add, double* A, int ld) {
  // lower triangle incl. diagonal, mirrored into the upper triangle; warp w takes rows w, w + nwarps, ...
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = warp; i < n; i += nwarps)
    for (int j = lane; j <= i; j += 32) {
      double acc = c0;
      for (int dd = 0; dd < d; ++dd) {
        const double df = Xs[i * d + dd] - Xs[j * d + dd];
        acc = fma(df * df, nhw[dd], acc);
      }
      double v = exp_fast(acc);
      if (i == j) v += diag_add;
      A[i * ld + j] = v;
      A[j * ld + i] = v;
    }
}

// In-place lower Cholesky of A (n x n, stride ld) by the whole CTA, ONE barrier per column: the column is not scaled
// while the factorisation runs - the rank-1 update uses the unscaled column and 1/pivot
// (A_ij -= A_ik A_jk / piv), 1/sqrt(pivot) goes to rinv[k], and all columns are scaled in one parallel pass at the
// end.  The diagonal keeps the PIVOTS until then (log det = sum log pivots, in parallel).  On exit: strictly lower
// triangle = L, diagonal = L_kk, rinv[k] = 1/L_kk.  *s_info is the dpotrf code.
__device__ __forceinline__ void ch_cholesky(double* A, int n, int ld, double* rinv, int* s_info, double* s_logdet) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) { *s_info = 0; *s_logdet = 0.0; }
  __syncthreads();
  const int tr = tid & 15, tc = tid >> 4;
  for (int k = 0; k < n; ++k) {
    const double piv = A[k * ld + k];               // final since the previous column's barrier
    if (!(piv > 0.0)) {                             // uniform: every thread reads the same value
      if (tid == 0) *s_info = k + 1;
      __syncthreads();
      return;
    }
    const double r = rsqrt_fast(piv);
    if (tid == 0) rinv[k] = r;
    const double ip = r * r;                        // 1 / pivot
    // trailing update of the lower triangle: 16 x 16 thread grid, rows/cols strided by 16
    for (int i = k + 1 + tr; i < n; i += 16) {
      const double lik = A[i * ld + k] * ip;
      for (int j = k + 1 + tc; j <= i; j += 16) A[i * ld + j] = fma(-lik, A[j * ld + k], A[i * ld + j]);
    }
    __syncthreads();
  }
  double lg = 0.0;
  for (int k = tid; k < n; k += nt) lg += log_det(A[k * ld + k]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, off);
  if ((tid & 31) == 0) atomicAdd(s_logdet, lg);
  __syncthreads();                                  // rinv complete, pivots read
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, k = e - i * n;
    if (k < i) A[i * ld + k] *= rinv[k];
    else if (k == i) A[i * ld + i] *= rinv[i];      // pivot -> L_ii
  }
  __syncthreads();
}

// K3 generic: one CTA per (GP, start).
__global__ void __launch_bounds__(CH_THREADS) nlml_generic_kernel(int S, int n, int d, const double* __restrict__ Xn,
                                                                  const double* __restrict__ Yn,
                                                                  const double* __restrict__ theta_all, double jitter,
                                                                  double* __restrict__ nll, int32_t* __restrict__ status,
                                                                  const int32_t* __restrict__ gidx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = ch_ld(n);
  double* A = reinterpret_cast<double*>(smem_raw);
  double* Xs = A + size_t(n) * ld;
  double* z = Xs + n * d;
  double* nhw = z + 2 * n;
  double* rinv = z + n;                      // second half of the 2n scratch
  __shared__ int s_info;
  __shared__ double s_logdet;
  const int64_t bs = blockIdx.x;
  const int64_t b = gidx ? gidx[bs] : bs / S;
  const double* theta = theta_all + bs * (d + 2);
  ch_stage_inputs(Xn + b * int64_t(n) * d, theta, n, d, Xs, nhw);
  for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] = Yn[b * n + i];
  __syncthreads();
  const double sn2 = exp_det(2.0 * theta[d + 1]);
  ch_build_K(Xs, nhw, n, d, 2.0 * theta[d], sn2 + jitter, A, ld);
  __syncthreads();
  ch_cholesky(A, n, ld, rinv, &s_info, &s_logdet);
  __syncthreads();
  if (s_info != 0) {
    if (threadIdx.x == 0) { nll[bs] = CUDART_NAN; status[bs] = s_info; }
    return;
  }
  // forward substitution L z = y by warp 0 (column-oriented), then y'K^-1 y = |z|^2
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int k = 0; k < n; ++k) {
      const double zk = z[k] / A[k * ld + k];
      __syncwarp();
      if (lane == 0) z[k] = zk;
      for (int i = k + 1 + lane; i < n; i += 32) z[i] = fma(-A[i * ld + k], zk, z[i]);
      __syncwarp();
    }
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc = fma(z[i], z[i], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) { nll[bs] = acc + s_logdet; status[bs] = 0; }
  }
}

// K1: Gram matrix to global memory (API parity with Cov_mat; the fused kernels never materialise K).  The only
// HBM-bound kernel of the path: n^2 x 8 B written per matrix.  Each entry of the lower triangle is evaluated once
// (half the exponentials), mirrored through a shared-memory tile, and the matrix leaves in fully coalesced 16-byte
// stores.  kind 0: SE-ARD (RBF branch, utilities_2.py:57-59); 1: Matern-3/2 (:61-63); 2: Matern-5/2 (:64-66), with
// dist = sum (x - x')^2 / W and r = sqrt(dist) as the reference forms them from cdist.
__device__ __forceinline__ double cov_entry(const double* Xs, const double* nhw, int d, int i, int j, double c0, double sf2,
                                            int kind) {
  double acc = 0.0;                           // -dist / 2
  for (int dd = 0; dd < d; ++dd) {
    const double df = Xs[i * d + dd] - Xs[j * d + dd];
    acc = fma(df * df, nhw[dd], acc);
  }
  if (kind == 0) return exp_fast(c0 + acc);
  const double dist = -2.0 * acc, r = sqrt(dist);
  if (kind == 1) return sf2 * (1.0 + 1.7320508075688772 * r) * exp_det(-1.7320508075688772 * r);
  return sf2 * (1.0 + 2.23606797749979 * r + (5.0 / 3.0) * dist) * exp_det(-2.23606797749979 * r);
}

inline size_t cov_smem_bytes(int n, int d, bool tiled) {
  return sizeof(double) * (size_t(n) * d + CH_MAXD + (tiled ? size_t(n) * (n | 1) : 0));
}

template <bool TILED>
__global__ void __launch_bounds__(CH_THREADS) cov_kernel(int S, int n, int d, const double* __restrict__ Xn,
                                                         const double* __restrict__ theta_all, int add_noise,
                                                         double jitter, int kind, double* __restrict__ K) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Xs = reinterpret_cast<double*>(smem_raw);
  double* nhw = Xs + n * d;
  double* tile = nhw + CH_MAXD;               // [n][n] when TILED
  const int64_t bs = blockIdx.x;
  const int64_t b = bs / S;
  const double* theta = theta_all + bs * (d + 2);
  ch_stage_inputs(Xn + b * int64_t(n) * d, theta, n, d, Xs, nhw);
  __syncthreads();
  const double diag = (add_noise ? exp_det(2.0 * theta[d + 1]) : 0.0) + jitter;
  const double c0 = 2.0 * theta[d];
  const double sf2 = exp_det(c0);
  double* Kout = K + bs * int64_t(n) * n;
  // warp w evaluates rows w, w + nwarps, ...: lanes stride over the columns j <= i (no index decoding, row data reused)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (TILED) {
    const int ld = n | 1;                       // odd row stride: the mirrored (column) writes are bank-conflict free
    for (int i = warp; i < n; i += nwarps)
      for (int j = lane; j <= i; j += 32) {
        double v = cov_entry(Xs, nhw, d, i, j, c0, sf2, kind);
        if (i == j) v += diag;
        tile[i * ld + j] = v;
        tile[j * ld + i] = v;
      }
    __syncthreads();
    for (int i = warp; i < n; i += nwarps)      // coalesced: a warp writes consecutive addresses of one row
      for (int j = lane; j < n; j += 32) Kout[int64_t(i) * n + j] = tile[i * ld + j];
  } else {
    // larger n: no tile; the mirrored store goes to a different row per lane, but the whole matrix is L2-resident
    // (126 MB L2) until complete, so the partial sectors merge there and HBM still sees n^2 x 8 B per matrix
    for (int i = warp; i < n; i += nwarps)
      for (int j = lane; j <= i; j += 32) {
        double v = cov_entry(Xs, nhw, d, i, j, c0, sf2, kind);
        if (i == j) v += diag;
        Kout[int64_t(i) * n + j] = v;
        Kout[int64_t(j) * n + i] = v;
      }
  }
}

// K2 generic: Cholesky + explicit inverse + alpha, one CTA per GP.
__global__ void __launch_bounds__(CH_THREADS) factor_generic_kernel(int n, int d, const double* __restrict__ Xn,
                                                                    const double* __restrict__ Yn,
                                                                    const double* __restrict__ theta_all, double jitter,
                                                                    double* __restrict__ invK, double* __restrict__ alpha,
                                                                    double* __restrict__ logdet,
                                                                    int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = ch_ld(n);
  double* A = reinterpret_cast<double*>(smem_raw);
  double* Xs = A + size_t(n) * ld;
  double* y = Xs + n * d;
  double* z = y + n;
  double* rinv = z + n;
  double* nhw = rinv + n;
  __shared__ int s_info;
  __shared__ double s_logdet;
  const int64_t b = blockIdx.x;
  const double* theta = theta_all + b * (d + 2);
  ch_stage_inputs(Xn + b * int64_t(n) * d, theta, n, d, Xs, nhw);
  for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = Yn[b * n + i];
  __syncthreads();
  ch_build_K(Xs, nhw, n, d, 2.0 * theta[d], exp_det(2.0 * theta[d + 1]) + jitter, A, ld);
  __syncthreads();
  ch_cholesky(A, n, ld, rinv, &s_info, &s_logdet);
  __syncthreads();
  double* out = invK + b * int64_t(n) * n;
  if (s_info != 0) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) out[e] = CUDART_NAN;
    for (int i = threadIdx.x; i < n; i += blockDim.x) alpha[b * n + i] = CUDART_NAN;
    if (threadIdx.x == 0) { status[b] = s_info; if (logdet) logdet[b] = CUDART_NAN; }
    return;
  }
  // Inverse of the lower factor without a second matrix: the upper triangle is free after Cholesky, so
  // column c of X = L^-1 (forward substitution, one column per thread) is written to ROW c of the upper
  // triangle, U[c][i] = X[i][c] for i > c.  Threads only read L (lower + diagonal) and their own row of U.
  // pass 1: strictly-lower part of L^-1 into the strictly-upper triangle; the diagonal 1/L_cc into z-free scratch
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const double xcc = rinv[c];
    for (int i = c + 1; i < n; ++i) {
      // X_ic = -(L_ic X_cc + sum_{c<k<i} L_ik X_kc) / L_ii ; four independent partial sums hide the FMA latency
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
  __syncthreads();
  for (int c = threadIdx.x; c < n; c += blockDim.x) A[c * ld + c] = rinv[c];
  __syncthreads();
  // now U (upper incl. diagonal) holds X' where X = L^-1:  U[c][i] = X[i][c], i >= c.
  // invK[i][j] = sum_{k >= max(i,j)} X[k][i] X[k][j] = sum_k U[i][k] U[j][k]
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = warp; i < n; i += nwarps)
      for (int j = lane; j <= i; j += 32) {
        double a0 = 0.0, a1 = 0.0;
        int k = i;
        for (; k + 1 < n; k += 2) {
          a0 = fma(A[i * ld + k], A[j * ld + k], a0);
          a1 = fma(A[i * ld + k + 1], A[j * ld + k + 1], a1);
        }
        if (k < n) a0 = fma(A[i * ld + k], A[j * ld + k], a0);
        const double acc = a0 + a1;
        out[int64_t(i) * n + j] = acc;
        out[int64_t(j) * n + i] = acc;
      }
  }
  // alpha = X' (X y):  z = X y (z_k = sum_{i<=k} X[k][i] y_i = sum_i U[i][k] y_i), alpha_i = sum_{k>=i} U[i][k] z_k
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i <= k; ++i) acc = fma(A[i * ld + k], y[i], acc);
    z[k] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int k = i; k < n; ++k) acc = fma(A[i * ld + k], z[k], acc);
    alpha[b * n + i] = acc;
  }
  if (threadIdx.x == 0) { status[b] = 0; if (logdet) logdet[b] = s_logdet; }
}

// Normalisation (utilities_2.py:38-40): one CTA per GP.
__global__ void normalise_kernel(int n, int d, const double* __restrict__ X, const double* __restrict__ Y,
                                 double* __restrict__ x_mean, double* __restrict__ x_std, double* __restrict__ y_mean,
                                 double* __restrict__ y_std, double* __restrict__ Xn, double* __restrict__ Yn) {
  const int64_t b = blockIdx.x;
  __shared__ double s_mean[CH_MAXD + 1], s_std[CH_MAXD + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int col = warp; col <= d; col += nw) {       // column d is Y
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc += (col < d) ? X[(b * n + i) * d + col] : Y[b * n + i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    const double mu = acc / n;
    double v = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double df = ((col < d) ? X[(b * n + i) * d + col] : Y[b * n + i]) - mu;
      v = fma(df, df, v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) { s_mean[col] = mu; s_std[col] = sqrt(v / n); }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * d; e += blockDim.x) {
    const int dd = e % d;
    Xn[b * int64_t(n) * d + e] = (X[b * int64_t(n) * d + e] - s_mean[dd]) / s_std[dd];
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) Yn[b * n + i] = (Y[b * n + i] - s_mean[d]) / s_std[d];
  if (threadIdx.x < d) { x_mean[b * d + threadIdx.x] = s_mean[threadIdx.x]; x_std[b * d + threadIdx.x] = s_std[threadIdx.x]; }
  if (threadIdx.x == 0) { y_mean[b] = s_mean[d]; y_std[b] = s_std[d]; }
}

__global__ void pack_best_kernel(int64_t count, const double* __restrict__ score, const int64_t* __restrict__ idx,
                                 int64_t offset, double* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[2 * i] = score[i];
  reinterpret_cast<int64_t*>(out)[2 * i + 1] = idx[i] + offset;
}

__global__ void first_failure_kernel(int64_t count, const int32_t* __restrict__ status, unsigned long long* first) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count && status[i] != 0) atomicMin(first, (unsigned long long)i);
}

}  // namespace gprto
