// K4: fused cross-covariance -> posterior mean/variance -> acquisition -> per-GP arg-min.
// Replaces GP_model.calc_cov_sample / GP_inference_np and ITR_GP_RTO.GP_obj_f
// (reference: sub_uts/utilities_2.py:74-85, 235-278, 477-507).
#pragma once
#include "gprto_math.cuh"

namespace gprto {

struct K4Params {
  int64_t B, m;
  int n, d, acq;
  int64_t chunk;       // candidates per CTA (multiple of 32*warps)
  int nchunks;         // CTAs per GP
  const double *Xn, *invK, *alpha, *theta, *x_mean, *x_std, *y_mean, *y_std, *cand, *prior, *f_best;
  // fused sampling mode (cand == nullptr): candidate c of GP b = tr_point(seed, b, c, shape, centre[b], radius[b])
  const double *centre, *radius;
  uint64_t seed;
  int shape;
  double *mean, *var, *score;
  double* part_score;  // [B, nchunks] partial arg-min (nullptr if best not requested)
  int64_t* part_idx;
};

constexpr int K4_WARPS = 4;
constexpr int K4_THREADS = K4_WARPS * 32;
constexpr int K4_MINB = 4;               // CTAs per SM asked of ptxas for NB <= 8 (128 registers; 3, 5 and 6 measured the same)

// Data index handled by k-step s on a lane with threadID-in-group t: within an 8-block the four
// lanes of a quad own {2t, 2t+1}, which is exactly the pair of output columns the same lane holds in
// the DMMA accumulator, so the final k'(Mk) contraction needs no data movement.
__host__ __device__ constexpr int k4_pi(int s, int t) { return 8 * (s >> 1) + 2 * t + (s & 1); }

template <int NB>
__host__ __device__ constexpr int k4_frags() { return NB * (NB + 1); }

template <int NB, int D>
__host__ __device__ constexpr size_t k4_smem_bytes() {
  return sizeof(double) * (size_t(k4_frags<NB>()) * 32 + size_t(8 * NB) * D + 8 * NB + 3 * D + 8) +
         sizeof(double) * K4_WARPS + sizeof(int64_t) * K4_WARPS;
}

// One CTA = one GP x one chunk of its candidates.  Each warp walks the chunk 32 candidates at a time as
// four 8-candidate DMMA tiles:
//   A fragment (8 candidates x 4 data points)  = cross-covariances k*, computed in registers (2 NB exps per lane)
//   B fragment (4 data points x 8 rows of M)    = folded inverse M = tril(invK + invK') - diag(invK), staged once
//                                                 per CTA in shared memory in fragment order (conflict-free LDS.64)
//   C accumulators                              = (M k*) for the lane's own data points
// Only the NB(NB+1) non-zero fragments of the triangular M are visited: n(n+1)/2-type work instead of n^2.
template <int NB, int D>
__global__ void __launch_bounds__(K4_THREADS, (NB <= 8 ? K4_MINB : 1)) k4_dmma_kernel(const K4Params p) {
  constexpr int NP = 8 * NB;       // padded n
  constexpr int NS = 2 * NB;       // k-steps
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Mf = reinterpret_cast<double*>(smem_raw);          // [frags][32]
  double* Xs = Mf + k4_frags<NB>() * 32;                     // [NP][D] normalised data points
  double* As = Xs + NP * D;                                  // [NP] alpha
  double* cs = As + NP;                                      // [D] candidate scale a_d
  double* cm = cs + D;                                       // [D] x_mean
  double* sc = cm + D;                                       // [D] -1/(2 W_d)
  double* scal = sc + D;                                     // [8]: c0, sf2, y_std, y_mean, f_best
  double* red_s = scal + 8;                                  // [warps]
  int64_t* red_i = reinterpret_cast<int64_t*>(red_s + K4_WARPS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t b = blockIdx.x;
  const int n = p.n;
  const double* theta = p.theta + b * (D + 2);
  const double* invK = p.invK + b * int64_t(n) * n;

  // ---- prologue: stage the GP ----
  for (int e = tid; e < k4_frags<NB>() * 32; e += K4_THREADS) {
    const int f = e >> 5, l = e & 31;
    int bb = 0;
    while ((bb + 1) * (bb + 2) <= f) ++bb;        // f in [bb(bb+1), (bb+1)(bb+2))
    const int s = f - bb * (bb + 1);
    const int i = 8 * bb + (l >> 2), j = k4_pi(s, l & 3);
    double v = 0.0;
    if (i < n && j < n) {
      if (j < i) v = invK[int64_t(i) * n + j] + invK[int64_t(j) * n + i];
      else if (j == i) v = invK[int64_t(i) * n + i];
    }
    Mf[e] = v;
  }
  for (int e = tid; e < NP * D; e += K4_THREADS) {
    const int j = e / D, dd = e - j * D;
    Xs[e] = j < n ? p.Xn[(b * n + j) * D + dd] : 0.0;
  }
  for (int e = tid; e < NP; e += K4_THREADS) As[e] = e < n ? p.alpha[b * n + e] : 0.0;
  if (tid < D) {
    cs[tid] = 1.0 / p.x_std[b * D + tid];
    cm[tid] = p.x_mean[b * D + tid];
    sc[tid] = -0.5 * exp_det(-2.0 * theta[tid]);      // -1/(2 W_d)
  }
  if (tid == 0) {
    scal[0] = 2.0 * theta[D];
    scal[1] = exp_det(2.0 * theta[D]);
    scal[2] = p.y_std[b];
    scal[3] = p.y_mean[b];
    scal[4] = p.f_best ? p.f_best[b] : 0.0;
  }
  __syncthreads();

  const double c0 = scal[0], sf2 = scal[1], y_std = scal[2], y_mean = scal[3], f_best = scal[4];
  double a_d[D], mu_d[D], nhw[D];
#pragma unroll
  for (int dd = 0; dd < D; ++dd) { a_d[dd] = cs[dd]; mu_d[dd] = cm[dd]; nhw[dd] = sc[dd]; }

  const int64_t c_begin = int64_t(blockIdx.y) * p.chunk;
  const int64_t c_end = min(p.m, c_begin + p.chunk);
  const double* cand = p.cand ? p.cand + b * p.m * D : nullptr;
  const double* Xt = Xs + 2 * t * D;       // + (8*(s>>1) + (s&1))*D + dd   (compile-time offsets)
  const double* At = As + 2 * t;
  const double* Mlane = Mf + lane;

  double best_v = CUDART_INF;
  int64_t best_i = INT64_MAX;

  for (int64_t base = c_begin + warp * 32; base < c_end; base += K4_WARPS * 32) {
    const int64_t ce = base + 8 * t + g;          // the candidate this lane finishes in the epilogue
    const bool valid = ce < c_end;
    double xe[D];
    if (cand) {
#pragma unroll
      for (int dd = 0; dd < D; ++dd) xe[dd] = valid ? (cand[ce * D + dd] - mu_d[dd]) * a_d[dd] : 0.0;
    } else {
      double xp[D];
      tr_point<D>(p.seed, b, ce, p.shape, p.centre + b * D, p.radius + b * (p.shape == 2 ? D + 1 : 1), xp);
#pragma unroll
      for (int dd = 0; dd < D; ++dd) xe[dd] = valid ? (xp[dd] - mu_d[dd]) * a_d[dd] : 0.0;
    }

    double q_mine = 0.0, m_mine = 0.0;
#pragma unroll 1                         // one tile at a time: two at once spill (measured slower)
    for (int u = 0; u < 4; ++u) {
      double x[D];
#pragma unroll
      for (int dd = 0; dd < D; ++dd) x[dd] = __shfl_sync(0xffffffffu, xe[dd], 4 * g + u);
      // cross-covariances for candidate (8u + g) and data points pi(s, t)
      double kv[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        double acc = c0;
#pragma unroll
        for (int dd = 0; dd < D; ++dd) {
          // difference first, scale second: keeps the rounding error relative to |x - X| (as the reference's
          // cdist does) instead of to |x|/length-scale
          const double df = x[dd] - Xt[(8 * (s >> 1) + (s & 1)) * D + dd];
          acc = fma(df * df, nhw[dd], acc);
        }
        kv[s] = exp_fast(acc);
      }
      double C0[NB], C1[NB];
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) { C0[bb] = 0.0; C1[bb] = 0.0; }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int bb = (s >> 1); bb < NB; ++bb) {
          dmma884(C0[bb], C1[bb], kv[s], Mlane[(bb * (bb + 1) + s) * 32]);
        }
      }
      double qp = 0.0, mp = 0.0;
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) {
        qp = fma(C0[bb], kv[2 * bb], qp);
        qp = fma(C1[bb], kv[2 * bb + 1], qp);
        mp = fma(kv[2 * bb], At[8 * bb], mp);
        mp = fma(kv[2 * bb + 1], At[8 * bb + 1], mp);
      }
      qp += __shfl_xor_sync(0xffffffffu, qp, 1);
      mp += __shfl_xor_sync(0xffffffffu, mp, 1);
      qp += __shfl_xor_sync(0xffffffffu, qp, 2);
      mp += __shfl_xor_sync(0xffffffffu, mp, 2);
      if (t == u) { q_mine = qp; m_mine = mp; }
    }

    if (valid) {
      const double mean = fma(m_mine, y_std, y_mean);
      const double var = fmax(0.0, sf2 - q_mine) * (y_std * y_std);
      const int64_t o = b * p.m + ce;
      if (p.mean) p.mean[o] = mean;
      if (p.var) p.var[o] = var;
      if (p.score || p.part_score) {
        const double pr = p.prior ? p.prior[o] : 0.0;
        const double sv = acquisition(p.acq, mean, var, pr, f_best);
        if (p.score) p.score[o] = sv;
        if (better(sv, ce, best_v, best_i)) { best_v = sv; best_i = ce; }
      }
    }
  }

  if (p.part_score) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best_v, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, best_i, off);
      if (better(ov, oi, best_v, best_i)) { best_v = ov; best_i = oi; }
    }
    if (lane == 0) { red_s[warp] = best_v; red_i[warp] = best_i; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < K4_WARPS; ++w)
        if (better(red_s[w], red_i[w], best_v, best_i)) { best_v = red_s[w]; best_i = red_i[w]; }
      p.part_score[b * p.nchunks + blockIdx.y] = best_v;
      p.part_idx[b * p.nchunks + blockIdx.y] = best_i;
    }
  }
}

#ifdef GPRTO_API_TU      // non-template kernels: defined in the API translation unit only
// Generic path: any n <= K4G_MAXN, any d <= K4G_MAXD; one thread per candidate, k* in local memory,
// full (non-symmetrised) quadratic form exactly as the reference writes it (:267-268).  Used outside the
// DMMA path's limits and as an independent cross-check of it in the tests.
constexpr int K4G_MAXN = 512;
constexpr int K4G_MAXD = 16;
constexpr int K4G_THREADS = 128;

__global__ void __launch_bounds__(K4G_THREADS) k4_generic_kernel(const K4Params p) {
  const int64_t b = blockIdx.x;
  const int n = p.n, d = p.d;
  const double* theta = p.theta + b * (d + 2);
  const double* invK = p.invK + b * int64_t(n) * n;
  const double* Xn = p.Xn + b * int64_t(n) * d;
  const double* alpha = p.alpha + b * n;
  __shared__ double red_s[K4G_THREADS / 32];
  __shared__ int64_t red_i[K4G_THREADS / 32];
  double invW[K4G_MAXD];
  for (int dd = 0; dd < d; ++dd) invW[dd] = exp_det(-2.0 * theta[dd]);
  const double sf2 = exp_det(2.0 * theta[d]);
  const double y_std = p.y_std[b], y_mean = p.y_mean[b];
  const double f_best = p.f_best ? p.f_best[b] : 0.0;
  const int64_t c_begin = int64_t(blockIdx.y) * p.chunk;
  const int64_t c_end = min(p.m, c_begin + p.chunk);
  double best_v = CUDART_INF;
  int64_t best_i = INT64_MAX;
  double k[K4G_MAXN];
  for (int64_t c = c_begin + threadIdx.x; c < c_end; c += K4G_THREADS) {
    double x[K4G_MAXD];
    for (int dd = 0; dd < d; ++dd)
      x[dd] = (p.cand[(b * p.m + c) * d + dd] - p.x_mean[b * d + dd]) / p.x_std[b * d + dd];
    double mn = 0.0;
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int dd = 0; dd < d; ++dd) {
        const double df = x[dd] - Xn[j * d + dd];
        s = fma(df * df, invW[dd], s);
      }
      k[j] = sf2 * exp_det(-0.5 * s);
      mn = fma(k[j], alpha[j], mn);
    }
    double q = 0.0;
    for (int i = 0; i < n; ++i) {
      double r = 0.0;
      for (int j = 0; j < n; ++j) r = fma(invK[int64_t(i) * n + j], k[j], r);
      q = fma(k[i], r, q);
    }
    const double mean = fma(mn, y_std, y_mean);
    const double var = fmax(0.0, sf2 - q) * (y_std * y_std);
    const int64_t o = b * p.m + c;
    if (p.mean) p.mean[o] = mean;
    if (p.var) p.var[o] = var;
    if (p.score || p.part_score) {
      const double pr = p.prior ? p.prior[o] : 0.0;
      const double sv = acquisition(p.acq, mean, var, pr, f_best);
      if (p.score) p.score[o] = sv;
      if (better(sv, c, best_v, best_i)) { best_v = sv; best_i = c; }
    }
  }
  if (p.part_score) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best_v, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, best_i, off);
      if (better(ov, oi, best_v, best_i)) { best_v = ov; best_i = oi; }
    }
    if (lane == 0) { red_s[warp] = best_v; red_i[warp] = best_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < K4G_THREADS / 32; ++w)
        if (better(red_s[w], red_i[w], best_v, best_i)) { best_v = red_s[w]; best_i = red_i[w]; }
      p.part_score[b * p.nchunks + blockIdx.y] = best_v;
      p.part_idx[b * p.nchunks + blockIdx.y] = best_i;
    }
  }
}

#endif  // GPRTO_API_TU

// Materialise the sampler's candidates (same function as K4's fused mode): cand[b, c, :] = centre[b] + radius[b] * u.
// With idx != nullptr only candidate idx[b] of each GP is produced (out[b, :]) - the coordinates of the per-GP best.
template <int D>
__global__ void sample_candidates_kernel(int64_t B, int64_t m, const double* __restrict__ centre, const double* __restrict__ radius,
                                         int shape, uint64_t seed, const int64_t* __restrict__ idx, double* __restrict__ out) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = idx ? B : B * m;
  if (e >= total) return;
  const int64_t b = idx ? e : e / m;
  const int64_t c = idx ? idx[e] : e - b * m;
  double x[D];
  tr_point<D>(seed, b, c, shape, centre + b * D, radius + b * (shape == 2 ? D + 1 : 1), x);
#pragma unroll
  for (int dd = 0; dd < D; ++dd) out[e * D + dd] = x[dd];
}

#ifdef GPRTO_API_TU
// Row-wise arg-min (numpy.argmin semantics) - also the second pass of K4's per-GP best and the
// multistart selection of utilities_2.py:166.  One warp per row.
__global__ void argmin_rows_kernel(int64_t B, int64_t S, const double* __restrict__ val,
                                   const int64_t* __restrict__ idx_in, double* __restrict__ best_val,
                                   int64_t* __restrict__ best_idx) {
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  double bv = CUDART_INF;
  int64_t bi = INT64_MAX;
  for (int64_t s = lane; s < S; s += 32) {
    const double v = val[row * S + s];
    const int64_t i = idx_in ? idx_in[row * S + s] : s;
    if (better(v, i, bv, bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { best_val[row] = bv; best_idx[row] = bi; }
}

#endif  // GPRTO_API_TU

}  // namespace gprto
