// K3 register-resident path: one CTA per (GP, hyper-parameter start); the whole lower triangle of
// K + (sn2 + jitter) I lives in REGISTERS as 8x8 FP64 tensor-core accumulator tiles and is factorised by a
// right-looking blocked Cholesky whose trailing updates are DMMA.8x8x4 instructions.
// Replaces GP_model.negative_loglikelihood (reference: sub_uts/utilities_2.py:92-114) for n <= 128.
//
//   tiles      (I, J), J <= I, I, J < NT, n_pad = 8 NT.  Warp w of NT/2 owns tile rows w and NT-1-w: NT+1 tiles
//              each (perfectly balanced), 2 doubles per lane per tile (the m8n8k4 C-fragment layout).
//   block step kb: (1) owners spill tile column kb to a shared-memory panel [n_pad][8] (row stride 12 doubles: both the
//              fragment loads and the row loads below are bank-conflict free or 2-way at worst);
//              (2) warp 0 factorises the 8x8 diagonal tile with shuffles (rsqrt + multiply, no divide on the chain);
//              the right-hand side block rides along as a ninth row (z = L^-1 y block by block, y'K^-1y = |z|^2), so no
//              separate triangular solve is needed;
//              (3) one thread per remaining row solves its 8 panel entries (L21 = A21 L11^-T);
//              (4) every warp updates its tiles with J > kb: C(I,J) -= L(I,kb) L(J,kb)' = 2 DMMA per tile; the warp that
//              holds diagonal tile (J,J) also applies y'_J -= L(J,kb) z_kb from the fragments it already loaded.
//              The panel is double-buffered: 3 barriers per block step.
//   K itself is never written to memory: each lane evaluates the SE-ARD entries of its own fragments.
#pragma once
#include "gprto_launch.h"
#include "gprto_math.cuh"
#ifdef K3_TIMING
#include <cstdio>
#endif

namespace gprto {

constexpr int K3R_LDP = 12;

template <int NT>
constexpr size_t k3r_smem_bytes(int d) {
  return sizeof(double) * (2 * size_t(8 * NT) * K3R_LDP + size_t(8 * NT) * d + 8 * NT + 16 + 8 + 16);
}

template <int NT, int DT>
__global__ void __launch_bounds__(16 * NT, (NT >= 12 ? 2 : (NT >= 6 ? 4 : 8))) k3_nlml_tile_kernel(int S, int n, int d_rt, const double* __restrict__ Xn,
                                                               const double* __restrict__ Yn,
                                                               const double* __restrict__ theta_all, double jitter,
                                                               double* __restrict__ nll, int32_t* __restrict__ status,
                                                               const int32_t* __restrict__ gidx) {
  constexpr int W = NT / 2, NP = 8 * NT, SLOTS = NT + 1, LDP = K3R_LDP, THREADS = 32 * W;
  const int d = DT > 0 ? DT : d_rt;      // DT > 0: compile-time input dimension (unrolled distance loops)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* P = reinterpret_cast<double*>(smem_raw);   // [2][NP][LDP]
  double* Xs = P + 2 * NP * LDP;                     // [NP][d]
  double* ys = Xs + NP * d;                          // [NP]
  double* nhw = ys + NP;                             // [16]
  double* rinv_s = nhw + 16;                         // [8]
  double* zs = rinv_s + 8;                           // [9] current block of z = L^-1 y, then |z_kb|^2
  __shared__ int s_info;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t bs = blockIdx.x;
  const int64_t b = gidx ? gidx[bs] : bs / S;      // gidx: problem -> data set indirection (indexed entry point)
  const double* theta = theta_all + bs * (d + 2);

  for (int e = tid; e < NP * d; e += THREADS) Xs[e] = (e / d) < n ? Xn[b * int64_t(n) * d + e] : 0.0;
  for (int e = tid; e < NP; e += THREADS) ys[e] = e < n ? Yn[b * n + e] : 0.0;
  if (tid < d) nhw[tid] = -0.5 * exp_det(-2.0 * theta[tid]);
  if (tid == 0) s_info = 0;
  __syncthreads();

  const double c0 = 2.0 * theta[d];
  const double diag_add = exp_det(2.0 * theta[d + 1]) + jitter;
  const int rowA = w, rowB = NT - 1 - w;

  // ---- build the tiles of K in registers (branch-free padding: identity block outside n)
  double C0[SLOTS], C1[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int I = (s <= w) ? rowA : rowB;
    const int J = (s <= w) ? s : s - w - 1;
    const int i = 8 * I + g, j0 = 8 * J + 2 * t;
    const double* xi = Xs + i * d;
    const double* xj = Xs + j0 * d;
    double a0 = c0, a1 = c0;
    if (DT > 0) {
#pragma unroll
      for (int dd = 0; dd < (DT > 0 ? DT : 1); ++dd) {
        const double x = xi[dd], h = nhw[dd];
        const double d0 = x - xj[dd], d1 = x - xj[DT + dd];
        a0 = fma(d0 * d0, h, a0);
        a1 = fma(d1 * d1, h, a1);
      }
    } else {
      for (int dd = 0; dd < d; ++dd) {
        const double x = xi[dd], h = nhw[dd];
        const double d0 = x - xj[dd], d1 = x - xj[d + dd];
        a0 = fma(d0 * d0, h, a0);
        a1 = fma(d1 * d1, h, a1);
      }
    }
    const double v0 = exp_fast(a0) + (i == j0 ? diag_add : 0.0);
    const double v1 = exp_fast(a1) + (i == j0 + 1 ? diag_add : 0.0);
    const bool in0 = i < n && j0 < n, in1 = i < n && j0 + 1 < n;
    C0[s] = in0 ? v0 : (i == j0 ? 1.0 : 0.0);
    C1[s] = in1 ? v1 : (i == j0 + 1 ? 1.0 : 0.0);
  }

  double logdet = 0.0, quad = 0.0;
#ifdef K3_TIMING
  long long tk[6] = {0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define K3_MARK(i) { const long long tn = clock64(); tk[i] += tn - tprev; tprev = tn; }
#else
#define K3_MARK(i)
#endif
  K3_MARK(0)

#pragma unroll 1
  for (int kb = 0; kb < NT; ++kb) {
    double* buf = P + (kb & 1) * NP * LDP;
    // (1) spill tile column kb
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int I = (s <= w) ? rowA : rowB;
      const int J = (s <= w) ? s : s - w - 1;
      if (J == kb) *reinterpret_cast<double2*>(&buf[(8 * I + g) * LDP + 2 * t]) = make_double2(C0[s], C1[s]);
    }
    __syncthreads();
    K3_MARK(1)
    // (2) diagonal tile by warp 0: lane l < 8 owns row l, lane 8 owns the right-hand side block y'_kb as a ninth row
    // (so z_kb = L11^-1 y'_kb falls out of the same elimination).  The n-long pivot chain is the latency bottleneck of
    // the whole kernel, so the elimination is DIVISION- AND SQRT-FREE: step c forms
    //     a_jk <- d~ a_jk - (a_jc 2^-e) a_kc,   d~ = d 2^-e in [1,2),  e = exponent(d),  d = current (scaled) pivot,
    // i.e. the Schur complement times d~ (a power-of-two normalised fraction-free step; the accumulated scale
    // S_c = prod_{c'<c} d~_c' stays below 2^8).  The eight 1/sqrt are then independent:  L_jc = a'_jc rsqrt(d'_c S_c),
    // one per lane, in parallel, after the chain (chain per column: shuffle, 3 integer ops, multiply, shuffle, FMA).
    // Layout: the tile is spread over the warp like a DMMA accumulator (lane 4r+q holds a[r][2q], a[r][2q+1]) and the
    // right-hand side block sits on lanes 0..7, so a column step is 6 shuffles + 6 FP64 instructions per lane instead
    // of a 7-long serial loop per row.
    if (w == W - 1) {      // last warp: the scheduler favours higher warp ids and this chain is the critical path
      const int r = lane >> 2, q = lane & 3, k8 = lane & 7;
      const double2 v = *reinterpret_cast<const double2*>(&buf[(8 * kb + r) * LDP + 2 * q]);
      double x0 = v.x, x1 = v.y;
      double yv = ys[8 * kb + k8];
      double S = 1.0, myS = 1.0, mydp = 1.0;
      int info = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int cs = c >> 1;
        const double colsel = (c & 1) ? x1 : x0;                                   // column c of my row (valid where q == cs)
        const double dpiv = __shfl_sync(0xffffffffu, colsel, 4 * c + cs);          // a[c][c]
        const double arc = __shfl_sync(0xffffffffu, colsel, 4 * r + cs);           // a[r][c]
        const double ak0 = __shfl_sync(0xffffffffu, colsel, 4 * (2 * q) + cs);     // a[2q][c]
        const double ak1 = __shfl_sync(0xffffffffu, colsel, 4 * (2 * q + 1) + cs); // a[2q+1][c]
        const double ayk = __shfl_sync(0xffffffffu, colsel, 4 * k8 + cs);          // a[k8][c]
        const double yc = __shfl_sync(0xffffffffu, yv, c);
        if (!(dpiv > 0.0) && info == 0) info = 8 * kb + c + 1;
        if (lane == c) { myS = S; mydp = dpiv; }
        const int e = ((__double2hiint(dpiv) >> 20) & 0x7ff) - 1023;
        const double f = __hiloint2double((1023 - e) << 20, 0);                    // 2^-e, exact
        const double dt = dpiv * f;
        const double ac = arc * f, ycf = yc * f;
        const double n0 = fma(-ac, ak0, dt * x0), n1 = fma(-ac, ak1, dt * x1);
        const double ny = fma(-ycf, ayk, dt * yv);
        x0 = (2 * q > c) ? n0 : x0;                                                // columns <= c are final (frozen at scale S_col)
        x1 = (2 * q + 1 > c) ? n1 : x1;
        yv = (k8 > c) ? ny : yv;
        S *= dt;
      }
      const double tmine = lane < 8 ? rsqrt_fast(mydp * myS) : 1.0;                // lane c: rsqrt(d'_c S_c)
      const double t0 = __shfl_sync(0xffffffffu, tmine, 2 * q), t1 = __shfl_sync(0xffffffffu, tmine, 2 * q + 1);
      *reinterpret_cast<double2*>(&buf[(8 * kb + r) * LDP + 2 * q]) = make_double2(x0 * t0, x1 * t1);
      double zq = 0.0, lg = 0.0;
      if (lane < 8) {
        const double z = yv * tmine;
        zs[lane] = z;
        zq = z * z;
        rinv_s[lane] = tmine * myS;                                                // 1/L_cc = rsqrt(d'_c / S_c)
        lg = 2.0 * log_det(mydp * tmine);                                              // log pivot = 2 log L_cc
      }
#pragma unroll
      for (int off = 1; off < 8; off <<= 1) {
        lg += __shfl_xor_sync(0xffffffffu, lg, off);
        zq += __shfl_xor_sync(0xffffffffu, zq, off);
      }
      if (lane == 0) {
        logdet += lg;
        zs[8] = zq;
        if (info != 0) s_info = info;
      }
    }
    K3_MARK(2)
    __syncthreads();
    K3_MARK(5)
    if (s_info != 0) break;
    // (3) panel rows (L21 = A21 L11^-T), one thread per remaining row
    {
      const int nrows = NP - 8 * (kb + 1);
      if (tid < nrows) {
        const int row = 8 * (kb + 1) + tid;
        double x[8], rv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) rv[c] = rinv_s[c];
        const double2* pr = reinterpret_cast<const double2*>(&buf[row * LDP]);
#pragma unroll
        for (int c = 0; c < 4; ++c) { const double2 v = pr[c]; x[2 * c] = v.x; x[2 * c + 1] = v.y; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          x[c] *= rv[c];
#pragma unroll
          for (int c2 = c + 1; c2 < 8; ++c2) x[c2] = fma(-buf[(8 * kb + c2) * LDP + c], x[c], x[c2]);   // L11[c2][c], broadcast
        }
        double2* pw = reinterpret_cast<double2*>(&buf[row * LDP]);
#pragma unroll
        for (int c = 0; c < 4; ++c) pw[c] = make_double2(x[2 * c], x[2 * c + 1]);
      }
      if (tid == 32 * (W - 1)) quad += zs[8];
    }
    __syncthreads();
    K3_MARK(3)
    // (4) trailing update with the FP64 tensor pipe
    double aA[2], aB[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      aA[s] = buf[(8 * rowA + g) * LDP + 4 * s + t];
      aB[s] = buf[(8 * rowB + g) * LDP + 4 * s + t];
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int J = (s <= w) ? s : s - w - 1;
      if (J > kb) {
        const double b0 = -buf[(8 * J + g) * LDP + t];
        const double b1 = -buf[(8 * J + g) * LDP + 4 + t];
        dmma884(C0[s], C1[s], (s <= w) ? aA[0] : aB[0], b0);
        dmma884(C0[s], C1[s], (s <= w) ? aA[1] : aB[1], b1);
        const int I = (s <= w) ? rowA : rowB;
        if (I == J) {      // the diagonal tile of row block J also carries y'_J -= L(J,kb) z_kb  (b0, b1 are already negated)
          double part = fma(b0, zs[t], b1 * zs[4 + t]);
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          if (t == 0) ys[8 * J + g] += part;
        }
      }
    }
    K3_MARK(4)
  }
#ifdef K3_TIMING
  if (tid == 0 && (bs == 0 || bs == 20000)) printf("K3 timing block %lld: build %lld spill+bar %lld diag %lld (bar after diag %lld) panel+bar %lld update %lld cycles\n", (long long)bs, tk[0], tk[1], tk[2], tk[5], tk[3], tk[4]);
#endif

  if (tid == 32 * (W - 1)) {      // lane 0 of the diagonal-tile warp holds logdet and quad
    const int info = s_info;
    status[bs] = info;
    nll[bs] = info == 0 ? quad + logdet : CUDART_NAN;
  }
}

template <int NT, int DT>
int launch_k3r_d(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                 double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  const size_t smem = k3r_smem_bytes<NT>(d);
  static PerDeviceOnce once;
  if (once.first()) {
    if (k3r_smem_bytes<NT>(K3R_MAXD) > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k3_nlml_tile_kernel<NT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           int(k3r_smem_bytes<NT>(K3R_MAXD)));
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    // without this the driver may size the shared-memory carve-out for a single block of this kernel
    cudaFuncSetAttribute(k3_nlml_tile_kernel<NT, DT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  k3_nlml_tile_kernel<NT, DT><<<(unsigned)(B * S), 16 * NT, smem, s>>>(S, n, d, Xn, Yn, theta, jitter, nll, status, gidx);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

template <int NT>
int launch_k3r(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
               double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  switch (d) {   // the dimensions of the published problems get unrolled distance loops
    case 2: return launch_k3r_d<NT, 2>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 3: return launch_k3r_d<NT, 3>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return launch_k3r_d<NT, 0>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
}

#ifdef GPRTO_K3_TU
int launch_nlml_reg(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                    double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx) {
  if (n > K3R_MAXN || d > K3R_MAXD) return GPRTO_EUNSUPPORTED;
  const int nt = (n + 15) / 16 * 2;       // even number of 8-row tiles
  switch (nt) {
    case 2: return launch_k3r<2>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 4: return launch_k3r<4>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 6: return launch_k3r<6>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 8: return launch_k3r<8>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 10: return launch_k3r<10>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 12: return launch_k3r<12>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 14: return launch_k3r<14>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 16: return launch_k3r<16>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return GPRTO_EUNSUPPORTED;
}
#endif  // GPRTO_K3_TU

}  // namespace gprto
