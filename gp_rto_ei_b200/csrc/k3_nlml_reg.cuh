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
#include "gprto_math.cuh"

namespace gprto {

constexpr int K3R_MAXN = 128;
constexpr int K3R_MAXD = 8;
constexpr int K3R_LDP = 12;

template <int NT>
constexpr size_t k3r_smem_bytes(int d) {
  return sizeof(double) * (2 * size_t(8 * NT) * K3R_LDP + size_t(8 * NT) * d + 8 * NT + 16 + 8 + 16);
}

template <int NT, int DT>
__global__ void __launch_bounds__(16 * NT, (NT >= 12 ? 2 : (NT >= 6 ? 4 : 8))) k3_nlml_tile_kernel(int S, int n, int d_rt, const double* __restrict__ Xn,
                                                               const double* __restrict__ Yn,
                                                               const double* __restrict__ theta_all, double jitter,
                                                               double* __restrict__ nll, int32_t* __restrict__ status) {
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
  const int64_t b = bs / S;
  const double* theta = theta_all + bs * (d + 2);

  for (int e = tid; e < NP * d; e += THREADS) Xs[e] = (e / d) < n ? Xn[b * int64_t(n) * d + e] : 0.0;
  for (int e = tid; e < NP; e += THREADS) ys[e] = e < n ? Yn[b * n + e] : 0.0;
  if (tid < d) nhw[tid] = -0.5 * exp(-2.0 * theta[tid]);
  if (tid == 0) s_info = 0;
  __syncthreads();

  const double c0 = 2.0 * theta[d];
  const double diag_add = exp(2.0 * theta[d + 1]) + jitter;
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
    // (2) diagonal tile by warp 0: lane l < 8 owns row l, lane 8 owns the right-hand side block y'_kb as a ninth row
    // (so z_kb = L11^-1 y'_kb falls out of the same elimination); pivots travel by shuffle, 1/sqrt by rsqrt_fast
    if (w == 0) {
      const int l = lane < 9 ? lane : 0;
      double dg[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) dg[c] = l < 8 ? buf[(8 * kb + l) * LDP + c] : ys[8 * kb + c];
      double mypiv = 1.0, myr = 0.0;
      int info = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double piv = __shfl_sync(0xffffffffu, dg[c], c);
        if (!(piv > 0.0) && info == 0) info = 8 * kb + c + 1;
        const double r = rsqrt_fast(piv);
        if (l == c) { mypiv = piv; myr = r; }
        dg[c] = (l == c) ? piv * r : dg[c] * r;
#pragma unroll
        for (int j = c + 1; j < 8; ++j) {
          const double ljc = __shfl_sync(0xffffffffu, dg[c], j);
          dg[j] = fma(-dg[c], ljc, dg[j]);
        }
      }
      double lg = 0.0;
      if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) buf[(8 * kb + l) * LDP + c] = dg[c];
        rinv_s[l] = myr;
        lg = log(mypiv);
      } else if (lane == 8) {
        double q = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) { zs[c] = dg[c]; q = fma(dg[c], dg[c], q); }
        zs[8] = q;
      }
      lg += __shfl_xor_sync(0xffffffffu, lg, 1);
      lg += __shfl_xor_sync(0xffffffffu, lg, 2);
      lg += __shfl_xor_sync(0xffffffffu, lg, 4);
      if (lane == 0) {
        logdet += lg;
        if (info != 0) s_info = info;
      }
    }
    __syncthreads();
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
      if (tid == 0) quad += zs[8];
    }
    __syncthreads();
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
  }

  if (tid == 0) {
    const int info = s_info;
    status[bs] = info;
    nll[bs] = info == 0 ? quad + logdet : CUDART_NAN;
  }
}

template <int NT, int DT>
int launch_k3r_d(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                 double* nll, int32_t* status, cudaStream_t s) {
  const size_t smem = k3r_smem_bytes<NT>(d);
  static bool configured = false;
  if (!configured) {
    if (k3r_smem_bytes<NT>(K3R_MAXD) > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k3_nlml_tile_kernel<NT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           int(k3r_smem_bytes<NT>(K3R_MAXD)));
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    // without this the driver may size the shared-memory carve-out for a single block of this kernel
    cudaFuncSetAttribute(k3_nlml_tile_kernel<NT, DT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  k3_nlml_tile_kernel<NT, DT><<<(unsigned)(B * S), 16 * NT, smem, s>>>(S, n, d, Xn, Yn, theta, jitter, nll, status);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

template <int NT>
int launch_k3r(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
               double* nll, int32_t* status, cudaStream_t s) {
  switch (d) {   // the dimensions of the published problems get unrolled distance loops
    case 2: return launch_k3r_d<NT, 2>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
    case 3: return launch_k3r_d<NT, 3>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
  }
  return launch_k3r_d<NT, 0>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
}

inline int launch_nlml_reg(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta,
                           double jitter, double* nll, int32_t* status, cudaStream_t s) {
  if (n > K3R_MAXN || d > K3R_MAXD) return GPRTO_EUNSUPPORTED;
  const int nt = (n + 15) / 16 * 2;       // even number of 8-row tiles
  switch (nt) {
    case 2: return launch_k3r<2>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
    case 4: return launch_k3r<4>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
    case 6: return launch_k3r<6>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
    case 8: return launch_k3r<8>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
    case 10: return launch_k3r<10>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
    case 12: return launch_k3r<12>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
    case 14: return launch_k3r<14>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
    case 16: return launch_k3r<16>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s);
  }
  return GPRTO_EUNSUPPORTED;
}

}  // namespace gprto
