// K2, register-tile / DMMA variant (32 < n <= 128): final factor of a fitted GP on the design of the look-ahead NLML
// kernel (k3_nlml_la.cuh).  Replaces the tail of GP_model.determine_hyperparameters (reference:
// sub_uts/utilities_2.py:168-175: Kopt = K + sn2 I, invKopt = solve(Kopt, I)) and the alpha = invK y that
// GP_inference_np re-forms on every call (:267).
//
// One CTA (8 warps) per GP.  Four phases, all on 8x8 FP64 tensor-core tiles:
//   1  Cholesky K = L L' exactly as k3_nlml_la_kernel (tiles of K in registers, dedicated panel warp one column
//      ahead, scaled division-free diagonal tile, DMMA trailing update) - but every tile column is spilled to ITS OWN
//      shared-memory panel, so that L (packed by tile columns, tile (I,J) at index tau = off(J) + I - J) survives the
//      factorisation.  Riding along: z = L^-1 y (extra row of the panel solve, as in K3) and the inverses of the
//      diagonal tiles (eight more "rows" e_c L11^-T of the same panel solve), which replace the diagonal tiles of L.
//   2  X = L^-1 in place, block row by block row: X(I,J) = -Linv(I,I) sum_{J<=K<I} L(I,K) X(K,J); row I of L is dead
//      once row I of X exists, so X overwrites L (two barriers per block row).  The accumulator is turned into the
//      B operand of the second product with two shuffles per k-step.
//   3  invK = X'X: tile (I,J) = sum_{K>=I} X(K,I)' X(K,J), all tiles independent, straight from the panels to global
//      memory (lower triangle computed once, mirrored on the store -> exactly symmetric output).
//   4  alpha = X'z, log det from the running pivot product of phase 1.
// Shared memory: 8*TILES rows x 12 doubles (104 KB at n = 128: two CTAs per SM, 28 KB at n = 64).
#pragma once
#include "gprto_launch.h"
#include "gprto_math.cuh"
#include "k3_nlml_la.cuh"

namespace gprto {

template <int NT>
constexpr size_t k2la_smem_bytes(int d) {
  return sizeof(double) * (size_t(8 * K3LA<NT>::TILES) * K3R_LDP + size_t(8 * NT) * d + 2 * size_t(8 * NT) + 16 + 8 + 8) +
         2 * K3LA<NT>::TILES + 16;
}

template <int NT, int DT>
__global__ void __launch_bounds__(32 * K3LA<NT>::W, (NT <= 8 ? 3 : 2))
k2_factor_la_kernel(int n, int d_rt, const double* __restrict__ Xn, const double* __restrict__ Yn,
                    const double* __restrict__ theta_all, double jitter, double* __restrict__ invK,
                    double* __restrict__ alpha, double* __restrict__ logdet, int32_t* __restrict__ status) {
  using C = K3LA<NT>;
  constexpr int UW = C::UW, TILES = C::TILES, SLOTS = C::SLOTS, NP = C::NP, LDP = K3R_LDP, THREADS = C::THREADS, W = C::W;
  constexpr int UTHREADS = 32 * UW, ATHREADS = 32 * (UW + 1), PANEL_W = W - 1;
  static_assert(C::IDLE == 0, "no idle warp in this kernel");
  static_assert(NP + 9 <= UTHREADS, "one panel-solve row per update thread");
  const int d = DT > 0 ? DT : d_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Lp = reinterpret_cast<double*>(smem_raw);  // [8 * TILES][LDP]: tile tau at rows 8 tau .. 8 tau + 7
  double* Xs = Lp + 8 * TILES * LDP;                 // [NP][d]
  double* ys = Xs + NP * d;                          // [NP] right-hand side, updated in place (y')
  double* zv = ys + NP;                              // [NP] z = L^-1 y
  double* nhw = zv + NP;                             // [16]
  double* rinv_s = nhw + 16;                         // [8]
  double* misc = rinv_s + 8;                         // [8]: [0] log det
  unsigned char* tI = reinterpret_cast<unsigned char*>(misc + 8);   // [TILES]
  unsigned char* tJ = tI + TILES;                                   // [TILES]
  __shared__ int s_info;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t b = blockIdx.x;
  const double* theta = theta_all + b * (d + 2);
  auto off = [](int J) { return J * NT - J * (J - 1) / 2; };      // first tile of tile column J

  for (int e = tid; e < NP * d; e += THREADS) Xs[e] = (e / d) < n ? Xn[b * int64_t(n) * d + e] : 0.0;
  for (int e = tid; e < NP; e += THREADS) ys[e] = e < n ? Yn[b * n + e] : 0.0;
  if (tid < d) nhw[tid] = -0.5 * exp_det(-2.0 * theta[tid]);
  if (tid == 0) s_info = 0;
  for (int e = tid; e < TILES; e += THREADS) {
    int J = 0;
    while ((J + 1) * NT - (J + 1) * J / 2 <= e) ++J;
    tJ[e] = (unsigned char)J;
    tI[e] = (unsigned char)(J + e - (J * NT - J * (J - 1) / 2));
  }
  __syncthreads();

  // ============================================ phase 1: Cholesky ============================================
  if (w == PANEL_W) {
    double pm = 1.0;
    int pe = 0, info = 0;
    k3la_bar_sync(1, ATHREADS);                                         // X(-1): column 0 is in its panel
    k3_diag_tile_lean(Lp, LDP, rinv_s, lane, 0, pm, pe, info);
    if (lane == 0 && info) s_info = info;
    k3la_bar_sync(2, ATHREADS);                                         // Y(-1)
#pragma unroll 1
    for (int kb = 0; kb + 1 < NT; ++kb) {
      if (s_info != 0) break;
      k3la_bar_sync(1, ATHREADS);                                       // X(kb): column kb+1 updated and spilled
      k3_diag_tile_lean(Lp + 8 * off(kb + 1) * LDP, LDP, rinv_s, lane, 8 * (kb + 1), pm, pe, info);
      if (lane == 0 && info) s_info = info;
      k3la_bar_sync(2, ATHREADS);                                       // Y(kb)
    }
    if (lane == 0) misc[0] = log_det(pm) + double(pe) * 0.693147180559945309417;
  } else {
    const int uw = w, utid = 32 * uw + lane;
    const double c0 = 2.0 * theta[d];
    const double diag_add = exp_det(2.0 * theta[d + 1]) + jitter;
    double C0[SLOTS], C1[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int tau = s * UW + uw;
      const int I = tau < TILES ? tI[tau] : 0, J = tau < TILES ? tJ[tau] : 0;
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
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int tau = s * UW + uw;
      if (tau < NT) *reinterpret_cast<double2*>(&Lp[(8 * tau + g) * LDP + 2 * t]) = make_double2(C0[s], C1[s]);
    }
    k3la_bar_arrive(1, ATHREADS);                                       // X(-1)
    k3la_bar_sync(2, ATHREADS);                                         // Y(-1)

#pragma unroll 1
    for (int kb = 0; kb < NT; ++kb) {
      if (s_info != 0) break;
      double* Lk = Lp + 8 * off(kb) * LDP;                              // panel of tile column kb: rows 0..7 = diagonal tile
      // (3) panel solve, one thread per row: rows below the diagonal tile (L21 = A21 L11^-T), the right-hand side block
      // (z_kb = L11^-1 y'_kb) and the eight unit rows e_c (row c of L11^-T = column c of L11^-1)
      const int nrows = NP - 8 * (kb + 1);
      const int rr = utid;
      double x[8];
      const bool lrow = rr < nrows, yrow = rr == nrows, irow = rr > nrows && rr <= nrows + 8;
      if (lrow || yrow || irow) {
        double rv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) rv[c] = rinv_s[c];
        if (lrow) {
          const double2* pr = reinterpret_cast<const double2*>(&Lk[(8 + rr) * LDP]);
#pragma unroll
          for (int c = 0; c < 4; ++c) { const double2 v = pr[c]; x[2 * c] = v.x; x[2 * c + 1] = v.y; }
        } else if (yrow) {
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = ys[8 * kb + c];
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = (c == rr - nrows - 1) ? 1.0 : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          x[c] *= rv[c];
#pragma unroll
          for (int c2 = c + 1; c2 < 8; ++c2) x[c2] = fma(-Lk[c2 * LDP + c], x[c], x[c2]);
        }
        if (lrow) {
          double2* pw = reinterpret_cast<double2*>(&Lk[(8 + rr) * LDP]);
#pragma unroll
          for (int c = 0; c < 4; ++c) pw[c] = make_double2(x[2 * c], x[2 * c + 1]);
        } else if (yrow) {
#pragma unroll
          for (int c = 0; c < 8; ++c) zv[8 * kb + c] = x[c];
        }
      }
      k3la_bar_sync(3, UTHREADS);                                       // Z(kb): panel kb complete, L11 no longer read
      if (irow) {                                                       // Linv(kb,kb) replaces the diagonal tile of L
        const int c = rr - nrows - 1;
#pragma unroll
        for (int j = 0; j < 8; ++j) Lk[j * LDP + c] = x[j];
      }
      if (kb + 1 == NT) break;
      // (4) trailing update (see k3_nlml_la.cuh): tiles of column kb+1 first, spilled to their own panel
      const int off1 = off(kb + 1), off2 = off(kb + 2);
      const int s_lo = (off1 - uw + UW - 1) / UW;
      bool arrived = false;
      auto tile_op = [&](auto sc) -> bool {
        constexpr int s = decltype(sc)::value;
        const int tau = s * UW + uw;
        if (tau >= TILES) return false;
        if (!arrived && tau >= off2) {
          k3la_bar_arrive(1, ATHREADS);                                 // X(kb)
          arrived = true;
        }
        const int I = tI[tau], J = tJ[tau];
        const double* LJ = Lk + (8 * (J - kb) + g) * LDP;
        const double* LI = Lk + (8 * (I - kb) + g) * LDP;
        const double b0 = -LJ[t], b1 = -LJ[4 + t];
        dmma884(C0[s], C1[s], LI[t], b0);
        dmma884(C0[s], C1[s], LI[4 + t], b1);
        if (I == J) {
          double part = fma(b0, zv[8 * kb + t], b1 * zv[8 * kb + 4 + t]);
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          if (t == 0) ys[8 * J + g] += part;
        }
        if (tau < off2) *reinterpret_cast<double2*>(&Lp[(8 * tau + g) * LDP + 2 * t]) = make_double2(C0[s], C1[s]);
        return true;
      };
#define K2LA_CASE(S) case S: if constexpr (S < SLOTS) { if (!tile_op(std::integral_constant<int, S>{})) goto k2la_done; } [[fallthrough]];
      switch (s_lo) {
        K2LA_CASE(0) K2LA_CASE(1) K2LA_CASE(2) K2LA_CASE(3) K2LA_CASE(4) K2LA_CASE(5) K2LA_CASE(6) K2LA_CASE(7)
        K2LA_CASE(8) K2LA_CASE(9) K2LA_CASE(10) K2LA_CASE(11) K2LA_CASE(12) K2LA_CASE(13) K2LA_CASE(14) K2LA_CASE(15)
        K2LA_CASE(16) K2LA_CASE(17) K2LA_CASE(18) K2LA_CASE(19) K2LA_CASE(20) K2LA_CASE(21) K2LA_CASE(22) K2LA_CASE(23)
        default: break;
      }
#undef K2LA_CASE
    k2la_done:
      if (!arrived) k3la_bar_arrive(1, ATHREADS);
      k3la_bar_sync(2, ATHREADS);                                       // Y(kb)
    }
  }
  __syncthreads();

  double* out = invK + b * int64_t(n) * n;
  if (s_info != 0) {
    for (int e = tid; e < n * n; e += THREADS) out[e] = CUDART_NAN;
    for (int i = tid; i < n; i += THREADS) alpha[b * n + i] = CUDART_NAN;
    if (tid == 0) { status[b] = s_info; if (logdet) logdet[b] = CUDART_NAN; }
    return;
  }

  // ============================================ phase 2: X = L^-1 in place ============================================
#pragma unroll 1
  for (int I = 1; I < NT; ++I) {
    const double* Dt = Lp + 8 * off(I) * LDP;                           // Linv(I,I)
    double R0[2], R1[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int J = w + W * q;
      R0[q] = R1[q] = 0.0;
      if (J < I) {
        double c0 = 0.0, c1 = 0.0;
        for (int K = J; K < I; ++K) {
          const double* Lt = Lp + 8 * (off(K) + I - K) * LDP;           // L(I,K)
          const double* Xt = Lp + 8 * (off(J) + K - J) * LDP;           // X(K,J)  (K = J: Linv(J,J))
          dmma884(c0, c1, Lt[g * LDP + t], Xt[t * LDP + g]);
          dmma884(c0, c1, Lt[g * LDP + 4 + t], Xt[(4 + t) * LDP + g]);
        }
        // accumulator -> B operand: B[k = t][n = g] of k-step s is acc[4s + t][g], held by lane 4(4s + t) + (g >> 1)
        const int src0 = 4 * t + (g >> 1), src1 = 4 * (4 + t) + (g >> 1);
        const double u0 = __shfl_sync(0xffffffffu, c0, src0), u1 = __shfl_sync(0xffffffffu, c1, src0);
        const double v0 = __shfl_sync(0xffffffffu, c0, src1), v1 = __shfl_sync(0xffffffffu, c1, src1);
        const double bq0 = (g & 1) ? u1 : u0, bq1 = (g & 1) ? v1 : v0;
        double r0 = 0.0, r1 = 0.0;
        dmma884(r0, r1, Dt[g * LDP + t], bq0);
        dmma884(r0, r1, Dt[g * LDP + 4 + t], bq1);
        R0[q] = -r0; R1[q] = -r1;
      }
    }
    __syncthreads();                                                    // every L(I,.) of this block row has been read
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int J = w + W * q;
      if (J < I) *reinterpret_cast<double2*>(&Lp[(8 * (off(J) + I - J) + g) * LDP + 2 * t]) = make_double2(R0[q], R1[q]);
    }
    __syncthreads();
  }

  // ============================================ phase 3: invK = X'X ============================================
  const bool even_n = (n & 1) == 0;
#pragma unroll 1
  for (int tau = w; tau < TILES; tau += W) {
    const int I = tI[tau], J = tJ[tau];
    double c0 = 0.0, c1 = 0.0;
    for (int K = I; K < NT; ++K) {
      const double* XI = Lp + 8 * (off(I) + K - I) * LDP;               // X(K,I)
      const double* XJ = Lp + 8 * (off(J) + K - J) * LDP;               // X(K,J)
      dmma884(c0, c1, XI[t * LDP + g], XJ[t * LDP + g]);
      dmma884(c0, c1, XI[(4 + t) * LDP + g], XJ[(4 + t) * LDP + g]);
    }
    const int i = 8 * I + g, j = 8 * J + 2 * t;
    if (i < n) {
      if (I != J) {
        if (j + 1 < n && even_n) *reinterpret_cast<double2*>(&out[int64_t(i) * n + j]) = make_double2(c0, c1);
        else {
          if (j < n) out[int64_t(i) * n + j] = c0;
          if (j + 1 < n) out[int64_t(i) * n + j + 1] = c1;
        }
        if (j < n) out[int64_t(j) * n + i] = c0;
        if (j + 1 < n) out[int64_t(j + 1) * n + i] = c1;
      } else {                                                          // diagonal tile: lower half, mirrored
        if (j <= i && j < n) { out[int64_t(i) * n + j] = c0; out[int64_t(j) * n + i] = c0; }
        if (j + 1 <= i && j + 1 < n) { out[int64_t(i) * n + j + 1] = c1; out[int64_t(j + 1) * n + i] = c1; }
      }
    }
  }

  // ============================================ phase 4: alpha = X'z, log det ============================================
  if (tid < n) {
    const int i = tid, Ib = i >> 3, c = i & 7;
    double a0 = 0.0, a1 = 0.0;
    for (int K = Ib; K < NT; ++K) {
      const double* Xt = Lp + 8 * (off(Ib) + K - Ib) * LDP + c;
      const double* zk = zv + 8 * K;
#pragma unroll
      for (int r = 0; r < 8; r += 2) {
        a0 = fma(Xt[r * LDP], zk[r], a0);
        a1 = fma(Xt[(r + 1) * LDP], zk[r + 1], a1);
      }
    }
    alpha[b * n + i] = a0 + a1;
  }
  if (tid == 0) { status[b] = 0; if (logdet) logdet[b] = misc[0]; }
}

template <int NT, int DT>
int launch_k2la_d(int64_t B, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* invK,
                  double* alpha, double* logdet, int32_t* status, cudaStream_t s) {
  const size_t smem = k2la_smem_bytes<NT>(d);
  static PerDeviceOnce once;
  if (once.first()) {
    if (k2la_smem_bytes<NT>(K3R_MAXD) > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k2_factor_la_kernel<NT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           int(k2la_smem_bytes<NT>(K3R_MAXD)));
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    cudaFuncSetAttribute(k2_factor_la_kernel<NT, DT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  k2_factor_la_kernel<NT, DT><<<(unsigned)B, 32 * K3LA<NT>::W, smem, s>>>(n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

template <int NT>
int launch_k2la(int64_t B, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* invK,
                double* alpha, double* logdet, int32_t* status, cudaStream_t s) {
  switch (d) {
    case 2: return launch_k2la_d<NT, 2>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
    case 3: return launch_k2la_d<NT, 3>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
  }
  return launch_k2la_d<NT, 0>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
}

// 32 < n <= 128 (NT = 6 .. 16, even); smaller and larger n stay on the shared-memory kernel of k23_chol.cuh.
int launch_factor_la(int64_t B, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter, double* invK,
                     double* alpha, double* logdet, int32_t* status, cudaStream_t s) {
  if (n > K3R_MAXN || d > K3R_MAXD || n <= 32) return GPRTO_EUNSUPPORTED;
  const int nt = (n + 15) / 16 * 2;
  switch (nt) {
    case 6: return launch_k2la<6>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
    case 8: return launch_k2la<8>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
    case 10: return launch_k2la<10>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
    case 12: return launch_k2la<12>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
    case 14: return launch_k2la<14>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
    case 16: return launch_k2la<16>(B, n, d, Xn, Yn, theta, jitter, invK, alpha, logdet, status, s);
  }
  return GPRTO_EUNSUPPORTED;
}

}  // namespace gprto
