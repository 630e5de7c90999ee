// K3, service-scheduler variant: the look-ahead Cholesky of k3_nlml_la.cuh re-organised around one measured fact
// (tools/fp64_latency.cu, profiles/r02_fp64_latency.json): the FP64 pipe of a B200 SM is PER SCHEDULER (4 per SM), and a
// dependent DFMA / shuffle chain that shares its scheduler with a warp streaming DMMA.8x8x4 runs 4-5x slower
// (DFMA 8.2 -> 40 cycles, one column step of the diagonal-tile elimination 56 -> 241 cycles), while DMMA traffic on the
// OTHER three schedulers does not disturb it at all.  In the look-ahead kernel the pivot chain (65 % of the run time)
// shares its scheduler with an update warp of its own CTA and with two warps of the co-resident CTA.
//
// Here the two warps of a CTA that live on scheduler 3 (warp ids 3 and 7: a warp's scheduler is its id mod 4) are SERVICE
// warps and own no matrix tiles, so that scheduler never sees a DMMA - from either of the two CTAs of the SM:
//   warp 7  DIAG : factorises the 8x8 diagonal tile of column kb (scaled division-free elimination, k3_diag_tile_lean)
//   warp 3  ROWS : panel solve L21 = A21 L11^-T for all rows of column kb, four rows per lane for ILP, plus the right-hand
//                  side block z_kb = L11^-1 y'_kb and the running |z|^2
//   warps 0-2, 4-6 : own the NT(NT+1)/2 tiles round-robin (column-major tile order), apply the trailing update with DMMA;
//                  the tiles of column kb+1 first (4a), then hand it to DIAG and update the rest (4b) while DIAG and ROWS
//                  work on column kb+1.
// Producer/consumer named barriers: X (update -> DIAG: column spilled), D (DIAG -> ROWS: L11 and 1/diag ready),
// R (ROWS -> update: panel solved; also the only barrier the update warps synchronise on among themselves).
#pragma once
#include "gprto_launch.h"
#include "gprto_math.cuh"
#include "k3_nlml_la.cuh"

namespace gprto {

template <int NT>
struct K3SV {
  static constexpr int W = 8;
  static constexpr int UW = 6;
  static constexpr int TILES = NT * (NT + 1) / 2;
  static constexpr int SLOTS = (TILES + UW - 1) / UW;
  static constexpr int NP = 8 * NT;
  static constexpr int THREADS = 32 * W;
  static constexpr int RPL = (NP + 31) / 32;       // panel rows per lane of the ROWS warp (incl. the right-hand side row)
};

template <int NT>
constexpr size_t k3sv_smem_bytes(int d) {
  return sizeof(double) * (2 * size_t(8 * NT) * K3R_LDP + size_t(8 * NT) * d + 8 * NT + 16 + 16 + 2 * 16) + 2 * K3SV<NT>::TILES + 16;
}

template <int NT, int DT>
__global__ void __launch_bounds__(256, (NT <= 8 ? 3 : 2))
k3_nlml_sv_kernel(int S, int n, int d_rt, const double* __restrict__ Xn, const double* __restrict__ Yn,
                  const double* __restrict__ theta_all, double jitter, double* __restrict__ nll,
                  int32_t* __restrict__ status, const int32_t* __restrict__ gidx) {
  using C = K3SV<NT>;
  constexpr int UW = C::UW, TILES = C::TILES, SLOTS = C::SLOTS, NP = C::NP, LDP = K3R_LDP, THREADS = C::THREADS, RPL = C::RPL;
  constexpr int UTHREADS = 32 * UW;
  constexpr int DIAG_W = 7, ROWS_W = 3;
  constexpr int BAR_X = 1, BAR_D = 2, BAR_R = 3;
  const int d = DT > 0 ? DT : d_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* P = reinterpret_cast<double*>(smem_raw);   // [2][NP][LDP]
  double* Xs = P + 2 * NP * LDP;                     // [NP][d]
  double* ys = Xs + NP * d;                          // [NP]
  double* nhw = ys + NP;                             // [16]
  double* rinv2 = nhw + 16;                          // [2][8]
  double* zs2 = rinv2 + 16;                          // [2][16]: z block (8), double-buffered
  unsigned char* tI = reinterpret_cast<unsigned char*>(zs2 + 32);   // [TILES]
  unsigned char* tJ = tI + TILES;                                   // [TILES]
  __shared__ int s_info;
  __shared__ int s_smsp[8];
  __shared__ double s_out[2];                        // [0] |z|^2, [1] log det

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t bs = blockIdx.x;
  const int64_t b = gidx ? gidx[bs] : bs / S;
  const double* theta = theta_all + bs * (d + 2);

  for (int e = tid; e < NP * d; e += THREADS) Xs[e] = (e / d) < n ? Xn[b * int64_t(n) * d + e] : 0.0;
  for (int e = tid; e < NP; e += THREADS) ys[e] = e < n ? Yn[b * n + e] : 0.0;
  if (tid < d) nhw[tid] = -0.5 * exp(-2.0 * theta[tid]);
  if (tid == 0) s_info = 0;
  for (int e = tid; e < TILES; e += THREADS) {
    int J = 0;
    while ((J + 1) * NT - (J + 1) * J / 2 <= e) ++J;
    tJ[e] = (unsigned char)J;
    tI[e] = (unsigned char)(J + e - (J * NT - J * (J - 1) / 2));
  }
  // Roles follow the HARDWARE warp slot, not the logical warp id: a warp's scheduler is %warpid mod 4, and the second CTA
  // of an SM gets its slots rotated (measured, tools/warpid_probe.cu: logical warps 0..7 -> slots 9,10,11,8,13,14,15,12), so
  // "warps 3 and 7" of both CTAs do NOT share a scheduler.  The two warps of this CTA on scheduler 3 become the service
  // warps; if the slot layout is not the expected two-per-scheduler one, fall back to logical warps 3 and 7.
  {
    unsigned hw;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
    if (lane == 0) s_smsp[w] = int(hw & 3u);
  }
  __syncthreads();
  int role_w;                                        // DIAG_W, ROWS_W, or the update-warp index 0..5 mapped onto the other ids
  {
    int nsv = 0, below_sv = 0, below_up = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool sv = s_smsp[i] == 3;
      nsv += sv;
      if (i < w) { below_sv += sv; below_up += !sv; }
    }
    const bool mine = s_smsp[w] == 3;
    if (nsv == 2) role_w = mine ? (below_sv == 1 ? DIAG_W : ROWS_W) : (below_up < ROWS_W ? below_up : below_up + 1);
    else role_w = w;
  }

  if (role_w == DIAG_W) {
    // ================================ DIAG warp ================================
    double pm = 1.0;
    int pe = 0, info = 0;
    K3LA_T0
#pragma unroll 1
    for (int kb = 0; kb < NT; ++kb) {
      k3la_bar_sync(BAR_X, UTHREADS + 32);                              // X(kb): diagonal tile of column kb is in the panel
      K3LA_MARK(0)
      k3_diag_tile_lean(P + (kb & 1) * NP * LDP + 8 * kb * LDP, LDP, rinv2 + 8 * (kb & 1), lane, 8 * kb, pm, pe, info);
      if (lane == 0 && info) s_info = info;
      k3la_bar_arrive(BAR_D, 64);                                       // D(kb)
      K3LA_MARK(1)
      if (info) break;
    }
#ifdef K3_TIMING
    if (lane == 0 && bs == 20000) printf("K3SV DIAG: wait X %lld diag %lld\n", tk[0], tk[1]);
#endif
    if (lane == 0) s_out[1] = log(pm) + double(pe) * 0.693147180559945309417;
  } else if (role_w == ROWS_W) {
    // ================================ ROWS warp ================================
    double quad = 0.0;
    K3LA_T0
#pragma unroll 1
    for (int kb = 0; kb < NT; ++kb) {
      k3la_bar_sync(BAR_D, 64);                                         // D(kb)
      K3LA_MARK(0)
      if (s_info != 0) { k3la_bar_arrive(BAR_R, UTHREADS + 32); break; }
      double* buf = P + (kb & 1) * NP * LDP;
      const double* rinv = rinv2 + 8 * (kb & 1);
      const int nrows = NP - 8 * (kb + 1);                              // rows below the diagonal tile; row index nrows = right-hand side
      double x[RPL][8];
      bool act[RPL], yrow[RPL];
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int rr = lane + 32 * q;
        act[q] = rr <= nrows;
        yrow[q] = rr == nrows;
        if (act[q] && !yrow[q]) {
          const double2* pr = reinterpret_cast<const double2*>(&buf[(8 * (kb + 1) + rr) * LDP]);
#pragma unroll
          for (int c = 0; c < 4; ++c) { const double2 v = pr[c]; x[q][2 * c] = v.x; x[q][2 * c + 1] = v.y; }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) x[q][c] = yrow[q] ? ys[8 * kb + c] : 0.0;
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double rv = rinv[c];
#pragma unroll
        for (int q = 0; q < RPL; ++q) x[q][c] *= rv;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) {
          const double l = buf[(8 * kb + c2) * LDP + c];                // L11[c2][c], broadcast
#pragma unroll
          for (int q = 0; q < RPL; ++q) x[q][c2] = fma(-l, x[q][c], x[q][c2]);
        }
      }
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int rr = lane + 32 * q;
        if (act[q] && !yrow[q]) {
          double2* pw = reinterpret_cast<double2*>(&buf[(8 * (kb + 1) + rr) * LDP]);
#pragma unroll
          for (int c = 0; c < 4; ++c) pw[c] = make_double2(x[q][2 * c], x[q][2 * c + 1]);
        } else if (yrow[q]) {
          double* zw = zs2 + (kb & 1) * 16;
#pragma unroll
          for (int c = 0; c < 8; ++c) { zw[c] = x[q][c]; quad = fma(x[q][c], x[q][c], quad); }
        }
      }
      k3la_bar_arrive(BAR_R, UTHREADS + 32);                            // R(kb)
      K3LA_MARK(1)
    }
#ifdef K3_TIMING
    if (lane == 0 && bs == 20000) printf("K3SV ROWS: wait D %lld rows %lld\n", tk[0], tk[1]);
#endif
    // exactly one lane per step held the right-hand side row
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) quad += __shfl_xor_sync(0xffffffffu, quad, off);
    if (lane == 0) s_out[0] = quad;
  } else {
    // ================================ update warps ================================
    const int uw = role_w < ROWS_W ? role_w : role_w - 1;
    const double c0 = 2.0 * theta[d];
    const double diag_add = exp(2.0 * theta[d + 1]) + jitter;
    double C0[SLOTS], C1[SLOTS];
    unsigned ij[(SLOTS + 3) / 4];                   // (I, J) of this warp's slots, one byte each: no table load on the update path
#pragma unroll
    for (int q = 0; q < (SLOTS + 3) / 4; ++q) ij[q] = 0u;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int tau = s * UW + uw;
      const int I = tau < TILES ? tI[tau] : 0, J = tau < TILES ? tJ[tau] : 0;
      ij[s >> 2] |= unsigned(I | (J << 4)) << (8 * (s & 3));
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
      if (tau < NT) *reinterpret_cast<double2*>(&P[(8 * tI[tau] + g) * LDP + 2 * t]) = make_double2(C0[s], C1[s]);
    }
    k3la_bar_arrive(BAR_X, UTHREADS + 32);                              // X(0)
    K3LA_T0

#pragma unroll 1
    for (int kb = 0; kb < NT; ++kb) {
      k3la_bar_sync(BAR_R, UTHREADS + 32);                              // R(kb): panel kb solved; all update warps are past step kb-1
      K3LA_MARK(0)
      if (s_info != 0 || kb + 1 == NT) break;
      double* buf = P + (kb & 1) * NP * LDP;
      double* nbuf = P + ((kb + 1) & 1) * NP * LDP;
      const double* zs = zs2 + (kb & 1) * 16;
      const int off1 = (kb + 1) * NT - (kb + 1) * kb / 2;                // first tile of column kb+1
      const int off2 = (kb + 2) * NT - (kb + 2) * (kb + 1) / 2;          // first tile of column kb+2
      const int s_lo = (off1 - uw + UW - 1) / UW;                        // first active slot of this warp
      const int s_mid = (off2 - uw + UW - 1) / UW;                       // first slot beyond column kb+1
      // one tile: C(I,J) -= L(I,kb) L(J,kb)'; the diagonal tile of row block J also carries y'_J -= L(J,kb) z_kb
      auto rhs_update = [&](int J, double b0, double b1) {
        double part = fma(b0, zs[t], b1 * zs[4 + t]);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (t == 0) ys[8 * J + g] += part;
      };
      // (4a) the tiles of column kb+1 (at most ceil((NT-1)/UW) per warp), spilled to the other panel buffer, then X(kb+1)
      auto single_op = [&](auto sc) {
        constexpr int s = decltype(sc)::value;
        const int I = (ij[s >> 2] >> (8 * (s & 3))) & 15, J = (ij[s >> 2] >> (8 * (s & 3) + 4)) & 15;
        const double* LI = buf + (8 * I + g) * LDP;
        const double* LJ = buf + (8 * J + g) * LDP;
        const double b0 = -LJ[t], b1 = -LJ[4 + t];
        dmma884(C0[s], C1[s], LI[t], b0);
        dmma884(C0[s], C1[s], LI[4 + t], b1);
        if (I == J) rhs_update(J, b0, b1);
        *reinterpret_cast<double2*>(&nbuf[(8 * I + g) * LDP + 2 * t]) = make_double2(C0[s], C1[s]);
      };
#define K3SV_ONE(S) case S: if constexpr (S < SLOTS) single_op(std::integral_constant<int, S>{}); break;
      for (int s4 = s_lo; s4 < s_mid; ++s4) {
        switch (s4) {
          K3SV_ONE(0) K3SV_ONE(1) K3SV_ONE(2) K3SV_ONE(3) K3SV_ONE(4) K3SV_ONE(5) K3SV_ONE(6) K3SV_ONE(7)
          K3SV_ONE(8) K3SV_ONE(9) K3SV_ONE(10) K3SV_ONE(11) K3SV_ONE(12) K3SV_ONE(13) K3SV_ONE(14) K3SV_ONE(15)
          K3SV_ONE(16) K3SV_ONE(17) K3SV_ONE(18) K3SV_ONE(19) K3SV_ONE(20) K3SV_ONE(21) K3SV_ONE(22) K3SV_ONE(23)
          default: break;
        }
      }
#undef K3SV_ONE
      K3LA_MARK(1)
      k3la_bar_arrive(BAR_X, UTHREADS + 32);                            // X(kb+1): column kb+1 handed to the DIAG warp
      // (4b) the rest, TWO tiles at a time: all eight fragment loads first, then the DMMAs of the two tiles interleaved, so
      // that neither the shared-memory latency nor the dependent DMMA pair of one tile stalls the warp (the update warps
      // are latency-bound, not pipe-bound: one tile at a time costs ~270 cycles for 2 x 16 cycles of FP64-pipe work)
      auto pair_op = [&](auto sc) -> bool {
        constexpr int s = decltype(sc)::value;
        constexpr bool has2 = (s + 1 < SLOTS);
        if (s * UW + uw >= TILES) return false;
        const int I0 = (ij[s >> 2] >> (8 * (s & 3))) & 15, J0 = (ij[s >> 2] >> (8 * (s & 3) + 4)) & 15;
        const double* LI0 = buf + (8 * I0 + g) * LDP;
        const double* LJ0 = buf + (8 * J0 + g) * LDP;
        const double a00 = LI0[t], a01 = LI0[4 + t], b00 = -LJ0[t], b01 = -LJ0[4 + t];
        if constexpr (has2) {
          constexpr int s1 = has2 ? s + 1 : s;
          if (s1 * UW + uw < TILES) {
            const int I1 = (ij[s1 >> 2] >> (8 * (s1 & 3))) & 15, J1 = (ij[s1 >> 2] >> (8 * (s1 & 3) + 4)) & 15;
            const double* LI1 = buf + (8 * I1 + g) * LDP;
            const double* LJ1 = buf + (8 * J1 + g) * LDP;
            const double a10 = LI1[t], a11 = LI1[4 + t], b10 = -LJ1[t], b11 = -LJ1[4 + t];
            dmma884(C0[s], C1[s], a00, b00);
            dmma884(C0[s1], C1[s1], a10, b10);
            dmma884(C0[s], C1[s], a01, b01);
            dmma884(C0[s1], C1[s1], a11, b11);
            if (I0 == J0) rhs_update(J0, b00, b01);
            if (I1 == J1) rhs_update(J1, b10, b11);
            return true;
          }
        }
        dmma884(C0[s], C1[s], a00, b00);
        dmma884(C0[s], C1[s], a01, b01);
        if (I0 == J0) rhs_update(J0, b00, b01);
        return false;
      };
#define K3SV_PAIR(S) case S: if constexpr (S < SLOTS) { if (!pair_op(std::integral_constant<int, S>{})) goto k3sv_done; } [[fallthrough]];
#ifndef K3SV_PAIRING
      auto rest_op = [&](auto sc) -> bool {
        constexpr int s = decltype(sc)::value;
        if (s * UW + uw >= TILES) return false;
        const int I = (ij[s >> 2] >> (8 * (s & 3))) & 15, J = (ij[s >> 2] >> (8 * (s & 3) + 4)) & 15;
        const double* LI = buf + (8 * I + g) * LDP;
        const double* LJ = buf + (8 * J + g) * LDP;
        const double b0 = -LJ[t], b1 = -LJ[4 + t];
        dmma884(C0[s], C1[s], LI[t], b0);
        dmma884(C0[s], C1[s], LI[4 + t], b1);
        if (I == J) rhs_update(J, b0, b1);
        return true;
      };
#define K3SV_REST(S) case S: if constexpr (S < SLOTS) { if (!rest_op(std::integral_constant<int, S>{})) goto k3sv_done; } [[fallthrough]];
      switch (s_mid) {
        K3SV_REST(0) K3SV_REST(1) K3SV_REST(2) K3SV_REST(3) K3SV_REST(4) K3SV_REST(5) K3SV_REST(6) K3SV_REST(7)
        K3SV_REST(8) K3SV_REST(9) K3SV_REST(10) K3SV_REST(11) K3SV_REST(12) K3SV_REST(13) K3SV_REST(14) K3SV_REST(15)
        K3SV_REST(16) K3SV_REST(17) K3SV_REST(18) K3SV_REST(19) K3SV_REST(20) K3SV_REST(21) K3SV_REST(22) K3SV_REST(23)
        default: break;
      }
#undef K3SV_REST
      (void)pair_op;
#else
      if (s_mid & 1) {
        switch (s_mid) {
          K3SV_PAIR(1) K3SV_PAIR(3) K3SV_PAIR(5) K3SV_PAIR(7) K3SV_PAIR(9) K3SV_PAIR(11) K3SV_PAIR(13) K3SV_PAIR(15)
          K3SV_PAIR(17) K3SV_PAIR(19) K3SV_PAIR(21) K3SV_PAIR(23)
          default: break;
        }
      } else {
        switch (s_mid) {
          K3SV_PAIR(0) K3SV_PAIR(2) K3SV_PAIR(4) K3SV_PAIR(6) K3SV_PAIR(8) K3SV_PAIR(10) K3SV_PAIR(12) K3SV_PAIR(14)
          K3SV_PAIR(16) K3SV_PAIR(18) K3SV_PAIR(20) K3SV_PAIR(22)
          default: break;
        }
      }
#endif
#undef K3SV_PAIR
    k3sv_done:;
      K3LA_MARK(2)
    }
#ifdef K3_TIMING
    if (lane == 0 && bs == 20000) printf("K3SV update warp %d: wait R %lld 4a %lld 4b %lld\n", w, tk[0], tk[1], tk[2]);
#endif
  }
  __syncthreads();
  if (tid == 0) {
    const int inf = s_info;
    status[bs] = inf;
    nll[bs] = inf == 0 ? s_out[0] + s_out[1] : CUDART_NAN;
  }
}

template <int NT, int DT>
int launch_k3sv_d(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                  double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  const size_t smem = k3sv_smem_bytes<NT>(d);
  static PerDeviceOnce once;
  if (once.first()) {
    if (k3sv_smem_bytes<NT>(K3R_MAXD) > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k3_nlml_sv_kernel<NT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           int(k3sv_smem_bytes<NT>(K3R_MAXD)));
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    cudaFuncSetAttribute(k3_nlml_sv_kernel<NT, DT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  k3_nlml_sv_kernel<NT, DT><<<(unsigned)(B * S), 256, smem, s>>>(S, n, d, Xn, Yn, theta, jitter, nll, status, gidx);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

template <int NT>
int launch_k3sv(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  switch (d) {
    case 2: return launch_k3sv_d<NT, 2>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 3: return launch_k3sv_d<NT, 3>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return launch_k3sv_d<NT, 0>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
}

#ifdef GPRTO_K3_TU
int launch_nlml_sv(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                   double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx) {
  if (n > K3R_MAXN || d > K3R_MAXD || n <= 32) return GPRTO_EUNSUPPORTED;
  const int nt = (n + 15) / 16 * 2;
  switch (nt) {
    case 6: return launch_k3sv<6>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 8: return launch_k3sv<8>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 10: return launch_k3sv<10>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 12: return launch_k3sv<12>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 14: return launch_k3sv<14>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 16: return launch_k3sv<16>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return GPRTO_EUNSUPPORTED;
}
#endif

}  // namespace gprto
