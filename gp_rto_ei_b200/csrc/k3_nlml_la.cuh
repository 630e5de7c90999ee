// K3, look-ahead variant (n > 48): same register-resident tiled Cholesky as k3_nlml_reg.cuh, restructured so that the
// n-long pivot chain no longer stalls the tensor-core work.
//
//   * warp 0 is a PANEL warp: it owns no matrix tiles; it factorises the 8x8 diagonal tile of column kb+1 (scaled
//     division-free elimination, see k3_nlml_reg.cuh) WHILE the update warps are still applying column kb to the
//     rest of the trailing matrix (look-ahead of depth one).
//   * the other NT/2-1 warps own the NT(NT+1)/2 tiles round-robin in column-major tile order, so both the "next
//     column first" part (4a) and the remainder (4b) of every update are balanced across warps; a tile's (I, J) comes
//     from a 2-byte shared-memory table.
//   * per block step: update warps  4a: tiles of column kb+1 -> spill to the other panel buffer -> bar.arrive(X)
//                                   4b: remaining tiles                      | panel warp: bar.sync(X); diag(kb+1)
//                     bar.sync(Y) all; update warps: panel rows of column kb+1; bar.sync(Z) among update warps.
//     z and the panel are double-buffered; the right-hand side update rides on the diagonal tiles as before.
#pragma once
#include "gprto_launch.h"
#include "gprto_math.cuh"
#include "k3_nlml_reg.cuh"
#include <type_traits>
#ifdef K3_TRACE
#include <cstdlib>
#endif

#ifdef K3_TIMING
#define K3LA_T0 long long tk[6] = {0, 0, 0, 0, 0, 0}; long long tprev = clock64();
#define K3LA_MARK(i) { const long long tn = clock64(); tk[i] += tn - tprev; tprev = tn; }
#else
#define K3LA_T0
#define K3LA_MARK(i)
#endif

namespace gprto {

// -DK3_TRACE: absolute clock64 time stamps of one CTA (block K3_TRACE_BS), per warp / block step / event, for the
// timeline analysis of tools/k3_trace.py (debug builds only; the accessor below is not part of the ABI)
#ifdef K3_TRACE
#ifndef K3_TRACE_BS
#define K3_TRACE_BS 20000
#endif
__device__ long long g_k3_trace[8 * 32 * 8];
#define K3TR(kb, ev) { if (bs == K3_TRACE_BS && lane == 0) g_k3_trace[((tid >> 5) * 32 + (kb)) * 8 + (ev)] = clock64(); }
#else
#define K3TR(kb, ev)
#endif

// warps per CTA: 1 panel + W-1 update warps.  8 for every size: at NT = 8..14 that is more warps (and fewer tiles, i.e.
// registers, per warp) than the NT/2 of the plain register kernel, which buys a third resident CTA at NT = 8.
template <int NT>
struct K3LA {
  static constexpr int W = 8;
  // (idling the warp that shares the panel warp's scheduler was measured slower: 7.5e6 vs 8.3e6 evals/s at n = 128)
  static constexpr int IDLE = 0;
  static constexpr int UW = W - 1 - IDLE;
  static constexpr int TILES = NT * (NT + 1) / 2;
  static constexpr int SLOTS = (TILES + UW - 1) / UW;
  static constexpr int NP = 8 * NT;
  static constexpr int THREADS = 32 * W;
};

template <int NT>
constexpr size_t k3la_smem_bytes(int d) {
  return sizeof(double) * (2 * size_t(8 * NT) * K3R_LDP + size_t(8 * NT) * d + 8 * NT + 16 + 8 + 2 * 16) + 2 * K3LA<NT>::TILES + 16;
}

__device__ __forceinline__ void k3la_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void k3la_bar_arrive(int id, int n) {
  __threadfence_block();      // producer side of a producer/consumer barrier: make the panel / z writes visible first
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// Scaled, division- and sqrt-free elimination of the 8x8 diagonal tile spread over a warp in accumulator layout
// (lane 4r+q holds a[r][2q], a[r][2q+1]; lanes 0..7 hold the right-hand side block).  Writes L11 back to the panel,
// 1/L_cc to rinv, z = L11^-1 y' to zs[0..7] and |z|^2 to zs[8]; returns (log det contribution on lane 0, info).
__device__ __forceinline__ void k3_diag_tile(double* tile, int ldp, const double* yblk, double* rinv, double* zs, int lane,
                                             int col0, double& logdet_acc, int& info_out) {
  const int r = lane >> 2, q = lane & 3, k8 = lane & 7;
  const double2 v = *reinterpret_cast<const double2*>(&tile[r * ldp + 2 * q]);
  double x0 = v.x, x1 = v.y;
  double yv = yblk[k8];
  double S = 1.0, myS = 1.0, mydp = 1.0;
  int info = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int cs = c >> 1;
    const double colsel = (c & 1) ? x1 : x0;
    const double dpiv = __shfl_sync(0xffffffffu, colsel, 4 * c + cs);
    const double arc = __shfl_sync(0xffffffffu, colsel, 4 * r + cs);
    const double ak0 = __shfl_sync(0xffffffffu, colsel, 4 * (2 * q) + cs);
    const double ak1 = __shfl_sync(0xffffffffu, colsel, 4 * (2 * q + 1) + cs);
    const double ayk = __shfl_sync(0xffffffffu, colsel, 4 * k8 + cs);
    const double yc = __shfl_sync(0xffffffffu, yv, c);
    if (!(dpiv > 0.0) && info == 0) info = col0 + c + 1;
    if (lane == c) { myS = S; mydp = dpiv; }
    const int e = ((__double2hiint(dpiv) >> 20) & 0x7ff) - 1023;
    const double f = __hiloint2double((1023 - e) << 20, 0);
    const double dt = dpiv * f;
    const double ac = arc * f, ycf = yc * f;
    const double n0 = fma(-ac, ak0, dt * x0), n1 = fma(-ac, ak1, dt * x1);
    const double ny = fma(-ycf, ayk, dt * yv);
    x0 = (2 * q > c) ? n0 : x0;
    x1 = (2 * q + 1 > c) ? n1 : x1;
    yv = (k8 > c) ? ny : yv;
    S *= dt;
  }
  const double tmine = lane < 8 ? rsqrt_fast(mydp * myS) : 1.0;
  const double t0 = __shfl_sync(0xffffffffu, tmine, 2 * q), t1 = __shfl_sync(0xffffffffu, tmine, 2 * q + 1);
  *reinterpret_cast<double2*>(&tile[r * ldp + 2 * q]) = make_double2(x0 * t0, x1 * t1);
  double zq = 0.0, lg = 0.0;
  if (lane < 8) {
    const double z = yv * tmine;
    zs[lane] = z;
    zq = z * z;
    rinv[lane] = tmine * myS;
    lg = 2.0 * log_det(mydp * tmine);
  }
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
    lg += __shfl_xor_sync(0xffffffffu, lg, off);
    zq += __shfl_xor_sync(0xffffffffu, zq, off);
  }
  if (lane == 0) zs[8] = zq;
  logdet_acc += lg;
  info_out = info;
}

// Lean variant for the look-ahead kernel: the panel warp is the critical path, so everything that can leave its
// chain does: the right-hand side block is solved by the update warps with the panel rows, and instead of 8 logarithms
// per tile the pivots are multiplied up (lane 0 keeps a running mantissa / exponent pair; one log per matrix at the end).
__device__ __forceinline__ void k3_diag_tile_lean(double* tile, int ldp, double* rinv, int lane, int col0, double& pm, int& pe,
                                                  int& info_out) {
  const int r = lane >> 2, q = lane & 3;
  const double2 v = *reinterpret_cast<const double2*>(&tile[r * ldp + 2 * q]);
  double x0 = v.x, x1 = v.y;
  double S = 1.0, myS = 1.0, mydp = 1.0;
  int info = 0, esum = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int cs = c >> 1;
    const double colsel = (c & 1) ? x1 : x0;
    const double dpiv = __shfl_sync(0xffffffffu, colsel, 4 * c + cs);
    const double arc = __shfl_sync(0xffffffffu, colsel, 4 * r + cs);
    const double ak0 = __shfl_sync(0xffffffffu, colsel, 4 * (2 * q) + cs);
    const double ak1 = __shfl_sync(0xffffffffu, colsel, 4 * (2 * q + 1) + cs);
    if (!(dpiv > 0.0) && info == 0) info = col0 + c + 1;
    if (lane == c) { myS = S; mydp = dpiv; }
    const int e = ((__double2hiint(dpiv) >> 20) & 0x7ff) - 1023;
    const double f = __hiloint2double((1023 - e) << 20, 0);
    const double dt = dpiv * f;
    const double ac = arc * f;
    const double n0 = fma(-ac, ak0, dt * x0), n1 = fma(-ac, ak1, dt * x1);
    x0 = (2 * q > c) ? n0 : x0;
    x1 = (2 * q + 1 > c) ? n1 : x1;
    // pivot p_c = d'_c / S_c = dt 2^e / S;  prod_c p_c = 2^(sum e) * prod dt / prod S_c  - tracked below without logs
    esum += e;
    S *= dt;
  }
  const double tmine = lane < 8 ? rsqrt_fast(mydp * myS) : 1.0;
  const double t0 = __shfl_sync(0xffffffffu, tmine, 2 * q), t1 = __shfl_sync(0xffffffffu, tmine, 2 * q + 1);
  *reinterpret_cast<double2*>(&tile[r * ldp + 2 * q]) = make_double2(x0 * t0, x1 * t1);
  if (lane < 8) rinv[lane] = tmine * myS;
  // product of the 8 pivots = prod_c (L_cc)^2, L_cc = d'_c t_c  (each L_cc is O(1)..O(1e3): no overflow over 8 factors)
  double lp = lane < 8 ? mydp * tmine : 1.0;
  lp *= __shfl_xor_sync(0xffffffffu, lp, 1);
  lp *= __shfl_xor_sync(0xffffffffu, lp, 2);
  lp *= __shfl_xor_sync(0xffffffffu, lp, 4);
  // running product on every lane (uniform): mantissa in [1,2), exponent separately
  const double prod = pm * lp * lp;
  const int ex = ((__double2hiint(prod) >> 20) & 0x7ff) - 1023;
  pm = __hiloint2double(__double2hiint(prod) - (ex << 20), __double2loint(prod));
  pe += ex;
  (void)esum;
  info_out = info;
}

template <int NT, int DT>
__global__ void __launch_bounds__(32 * K3LA<NT>::W, (NT <= 8 ? 3 : 2))      // 256 threads: <= 85 / 128 registers
k3_nlml_la_kernel(int S, int n, int d_rt, const double* __restrict__ Xn, const double* __restrict__ Yn,
                  const double* __restrict__ theta_all, double jitter, double* __restrict__ nll,
                  int32_t* __restrict__ status, const int32_t* __restrict__ gidx) {
  using C = K3LA<NT>;
  constexpr int UW = C::UW, TILES = C::TILES, SLOTS = C::SLOTS, NP = C::NP, LDP = K3R_LDP, THREADS = C::THREADS;
  constexpr int UTHREADS = 32 * UW;
  constexpr int ATHREADS = 32 * (UW + 1);      // threads taking part in the X / Y barriers
  constexpr int PANEL_W = C::W - 1, IDLE_W = C::IDLE ? C::W - 5 : -1;
  const int d = DT > 0 ? DT : d_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* P = reinterpret_cast<double*>(smem_raw);   // [2][NP][LDP]
  double* Xs = P + 2 * NP * LDP;                     // [NP][d]
  double* ys = Xs + NP * d;                          // [NP]
  double* nhw = ys + NP;                             // [16]
  double* rinv_s = nhw + 16;                         // [8]
  double* zs2 = rinv_s + 8;                          // [2][16]: z block (8), double-buffered; zs2[15] = running |z|^2
  double* quad_s = zs2 + 15;
  unsigned char* tI = reinterpret_cast<unsigned char*>(zs2 + 32);   // [TILES]
  unsigned char* tJ = tI + TILES;                                   // [TILES]
  __shared__ int s_info;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t bs = blockIdx.x;
  const int64_t b = gidx ? gidx[bs] : bs / S;      // gidx: problem -> data set indirection (indexed entry point)
  const double* theta = theta_all + bs * (d + 2);

  for (int e = tid; e < NP * d; e += THREADS) Xs[e] = (e / d) < n ? Xn[b * int64_t(n) * d + e] : 0.0;
  for (int e = tid; e < NP; e += THREADS) ys[e] = e < n ? Yn[b * n + e] : 0.0;
  if (tid < d) nhw[tid] = -0.5 * exp_det(-2.0 * theta[tid]);
  if (tid == 0) { s_info = 0; *quad_s = 0.0; }
  for (int e = tid; e < TILES; e += THREADS) {      // column-major tile order: tau = off(J) + (I - J), off(J) = J NT - J(J-1)/2
    int J = 0;
    while ((J + 1) * NT - (J + 1) * J / 2 <= e) ++J;
    tJ[e] = (unsigned char)J;
    tI[e] = (unsigned char)(J + e - (J * NT - J * (J - 1) / 2));
  }
  __syncthreads();

  // The panel warp is the LAST warp of the CTA: the scheduler favours higher warp ids among eligible warps, and the
  // pivot chain is the critical path (measured: as warp 0 it ran 2x slower, starved by the update warps).
  if (w == IDLE_W) return;
  if (w == PANEL_W) {
    // ================================ panel warp ================================
    double pm = 1.0;
    int pe = 0, info = 0;
    K3LA_T0
    K3TR(31, 0)
    k3la_bar_sync(1, ATHREADS);                                         // X(-1): column 0 is in panel buffer 0
    K3TR(31, 1)
    k3_diag_tile_lean(P + 0, LDP, rinv_s, lane, 0, pm, pe, info);
    if (lane == 0 && info) s_info = info;
    K3TR(31, 2)
    k3la_bar_sync(2, ATHREADS);                                         // Y(-1)
    K3TR(31, 3)
#pragma unroll 1
    for (int kb = 0; kb + 1 < NT; ++kb) {
      if (s_info != 0) break;
      double* nbuf = P + ((kb + 1) & 1) * NP * LDP;
      K3LA_MARK(0)
      K3TR(kb, 0)
      k3la_bar_sync(1, ATHREADS);                                       // X(kb): column kb+1 updated and spilled
      K3LA_MARK(1)
      K3TR(kb, 1)
      k3_diag_tile_lean(nbuf + 8 * (kb + 1) * LDP, LDP, rinv_s, lane, 8 * (kb + 1), pm, pe, info);
      if (lane == 0 && info) s_info = info;
      K3LA_MARK(2)
      K3TR(kb, 2)
      k3la_bar_sync(2, ATHREADS);                                       // Y(kb)
      K3LA_MARK(3)
      K3TR(kb, 3)
    }
    k3la_bar_sync(4, ATHREADS);                                         // F: the update warps have finished z (quad_s)
#ifdef K3_TIMING
    if (lane == 0 && bs == 20000) printf("K3LA panel warp: wait X %lld diag %lld wait Y %lld other %lld\n", tk[1], tk[2], tk[3], tk[0]);
#endif
    if (lane == 0) {
      const int inf = s_info;
      status[bs] = inf;
      // log det = log(prod pivots) = log(mantissa) + exponent * ln 2
      nll[bs] = inf == 0 ? *quad_s + log_det(pm) + double(pe) * 0.693147180559945309417 : CUDART_NAN;
    }
    return;
  }

  // ================================ update warps ================================
  const int uw = (IDLE_W >= 0 && w > IDLE_W) ? w - 1 : w, utid = 32 * uw + lane;
  const double c0 = 2.0 * theta[d];
  const double diag_add = exp_det(2.0 * theta[d + 1]) + jitter;
  K3TR(31, 0)
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
  // prologue: spill column 0 (tiles tau < NT), let the panel warp factor it, solve the panel rows of column 0
  K3TR(31, 1)
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int tau = s * UW + uw;
    if (tau < NT) *reinterpret_cast<double2*>(&P[(8 * tI[tau] + g) * LDP + 2 * t]) = make_double2(C0[s], C1[s]);
  }
  k3la_bar_arrive(1, ATHREADS);                                         // X(-1)
  k3la_bar_sync(2, ATHREADS);                                           // Y(-1): L11(0), rinv, z_0 ready
  K3TR(31, 2)
  K3LA_T0

#pragma unroll 1
  for (int kb = 0; kb < NT; ++kb) {
    if (s_info != 0) break;
    double* buf = P + (kb & 1) * NP * LDP;
    double* nbuf = P + ((kb + 1) & 1) * NP * LDP;
    const double* zs = zs2 + (kb & 1) * 16;
    K3TR(kb, 0)
    // (3) panel rows of column kb: L21 = A21 L11^-T, one thread per remaining row; the thread after the last row solves
    // the right-hand side block the same way (y is the extra row of the augmented matrix): z_kb = L11^-1 y'_kb
    {
      const int nrows = NP - 8 * (kb + 1);
      for (int rr = utid; rr <= nrows; rr += UTHREADS) {
        const bool yrow = rr == nrows;
        const int row = 8 * (kb + 1) + rr;
        double x[8], rv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) rv[c] = rinv_s[c];
        if (yrow) {
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = ys[8 * kb + c];
        } else {
          const double2* pr = reinterpret_cast<const double2*>(&buf[row * LDP]);
#pragma unroll
          for (int c = 0; c < 4; ++c) { const double2 v = pr[c]; x[2 * c] = v.x; x[2 * c + 1] = v.y; }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          x[c] *= rv[c];
#pragma unroll
          for (int c2 = c + 1; c2 < 8; ++c2) x[c2] = fma(-buf[(8 * kb + c2) * LDP + c], x[c], x[c2]);
        }
        if (yrow) {
          double qq = 0.0;
          double* zw = zs2 + (kb & 1) * 16;
#pragma unroll
          for (int c = 0; c < 8; ++c) { zw[c] = x[c]; qq = fma(x[c], x[c], qq); }
          *quad_s += qq;                                         // single writer per step
        } else {
          double2* pw = reinterpret_cast<double2*>(&buf[row * LDP]);
#pragma unroll
          for (int c = 0; c < 4; ++c) pw[c] = make_double2(x[2 * c], x[2 * c + 1]);
        }
      }
    }
    K3LA_MARK(0)
    K3TR(kb, 1)
    k3la_bar_sync(3, UTHREADS);                                        // Z(kb): panel kb complete (update warps only)
    K3LA_MARK(1)
    K3TR(kb, 2)
    if (kb + 1 == NT) break;
    // (4) trailing update.  This warp's tiles are tau = s UW + uw in column-major order, so the active ones (tau >=
    // off1) are a contiguous range of slots starting at s_lo, the tiles of column kb+1 coming first: jump straight to
    // slot s_lo (no per-slot predicate chain - the warps are instruction-latency bound), update and spill the column
    // kb+1 tiles (4a), arrive at X as soon as the first later tile is reached, then update the rest (4b).
    const int off1 = (kb + 1) * NT - (kb + 1) * kb / 2;                // first tile of column kb+1
    const int off2 = (kb + 2) * NT - (kb + 2) * (kb + 1) / 2;          // first tile of column kb+2
    const int s_lo = (off1 - uw + UW - 1) / UW;
    bool arrived = false;
    auto tile_op = [&](auto sc) -> bool {
      constexpr int s = decltype(sc)::value;
      const int tau = s * UW + uw;
      if (tau >= TILES) return false;
      if (!arrived && tau >= off2) {
        K3LA_MARK(2)
        K3TR(kb, 3)
        k3la_bar_arrive(1, ATHREADS);                                   // X(kb): hand column kb+1 to the panel warp
        arrived = true;
      }
      const int I = tI[tau], J = tJ[tau];
      const double b0 = -buf[(8 * J + g) * LDP + t], b1 = -buf[(8 * J + g) * LDP + 4 + t];
      dmma884(C0[s], C1[s], buf[(8 * I + g) * LDP + t], b0);
      dmma884(C0[s], C1[s], buf[(8 * I + g) * LDP + 4 + t], b1);
      if (I == J) {
        double part = fma(b0, zs[t], b1 * zs[4 + t]);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (t == 0) ys[8 * J + g] += part;
      }
      if (tau < off2) *reinterpret_cast<double2*>(&nbuf[(8 * I + g) * LDP + 2 * t]) = make_double2(C0[s], C1[s]);
      return true;
    };
#define K3LA_CASE(S) case S: if constexpr (S < SLOTS) { if (!tile_op(std::integral_constant<int, S>{})) goto k3la_done; } [[fallthrough]];
    switch (s_lo) {
      K3LA_CASE(0) K3LA_CASE(1) K3LA_CASE(2) K3LA_CASE(3) K3LA_CASE(4) K3LA_CASE(5) K3LA_CASE(6) K3LA_CASE(7)
      K3LA_CASE(8) K3LA_CASE(9) K3LA_CASE(10) K3LA_CASE(11) K3LA_CASE(12) K3LA_CASE(13) K3LA_CASE(14) K3LA_CASE(15)
      K3LA_CASE(16) K3LA_CASE(17) K3LA_CASE(18) K3LA_CASE(19) K3LA_CASE(20) K3LA_CASE(21) K3LA_CASE(22) K3LA_CASE(23)
      default: break;
    }
#undef K3LA_CASE
  k3la_done:
    if (!arrived) {
      K3LA_MARK(2)
      K3TR(kb, 3)
      k3la_bar_arrive(1, ATHREADS);
    }
    K3LA_MARK(3)
    K3TR(kb, 4)
    k3la_bar_sync(2, ATHREADS);                                         // Y(kb): diag(kb+1) done, all updates done
    K3LA_MARK(4)
    K3TR(kb, 5)
  }
  k3la_bar_arrive(4, ATHREADS);                                       // F
#ifdef K3_TIMING
  if (tid == 0 && bs == 20000) printf("K3LA update warp 0: panel rows %lld wait Z %lld 4a %lld 4b %lld wait Y %lld\n", tk[0], tk[1], tk[2], tk[3], tk[4]);
#endif
}

template <int NT, int DT>
int launch_k3la_d(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                  double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  size_t smem = k3la_smem_bytes<NT>(d);
  static PerDeviceOnce once;
#ifdef K3_TRACE
  // occupancy experiment: GPRTO_K3_SMEM_PAD=<bytes> pads the dynamic shared memory (120000 -> one CTA per SM)
  size_t pad = getenv("GPRTO_K3_SMEM_PAD") ? size_t(atol(getenv("GPRTO_K3_SMEM_PAD"))) : 0;
  if (pad > smem) {
    cudaFuncSetAttribute(k3_nlml_la_kernel<NT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pad));
    smem = pad;
  }
#endif
  if (once.first()) {
    if (k3la_smem_bytes<NT>(K3R_MAXD) > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k3_nlml_la_kernel<NT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           int(k3la_smem_bytes<NT>(K3R_MAXD)));
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    cudaFuncSetAttribute(k3_nlml_la_kernel<NT, DT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  k3_nlml_la_kernel<NT, DT><<<(unsigned)(B * S), 32 * K3LA<NT>::W, smem, s>>>(S, n, d, Xn, Yn, theta, jitter, nll, status, gidx);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

template <int NT>
int launch_k3la(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  switch (d) {
    case 2: return launch_k3la_d<NT, 2>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 3: return launch_k3la_d<NT, 3>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return launch_k3la_d<NT, 0>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
}

#ifdef GPRTO_K3_TU
// Look-ahead kernel where it measured faster than k3_nlml_tile_kernel (B200, d = 3, evals/s look-ahead vs plain):
// n = 48: 4.9e7 vs 4.6e7, n = 64: 3.4e7 vs 3.1e7, n = 80: 1.9e7 vs 2.1e7 (so not used), n = 96: 1.4e7 vs 1.1e7,
// n = 112: 1.1e7 vs 7.4e6, n = 128: 8.4e6 vs 5.9e6; returns GPRTO_EUNSUPPORTED otherwise (the caller falls back to
// launch_nlml_reg).
int launch_nlml_la(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                   double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx) {
  if (n > K3R_MAXN || d > K3R_MAXD || n <= 32) return GPRTO_EUNSUPPORTED;
  const int nt = (n + 15) / 16 * 2;
  switch (nt) {
    case 6: return launch_k3la<6>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 8: return launch_k3la<8>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 12: return launch_k3la<12>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 14: return launch_k3la<14>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 16: return launch_k3la<16>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return GPRTO_EUNSUPPORTED;
}
#endif  // GPRTO_K3_TU

}  // namespace gprto
