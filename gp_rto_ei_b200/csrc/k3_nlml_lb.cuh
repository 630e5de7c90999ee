// K3, look-ahead kernel with a tensor-core panel solve ("lb"): the register-resident tiled Cholesky of k3_nlml_la.cuh rebuilt
// after the round-2 timeline / stall-sampling analysis of that kernel (profiles/r02_k3_timeline.md): its update warps spent
// their time in (a) the one-thread-per-row forward substitution of the panel, (b) integer index arithmetic (div / mod by the
// warp count) and the fetch bubbles behind the taken branches of the per-slot switch, (c) operand loads that were not
// overlapped with the tensor-core instructions.  Replaces GP_model.negative_loglikelihood (reference:
// sub_uts/utilities_2.py:92-114).
//
//   * The panel warp eliminates the 8x8 diagonal tile as before (scaled, division- and sqrt-free) but carries an identity
//     block through the same row operations, so it hands over the INVERSE of the diagonal factor, Linv = L11^-1, written in
//     place of the diagonal tile.  The panel solve L21 = A21 L11^-T is then two DMMA.8x8x4 per 8x8 tile.
//   * Hand-over of the next diagonal tile: its owner solves tile (kb+1, kb), turns the accumulator into fragments with four
//     shuffles, updates the diagonal tile and arrives at a 64-thread named barrier; nobody else is on that path.  The other
//     tiles of column kb+1 are solved / updated by their owners before the Z barrier; z_kb = Linv y'_kb is a fragment
//     product on an otherwise idle warp, and y' -= L21 z_kb is a row-per-thread product after Z (no per-tile special case).
//   * Tiles are dealt to the 7 update warps in REVERSED column-major order: the tiles still alive at step kb are a PREFIX
//     of every warp's slots, so the trailing update is a straight unrolled loop with one not-taken exit test per slot; the
//     fragments of the next tile are loaded while the current one is multiplied.  Everything that depends only on
//     (step, warp) - slot counts, first tile row, roles - comes from a table built once per CTA.
//   * Tile column 0 never enters the register file (it is evaluated straight into the panel buffer).
//   * bar.arrive orders the prior shared-memory writes by itself (PTX ISA, "bar"): no fence in front of it.  The panel warp
//     arrives at Y and never waits there.
#pragma once
#include "gprto_launch.h"
#include "gprto_math.cuh"
#include "k3_nlml_la.cuh"
#include <type_traits>

namespace gprto {

template <int NT>
struct K3LB {
  static constexpr int W = 8, UW = 7;
  static constexpr int TILES = NT * (NT + 1) / 2;
  static constexpr int SLOTS_ALL = (TILES + UW - 1) / UW;                 // slots that hold some tile
  static constexpr int SLOTS = (TILES - NT + UW - 1) / UW;                // slots kept in registers (tile column 0 excluded)
  static constexpr int TABN = (SLOTS_ALL + 2) * UW;                       // tile table, padded (the prefetch runs two slots ahead)
  static constexpr int NP = 8 * NT;
  static constexpr int THREADS = 32 * W;
};

template <int NT>
constexpr size_t k3lb_smem_bytes(int d) {
  return sizeof(double) * (2 * size_t(8 * NT) * K3R_LDP + 8 * K3R_LDP + size_t(8 * NT) * d + 2 * 8 * NT + 16 + 16 + 2) +
         4 * K3LB<NT>::TABN + 2 * NT * K3LB<NT>::UW + 16;
}

__device__ __forceinline__ void k3lb_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// Shared-memory loads as volatile asm: the compiler keeps them where they are written (it sinks ordinary loads down to their
// first use, which undoes a software prefetch placed in front of the tensor-core instructions).
__device__ __forceinline__ double k3lb_lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double k3lb_lds64_32(uint32_t addr) {      // [addr + 32 bytes]: the k = 4..7 half of a fragment
  double v;
  asm volatile("ld.shared.f64 %0, [%1+32];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t k3lb_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Scaled, division- and sqrt-free elimination of the 8x8 diagonal tile (see k3_diag_tile_lean) with the identity carried
// along: after step c the rows r > c of W hold S_r times the rows of the unit-lower elimination matrix (W[r][r] = S_r), and
// L11^-1 = diag(t) W with t_r = rsqrt(d'_r S_r).  Writes L11^-1 (upper triangle exactly zero) over the tile.
// The power-of-two normalisation f = 2^-e(d) is applied AFTER the update, a <- f (d a_jk - a_jc a_kc): the same bits as
// scaling d and a_jc first (a power of two commutes with every rounding), but the integer extraction of the exponent runs
// beside the multiply-add instead of in front of it.  The pivots d'_c and scales S_c are picked up after the loop from the
// frozen diagonal entries of A and W (nothing but the update itself is left inside the loop).
// Returns the lane's L_cc (lanes 0..7, else 1) for k3_diag_tile_logdet, which is called after the tile has been handed over.
__device__ __forceinline__ void k3_diag_tile_inv(double* tile, int ldp, int lane, int col0, double& lp_out, int& info_out) {
  const int r = lane >> 2, q = lane & 3;
  const double2 v = *reinterpret_cast<const double2*>(&tile[r * ldp + 2 * q]);
  double x0 = v.x, x1 = v.y;
  double w0 = (r == 2 * q) ? 1.0 : 0.0, w1 = (r == 2 * q + 1) ? 1.0 : 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int cs = c >> 1;
    const double colsel = (c & 1) ? x1 : x0;
    const double dpiv = __shfl_sync(0xffffffffu, colsel, 4 * c + cs);
    const double arc = __shfl_sync(0xffffffffu, colsel, 4 * r + cs);
    const double ak0 = __shfl_sync(0xffffffffu, colsel, 4 * (2 * q) + cs);
    const double ak1 = __shfl_sync(0xffffffffu, colsel, 4 * (2 * q + 1) + cs);
    const double wc0 = __shfl_sync(0xffffffffu, w0, 4 * c + q);
    const double wc1 = __shfl_sync(0xffffffffu, w1, 4 * c + q);
    const double f = __hiloint2double(0x7fe00000 - (__double2hiint(dpiv) & 0x7ff00000), 0);      // 2^-e(dpiv)
    const double n0 = fma(-arc, ak0, dpiv * x0) * f, n1 = fma(-arc, ak1, dpiv * x1) * f;
    const double m0 = fma(-arc, wc0, dpiv * w0) * f, m1 = fma(-arc, wc1, dpiv * w1) * f;
    x0 = (2 * q > c) ? n0 : x0;
    x1 = (2 * q + 1 > c) ? n1 : x1;
    w0 = (r > c) ? m0 : w0;
    w1 = (r > c) ? m1 : w1;
  }
  // lane c < 8 fetches d'_c = A'[c][c] and S_c = W[c][c] from lane 4c + (c >> 1), component c & 1
  const int src = (4 * lane + (lane >> 1)) & 31;
  const double dx0 = __shfl_sync(0xffffffffu, x0, src), dx1 = __shfl_sync(0xffffffffu, x1, src);
  const double dw0 = __shfl_sync(0xffffffffu, w0, src), dw1 = __shfl_sync(0xffffffffu, w1, src);
  const double mydp = (lane & 1) ? dx1 : dx0, myS = (lane & 1) ? dw1 : dw0;
  const unsigned bad = __ballot_sync(0xffffffffu, lane < 8 && !(mydp > 0.0));      // dpotrf: first non-positive pivot
  const double tmine = lane < 8 ? rsqrt_fast(mydp * myS) : 1.0;
  const double tr = __shfl_sync(0xffffffffu, tmine, r);
  *reinterpret_cast<double2*>(&tile[r * ldp + 2 * q]) = make_double2(w0 * tr, w1 * tr);
  lp_out = lane < 8 ? mydp * tmine : 1.0;      // L_cc = d'_c t_c
  info_out = bad ? col0 + __ffs(bad) : 0;
}

// log det bookkeeping of one diagonal tile (off the hand-over path): the product of the 8 pivots L_cc^2 is folded into a
// running mantissa / exponent pair (one log per matrix at the end).
__device__ __forceinline__ void k3_diag_tile_logdet(double lp, double& pm, int& pe) {
  lp *= __shfl_xor_sync(0xffffffffu, lp, 1);
  lp *= __shfl_xor_sync(0xffffffffu, lp, 2);
  lp *= __shfl_xor_sync(0xffffffffu, lp, 4);
  const double prod = pm * lp * lp;
  const int ex = ((__double2hiint(prod) >> 20) & 0x7ff) - 1023;
  pm = __hiloint2double(__double2hiint(prod) - (ex << 20), __double2loint(prod));
  pe += ex;
}

// accumulator layout (lane (g,t) holds X[g][2t], X[g][2t+1]) -> A/B fragment layout (X[g][t], X[g][4+t])
__device__ __forceinline__ void k3lb_acc_to_frag(double x0, double x1, int src0, bool odd, double& f0, double& f1) {
  const double u0 = __shfl_sync(0xffffffffu, x0, src0), u1 = __shfl_sync(0xffffffffu, x1, src0);
  const double v0 = __shfl_sync(0xffffffffu, x0, src0 + 2), v1 = __shfl_sync(0xffffffffu, x1, src0 + 2);
  f0 = odd ? u1 : u0;
  f1 = odd ? v1 : v0;
}

// X = A Linv^T for one 8x8 tile: two independent DMMA (k = 0..3, 4..7) and one add each
__device__ __forceinline__ void k3lb_solve_tile(const double* rowp, double bl0, double bl1, double& x0, double& x1) {
  double xa0 = 0.0, xa1 = 0.0, xb0 = 0.0, xb1 = 0.0;
  dmma884(xa0, xa1, rowp[0], bl0);
  dmma884(xb0, xb1, rowp[4], bl1);
  x0 = xa0 + xb0;
  x1 = xa1 + xb1;
}

template <int NT, int DT>
__global__ void __launch_bounds__(K3LB<NT>::THREADS, (NT <= 8 ? 3 : 2))
k3_nlml_lb_kernel(int S, int n, int d_rt, const double* __restrict__ Xn, const double* __restrict__ Yn,
                  const double* __restrict__ theta_all, double jitter, double* __restrict__ nll,
                  int32_t* __restrict__ status, const int32_t* __restrict__ gidx) {
  using C = K3LB<NT>;
  constexpr int UW = C::UW, TILES = C::TILES, SLOTS = C::SLOTS, SLOTS_ALL = C::SLOTS_ALL, NP = C::NP, LDP = K3R_LDP;
  constexpr int THREADS = C::THREADS, TABN = C::TABN;
  constexpr int UTHREADS = 32 * UW, ATHREADS = 32 * (UW + 1), PANEL_W = C::W - 1;
  const int d = DT > 0 ? DT : d_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* P = reinterpret_cast<double*>(smem_raw);   // [2][NP][LDP]
  double* Ljs = P + 2 * NP * LDP;                    // [8][LDP]: solved tile (kb+1, kb) (it is not written back into the panel)
  double* Xs = Ljs + 8 * LDP;                        // [NP][d]
  double* ys = Xs + NP * d;                          // [NP]
  double* zall = ys + NP;                            // [NP]: z = L^-1 y
  double* nhw = zall + NP;                           // [16]
  double* quad_s = nhw + 16;                         // [2]: |z|^2
  double* spare = quad_s + 2;                        // [16]
  uint32_t* tOff = reinterpret_cast<uint32_t*>(spare + 16);   // [TABN]: byte offsets (768 I) | (768 J) << 16 of reversed tile tau_r
  unsigned short* stab = reinterpret_cast<unsigned short*>(tOff + TABN);   // [NT][UW]: per (step, warp) constants
  __shared__ int s_inf[2];                           // dpotrf code of the diagonal tile of column kb, slot kb & 1

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t bs = blockIdx.x;
  const int64_t b = gidx ? gidx[bs] : bs / S;
  const double* theta = theta_all + bs * (d + 2);

  for (int e = tid; e < NP * d; e += THREADS) Xs[e] = (e / d) < n ? Xn[b * int64_t(n) * d + e] : 0.0;
  for (int e = tid; e < NP; e += THREADS) ys[e] = e < n ? Yn[b * n + e] : 0.0;
  if (tid < d) nhw[tid] = -0.5 * exp_det(-2.0 * theta[tid]);
  if (tid == 0) { s_inf[0] = 0; s_inf[1] = 0; quad_s[0] = 0.0; }
  for (int e = tid; e < TABN; e += THREADS) {        // reversed column-major order: tau = TILES-1-e = off(J) + (I - J)
    int J = 0, I = 0;
    if (e < TILES) {
      const int tau = TILES - 1 - e;
      while ((J + 1) * NT - (J + 1) * J / 2 <= tau) ++J;
      I = J + tau - (J * NT - J * (J - 1) / 2);
    }
    tOff[e] = uint32_t(64 * LDP * I) | (uint32_t(64 * LDP * J) << 16);      // byte offsets of tile rows 8 I, 8 J in a panel
  }
  for (int e = tid; e < NT * UW; e += THREADS) {
    // step kb, update warp u: tiles right of column kb+1 are tau >= off2 <=> tau_r <= TILES-1-off2: the first n4b slots;
    // the ncol tiles of column kb+1 follow (slots n4b .. n4b+ncol-1); the diagonal tile (kb+1, kb+1) is tau_r = TILES-1-off1.
    const int kb = e / UW, u = e % UW;
    const int off1 = (kb + 1) * NT - (kb + 1) * kb / 2, off2 = (kb + 2) * NT - (kb + 2) * (kb + 1) / 2;
    const int last4b = TILES - 1 - off2, lastcol = TILES - 1 - off1;       // largest tau_r of each group (-1: none)
    const int n4b = last4b >= u ? (last4b - u) / UW + 1 : 0;
    const int nall = lastcol >= u ? (lastcol - u) / UW + 1 : 0;
    const int ncol = nall - n4b;
    const int dw = lastcol >= 0 ? lastcol % UW : 0;                        // owner of the next diagonal tile: hand-over warp
    const int zw = (dw + UW - 1) % UW;                                     // right-hand side warp
    const int shi = nall - 1;                                              // slot of this warp's last tile of column kb+1
    int Ihi = 0;
    if (ncol > 0) { const int tau = TILES - 1 - (shi * UW + u); Ihi = kb + 1 + (tau - off1); }
    stab[e] = (unsigned short)(n4b | (ncol << 5) | (Ihi << 7) | ((u == dw && kb + 1 < NT) ? 1 << 12 : 0) | (u == zw ? 1 << 13 : 0));
  }
  __syncthreads();

  if (w == PANEL_W) {
    // ================================ panel warp ================================
    double pm = 1.0, lp;
    int pe = 0, info = 0;
    K3TR(31, 0)
    k3la_bar_sync(1, 64);                                               // X(-1): tile (0,0) is in panel buffer 0
    K3TR(31, 1)
    k3_diag_tile_inv(P, LDP, lane, 0, lp, info);
    if (lane == 0) s_inf[0] = info;
    K3TR(31, 2)
    k3lb_bar_arrive(2, ATHREADS);                                       // Y(-1): the panel warp never waits at Y
    k3_diag_tile_logdet(lp, pm, pe);
#pragma unroll 1
    for (int kb = 0; kb + 1 < NT; ++kb) {
      if (info != 0) break;
      double* nbuf = P + ((kb + 1) & 1) * NP * LDP;
      K3TR(kb, 0)
      k3la_bar_sync(1, 64);                                             // X(kb): diagonal tile of column kb+1 updated and spilled
      K3TR(kb, 1)
      k3_diag_tile_inv(nbuf + 8 * (kb + 1) * LDP, LDP, lane, 8 * (kb + 1), lp, info);
      if (lane == 0) s_inf[(kb + 1) & 1] = info;
      K3TR(kb, 2)
      k3lb_bar_arrive(2, ATHREADS);                                     // Y(kb)
      k3_diag_tile_logdet(lp, pm, pe);
    }
    k3la_bar_sync(4, ATHREADS);                                         // F: the update warps have finished |z|^2 (quad_s)
    if (lane == 0) {
      status[bs] = info;
      nll[bs] = info == 0 ? quad_s[0] + log_det(pm) + double(pe) * 0.693147180559945309417 : CUDART_NAN;
    }
    return;
  }

  // ================================ update warps ================================
  const int uw = w, utid = tid;                                         // update warps are warps 0 .. UW-1
  const double c0 = 2.0 * theta[d];
  const double diag_add = exp_det(2.0 * theta[d + 1]) + jitter;
  const int lbase = g * LDP + t;                                        // lane offset of a fragment element inside a panel tile
  const int fsrc = 4 * g + (t >> 1);                                    // accumulator -> fragment shuffle source
  const bool fodd = t & 1;
  K3TR(31, 0)
  double C0[SLOTS], C1[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS_ALL; ++s) {
    const int taur = s * UW + uw;
    const uint32_t ij = tOff[taur];
    const int I8 = (ij & 0xffff) / (8 * LDP), J8 = (ij >> 16) / (8 * LDP);        // 8 I, 8 J
    const int i = I8 + g, j0 = J8 + 2 * t;
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
    const double k0 = in0 ? v0 : (i == j0 ? 1.0 : 0.0);
    const double k1 = in1 ? v1 : (i == j0 + 1 ? 1.0 : 0.0);
    if (s < SLOTS) { C0[s] = k0; C1[s] = k1; }
    // tile column 0 (tau_r >= TILES - NT) goes straight to panel buffer 0; only the slots from SLOTS-1 on can hold such tiles
    if (s >= SLOTS - 1 && taur >= TILES - NT && taur < TILES)
      *reinterpret_cast<double2*>(&P[(I8 + g) * LDP + 2 * t]) = make_double2(k0, k1);
  }
  K3TR(31, 1)
  if (uw == (TILES - 1) % UW) k3lb_bar_arrive(1, 64);                   // X(-1): the owner of tile (0,0) (tau_r = TILES-1)
  k3la_bar_sync(2, ATHREADS);                                           // Y(-1): Linv(0) ready, column 0 in the panel
  K3TR(31, 2)

  bool failed = false;
#pragma unroll 1
  for (int kb = 0; kb < NT; ++kb) {
    if (s_inf[kb & 1] != 0) { failed = true; break; }
    const int st = stab[kb * UW + uw];
    double* buf = P + (kb & 1) * NP * LDP;
    double* nbuf = P + ((kb + 1) & 1) * NP * LDP;
    double* zk = zall + 8 * kb;
    K3TR(kb, 0)
    // B fragment of the panel solve: Linv(kb) sits where the diagonal tile was; B[k][n] = Linv[n][k]
    const double* fb = buf + lbase;                                     // fragment base of this lane in the current panel
    const double bl0 = fb[8 * kb * LDP], bl1 = fb[8 * kb * LDP + 4];
    if (st & (1 << 13)) {
      // right-hand side warp: z_kb = Linv y'_kb; lane (g,t) adds Linv[g][t] y[t] + Linv[g][4+t] y[4+t], the quad holds z[g]
      double p = fma(bl0, ys[8 * kb + t], bl1 * ys[8 * kb + 4 + t]);
      p += __shfl_xor_sync(0xffffffffu, p, 1);
      p += __shfl_xor_sync(0xffffffffu, p, 2);
      if (t == 0) zk[g] = p;
    }
    if (kb + 1 == NT) break;

    // (4a) column kb+1: solve the panel tiles this warp needs, update and spill its tiles of the column
    K3TR(kb, 1)
    int ncol = (st >> 5) & 3;
    const int n4b = st & 31;
    if (ncol) {
      const bool desig = st & (1 << 12);
      double xj0, xj1, nb0, nb1;                                        // L(kb+1, kb): accumulator form and (negated) B fragment
      k3lb_solve_tile(fb + 8 * (kb + 1) * LDP, bl0, bl1, xj0, xj1);
      {
        double f0, f1;
        k3lb_acc_to_frag(xj0, xj1, fsrc, fodd, f0, f1);
        nb0 = -f0; nb1 = -f1;
      }
      int I = (st >> 7) & 31;                                           // row of this warp's last-slot tile of the column
      bool first = true;
      auto col_op = [&](auto sc) {
        constexpr int s = decltype(sc)::value;
        double f0, f1;
        if (first && desig) {                                           // the diagonal tile (kb+1, kb+1): operands already there
          f0 = -nb0; f1 = -nb1;
        } else {
          double x0, x1;
          k3lb_solve_tile(fb + 8 * I * LDP, bl0, bl1, x0, x1);
          *reinterpret_cast<double2*>(&buf[(8 * I + g) * LDP + 2 * t]) = make_double2(x0, x1);    // L(I, kb) for the trailing update
          k3lb_acc_to_frag(x0, x1, fsrc, fodd, f0, f1);
        }
        double u0 = 0.0, u1 = 0.0;
        dmma884(C0[s], C1[s], f0, nb0);
        dmma884(u0, u1, f1, nb1);
        C0[s] += u0; C1[s] += u1;
        *reinterpret_cast<double2*>(&nbuf[(8 * I + g) * LDP + 2 * t]) = make_double2(C0[s], C1[s]);
        if (first && desig) {
          k3lb_bar_arrive(1, 64);                                       // X(kb): diagonal tile of column kb+1 handed over
          K3TR(kb, 6)
          *reinterpret_cast<double2*>(&Ljs[g * LDP + 2 * t]) = make_double2(xj0, xj1);            // for y' -= L z below
        }
        first = false;
        I += UW;                                                        // previous slot = tau + UW = next row of the same column
      };
      // this warp's tiles of column kb+1 are slots n4b+ncol-1 (first) down to n4b
#define K3LB_CASE(S) case S: if constexpr (S < SLOTS) { col_op(std::integral_constant<int, S>{}); if (--ncol == 0) goto k3lb_col_done; } [[fallthrough]];
      switch (n4b + ncol - 1) {
        K3LB_CASE(23) K3LB_CASE(22) K3LB_CASE(21) K3LB_CASE(20) K3LB_CASE(19) K3LB_CASE(18) K3LB_CASE(17) K3LB_CASE(16)
        K3LB_CASE(15) K3LB_CASE(14) K3LB_CASE(13) K3LB_CASE(12) K3LB_CASE(11) K3LB_CASE(10) K3LB_CASE(9) K3LB_CASE(8)
        K3LB_CASE(7) K3LB_CASE(6) K3LB_CASE(5) K3LB_CASE(4) K3LB_CASE(3) K3LB_CASE(2) K3LB_CASE(1) K3LB_CASE(0)
        default: break;
      }
#undef K3LB_CASE
    k3lb_col_done:;
    }
    K3TR(kb, 2)
    k3la_bar_sync(3, UTHREADS);                                         // Z(kb): panel kb solved, z_kb written
    K3TR(kb, 3)

    // y'_r -= L(r, kb) z_kb for the rows below block kb, one row per thread (rows of block kb+1 come from Ljs)
    {
      const int nrows = NP - 8 * (kb + 1);
      if (utid < nrows) {
        const double* rp = utid < 8 ? Ljs + utid * LDP : buf + (8 * (kb + 1) + utid) * LDP;
        const double2 l0 = *reinterpret_cast<const double2*>(rp), l1 = *reinterpret_cast<const double2*>(rp + 2);
        const double2 l2 = *reinterpret_cast<const double2*>(rp + 4), l3 = *reinterpret_cast<const double2*>(rp + 6);
        const double2 z0 = *reinterpret_cast<const double2*>(zk), z1 = *reinterpret_cast<const double2*>(zk + 2);
        const double2 z2 = *reinterpret_cast<const double2*>(zk + 4), z3 = *reinterpret_cast<const double2*>(zk + 6);
        double sa = l0.x * z0.x, sb = l0.y * z0.y;
        sa = fma(l1.x, z1.x, sa); sb = fma(l1.y, z1.y, sb);
        sa = fma(l2.x, z2.x, sa); sb = fma(l2.y, z2.y, sb);
        sa = fma(l3.x, z3.x, sa); sb = fma(l3.y, z3.y, sb);
        ys[8 * (kb + 1) + utid] -= sa + sb;
      }
    }

    // (4b) trailing update of the tiles right of column kb+1: the first n4b slots, fragments of the next tile in flight
    {
      const uint32_t fb32 = (uint32_t)__cvta_generic_to_shared(fb);
      const uint32_t tb32 = (uint32_t)__cvta_generic_to_shared(tOff + uw);
      uint32_t ij = k3lb_lds32(tb32);
      uint32_t pa = fb32 + (ij & 0xffff), pb = fb32 + (ij >> 16);
      double a0 = k3lb_lds64(pa), b0 = k3lb_lds64(pb), a1 = k3lb_lds64_32(pa), b1 = k3lb_lds64_32(pb);
      ij = k3lb_lds32(tb32 + 4 * UW);
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (s >= n4b) break;
        pa = fb32 + (ij & 0xffff);
        pb = fb32 + (ij >> 16);
        ij = k3lb_lds32(tb32 + 4 * (s + 2) * UW);
        const double na0 = k3lb_lds64(pa), nb0 = k3lb_lds64(pb), na1 = k3lb_lds64_32(pa), nb1 = k3lb_lds64_32(pb);
        dmma884(C0[s], C1[s], a0, -b0);
        dmma884(C0[s], C1[s], a1, -b1);
        a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
      }
    }
    K3TR(kb, 4)
    k3la_bar_sync(2, ATHREADS);                                         // Y(kb): Linv(kb+1) ready, all updates done
    K3TR(kb, 5)
  }
  if (!failed) {                                                        // (uniform: a failed factorisation left the loop on every warp)
    k3la_bar_sync(3, UTHREADS);                                         // z complete
    if (uw == 0) {
      double qq = 0.0;
#pragma unroll
      for (int e = lane; e < NP; e += 32) qq = fma(zall[e], zall[e], qq);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, off);
      if (lane == 0) quad_s[0] = qq;
    }
  }
  k3lb_bar_arrive(4, ATHREADS);                                         // F
}

template <int NT, int DT>
int launch_k3lb_d(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                  double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  const size_t smem = k3lb_smem_bytes<NT>(d);
  static PerDeviceOnce once;
  if (once.first()) {
    if (k3lb_smem_bytes<NT>(K3R_MAXD) > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k3_nlml_lb_kernel<NT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           int(k3lb_smem_bytes<NT>(K3R_MAXD)));
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    cudaFuncSetAttribute(k3_nlml_lb_kernel<NT, DT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  k3_nlml_lb_kernel<NT, DT><<<(unsigned)(B * S), K3LB<NT>::THREADS, smem, s>>>(S, n, d, Xn, Yn, theta, jitter, nll, status, gidx);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

template <int NT>
int launch_k3lb(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  switch (d) {
    case 2: return launch_k3lb_d<NT, 2>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 3: return launch_k3lb_d<NT, 3>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return launch_k3lb_d<NT, 0>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
}

#ifdef GPRTO_K3_TU
int launch_nlml_lb(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                   double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx) {
  if (n > K3R_MAXN || d > K3R_MAXD || n <= 32) return GPRTO_EUNSUPPORTED;
  const int nt = (n + 15) / 16 * 2;
  switch (nt) {
    case 6: return launch_k3lb<6>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 8: return launch_k3lb<8>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 10: return launch_k3lb<10>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 12: return launch_k3lb<12>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 14: return launch_k3lb<14>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 16: return launch_k3lb<16>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return GPRTO_EUNSUPPORTED;
}
#endif  // GPRTO_K3_TU

}  // namespace gprto
