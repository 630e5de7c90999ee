// K3, "lc": the lb kernel (k3_nlml_lb.cuh) with the whole pivot chain inside ONE warp and no per-slot code outside the
// trailing loop.  Replaces GP_model.negative_loglikelihood (reference: sub_uts/utilities_2.py:92-114).
//
// In lb the next diagonal tile travels panel warp -> (barrier) -> owner warp: solve tile (kb+1, kb), update tile (kb+1, kb+1),
// spill -> (barrier) -> panel warp.  The timeline (profiles/r02_k3_timeline.md) shows 550-600 cycles for that hand-over at every
// step - the owner runs a different copy of the per-slot code each time, i.e. cold instruction-cache lines - plus two
// barrier wake-ups, against 1,300-1,500 cycles for the elimination of the diagonal tile itself.
//   * The panel warp keeps Linv(kb) in registers, turns it into B fragments with four shuffles, solves tile (kb+1, kb) itself
//     (two DMMA; it also writes the solved tile into the panel for everybody else), applies it to tile (kb+1, kb+1) (two DMMA)
//     - the result is already in the register layout of the elimination - and goes straight on with column kb+1.  What it
//     needs from the update warps was finished during the PREVIOUS step: tile (kb+1, kb) updated through panel kb-1 (in the
//     panel buffer) and tile (kb+1, kb+1) updated through panel kb-1, which its owner spills to a side buffer ("diag2" tile of
//     step kb-1) and never touches again.
//   * Update warps: the panel solve L(I, kb) = A(I, kb) Linv^T, I >= kb+2, is a short loop over tile rows (no tile registers
//     involved, the same code every step); after Z ONE unrolled loop updates every live tile of the warp - the tiles of
//     column kb+1 and the diag2 tile are simply the last slots of that prefix and are spilled by a predicated store.
//   * K evaluation: tile column 0 and tile (1,1) first, so that the panel warp factorises columns 0 and 1 while the rest of K is
//     still being evaluated; interior off-diagonal tiles skip the padding / diagonal selects.
//   barriers (ids alternate with the parity of the step, because the panel warp may arrive for the next phase before the
//   last update warp has left the previous one):
//     X'(kb)  update warps arrive at the end of step kb; the panel warp waits for X'(kb-1) before block kb (-> Linv(kb+1))
//     Z(kb)   panel kb solved in place (update warps wait; the panel warp arrives once it has stored tile (kb+1, kb))
//     Y(kb)   the panel warp arrives when Linv(kb+1) is in Li[(kb+1)&1]; the update warps wait for it at the top of step kb+1
#pragma once
#include <mutex>

#include "k3_nlml_lb.cuh"

// Table-based exp in the evaluation of K: n = 128 1.265e7 evals/s against 1.246e7 with the polynomial-only exp_det, n = 64 level
// (B200, d = 3); accuracy in the same class (0.98 ulp against < 1 ulp).  -DK3LC_TABLE_EXP=0 restores the bits of exp_det.
#ifndef K3LC_TABLE_EXP
#define K3LC_TABLE_EXP 1
#endif

namespace gprto {

// Dealing of the tiles to the 7 update warps.  Shipped: one by one in reversed column-major order (columns from the last to
// the first, rows from the bottom up, the diagonal tile last within its column), so that the tiles alive at step kb
// (J >= kb+2) are a prefix of every warp's slots and the tiles of column kb+1 follow.  Measured alternative (BY_ROWS): whole
// tile ROWS per warp, longest row first to the least loaded warp - consecutive slots then share the A fragment L(I, kb),
// which is reloaded only when the row changes (the trailing loop is bound by shared-memory bandwidth,
// profiles/r02_k3_timeline.md par. 4).  It cut the operand loads by ~45 % but costs balance (21 / 19 tiles per warp instead
// of 20 / 19 at n = 128, and late rows stay long while early ones die): n = 128 1.19e7 evals/s against 1.24e7, n = 96
// 2.15e7 against 2.12e7 - not shipped.
template <int NT>
struct K3Deal {
  static constexpr int UW = 7;
  static constexpr bool BY_ROWS = false;                // see above
  int owner[NT];
  int count[UW];
  int rows[UW];
  int slots_all, slots;
  constexpr K3Deal() : owner{}, count{}, rows{}, slots_all(0), slots(0) {
    for (int I = NT - 1; I >= 0; --I) {
      int u = 0;
      for (int v = 1; v < UW; ++v)
        if (count[v] < count[u]) u = v;
      owner[I] = u;
      count[u] += I + 1;
      rows[u] += 1;
    }
    for (int u = 0; u < UW; ++u) {
      if (count[u] > slots_all) slots_all = count[u];
      if (count[u] - rows[u] > slots) slots = count[u] - rows[u];
    }
    if (!BY_ROWS) {                                     // tiles dealt one by one in reversed column-major order
      slots_all = (NT * (NT + 1) / 2 + UW - 1) / UW;
      slots = (NT * (NT + 1) / 2 - NT + UW - 1) / UW;
    }
  }
};

template <int NT>
struct K3LC {
  static constexpr K3Deal<NT> deal{};
  static constexpr int UW = 7;
  static constexpr int SLOTS_ALL = deal.slots_all;        // most tiles any warp owns
  static constexpr int SLOTS = deal.slots;                // most tiles outside tile column 0 (those never enter registers)
  static constexpr int TABN = (SLOTS_ALL + 2) * UW;       // tile table, padded (the prefetch runs two slots ahead)
  static constexpr int STABN = NT * UW;
  // RS = number of slots per warp kept in registers (the last tile columns: they live longest and are updated most often);
  // the remaining slots - early columns, few updates each - live in shared memory: n = 128 fits three CTAs per SM.
  static constexpr int RS = NT >= 12 ? 7 : (NT == 10 ? 4 : SLOTS);
};

constexpr uint32_t K3LC_VALID = 1u << 15, K3LC_ARRIVE = 1u << 14, K3LC_OFFMASK = 0x3fff;

// Host: tile table [TABN] (byte offset 768 I | flags, 768 J << 16; entry s * UW + u = slot s of warp u) followed by the
// step table [NT][UW] (nact | nspill << 5 | diag2 << 10).
template <int NT>
void k3lc_build_tables(uint32_t* tab) {
  constexpr K3Deal<NT> deal{};
  constexpr int UW = 7, TABN = K3LC<NT>::TABN;
  for (int e = 0; e < TABN + NT * UW; ++e) tab[e] = 0;
  int cntJ2[UW] = {};
  int lists[UW][K3LC<NT>::SLOTS_ALL + 1][2];
  int len[UW] = {};
  int k = 0;                                            // position in the global order: columns from the last to the first,
  for (int J = NT - 1; J >= 0; --J)                     // rows from the bottom up, the diagonal tile last within its column
    for (int I = NT - 1; I >= J; --I, ++k) {
      const int u = K3Deal<NT>::BY_ROWS ? deal.owner[I] : k % UW;
      lists[u][len[u]][0] = I; lists[u][len[u]][1] = J; ++len[u];
    }
  for (int u = 0; u < UW; ++u) {
    for (int s = 0; s < len[u]; ++s) {
      tab[s * UW + u] = uint32_t(64 * K3R_LDP * lists[u][s][0]) | K3LC_VALID | (uint32_t(64 * K3R_LDP * lists[u][s][1]) << 16);
      if (lists[u][s][1] >= 2) ++cntJ2[u];
    }
    tab[cntJ2[u] * UW + u] |= K3LC_ARRIVE;              // first slot (from the top) with J <= 1: X'(-1) after it
  }
  for (int kb = 0; kb < NT; ++kb)
    for (int u = 0; u < UW; ++u) {
      int nall = 0, n4b = 0;
      for (int s = 0; s < len[u]; ++s) {
        if (lists[u][s][1] >= kb + 1) ++nall;
        if (lists[u][s][1] >= kb + 2) ++n4b;
      }
      bool diag_owner = false, d2 = false;
      for (int s = 0; s < len[u]; ++s) {
        if (lists[u][s][0] == kb + 1 && lists[u][s][1] == kb + 1) diag_owner = true;
        if (lists[u][s][0] == kb + 2 && lists[u][s][1] == kb + 2) d2 = true;
      }
      tab[TABN + kb * UW + u] = uint32_t(nall - (diag_owner ? 1 : 0)) | (uint32_t(n4b - (d2 ? 1 : 0)) << 5) | (d2 ? 1u << 10 : 0u);
    }
}

template <int NT, int RS>
constexpr size_t k3lc_smem_bytes(int d) {
  return sizeof(double) * (2 * size_t(8 * NT) * K3R_LDP + 4 * 8 * K3R_LDP + size_t(8 * NT) * d + 2 * 8 * NT + 16 + 32 + 2 +
                           size_t(K3LC<NT>::SLOTS - RS) * 7 * 64) +
         4 * K3LC<NT>::TABN + 4 * NT * 7 + 16;
}

// exp for the kernel matrix (K3LC_TABLE_EXP): a = (32 n + j) ln2/32 + r, |r| <= ln2/64, exp(a) = 2^n T[j] e^r
// with T[j] = 2^(j/32) correctly rounded (32 doubles in shared memory) and e^r - 1 = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120 +
// r^4/720) (truncation 3e-18), result T + T (e^r - 1) in one fma: 0.98 ulp maximum error against mpmath over [-700, 10]
// (exp_det: < 1 ulp), 11 FP64 operations instead of 15.  The argument is 2 log sf - r^2/2 <= 2 log sf < 709 inside the
// reference's hyper-parameter box, so only the underflow side is clamped.
__device__ const double k3lc_exp2_tab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0,
    0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0,
    0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0, 0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,
    0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

__device__ __forceinline__ double k3lc_exp_tab(double a, const double* tab) {
  const double LOG2E32 = 0x1.71547652b82fep+5, LN2_32_HI = 0x1.62e42fc000000p-6, LN2_32_LO = 0x1.7d1cf79abc9e4p-33;
  const double MAGIC = 6755399441055744.0;
  const double tt = fma(a, LOG2E32, MAGIC);
  const int m = __double2loint(tt);
  const double mf = tt - MAGIC;
  double r = fma(mf, -LN2_32_HI, a);
  r = fma(mf, -LN2_32_LO, r);
  const double T = tab[m & 31];
  double sp = 1.0 / 720.0;
  sp = fma(sp, r, 1.0 / 120.0);
  sp = fma(sp, r, 1.0 / 24.0);
  sp = fma(sp, r, 1.0 / 6.0);
  sp = fma(sp, r, 0.5);
  const double q = fma(r * r, sp, r);
  const double res = fma(T, q, T);
  const double sc = __hiloint2double(__double2hiint(res) + ((m >> 5) << 20), __double2loint(res));
  return a < -700.0 ? 0.0 : sc;
}

// exp_det without the overflow clamp (same bits as exp_det for a < 709)
__device__ __forceinline__ double k3lc_exp(double a) {
  const double LOG2E = 1.4426950408889634074, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  const double MAGIC = 6755399441055744.0;
  const double tt = fma(a, LOG2E, MAGIC);
  const int nn = __double2loint(tt);
  const double nf = tt - MAGIC;
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
  const double res = __hiloint2double(__double2hiint(p) + (int)((unsigned)nn << 20), __double2loint(p));
  return a < -700.0 ? 0.0 : res;
}

// k3_diag_tile_inv on a tile that is already in registers (x0, x1 = A[r][2q], A[r][2q+1]); returns Linv in the same layout
// (li0, li1) and writes it to `out` ([8][ldp]) for the update warps.  Two more instruction-count cuts (the warp is issue-
// and latency-bound, 1,300 - 2,800 cycles per tile inside the kernel against ~800 alone):
//   * nothing reads the columns of A at or left of the pivot column again (only L^-1 leaves this routine), so the update
//     is written back unconditionally; the pivot d'_c is latched by lane c when it is broadcast;
//   * the power-of-two normalisation is applied on even columns only.  A normalised step bounds every entry by 4x the
//     largest entry before it, an unnormalised one squares the bound: from |K| <= 2.2e4 (the reference's sf2 <= e^10) the
//     eight steps stay below 2e92, the scale S below 2e85 (binary64 overflows at 1.8e308).
__device__ __forceinline__ void k3_diag_tile_inv_regs(double x0, double x1, double* out, int ldp, int lane, int col0, double& li0,
                                                      double& li1, double& lp_out, int& info_out) {
  const int r = lane >> 2, q = lane & 3;
  double w0 = (r == 2 * q) ? 1.0 : 0.0, w1 = (r == 2 * q + 1) ? 1.0 : 0.0;
  double mydp = 1.0;
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
    if (lane == c) mydp = dpiv;
    double n0 = fma(-arc, ak0, dpiv * x0), n1 = fma(-arc, ak1, dpiv * x1);
    double m0 = fma(-arc, wc0, dpiv * w0), m1 = fma(-arc, wc1, dpiv * w1);
    if ((c & 1) == 0) {
      const double f = __hiloint2double(0x7fe00000 - (__double2hiint(dpiv) & 0x7ff00000), 0);      // 2^-e(dpiv)
      n0 *= f; n1 *= f; m0 *= f; m1 *= f;
    }
    x0 = n0;
    x1 = n1;
    w0 = (r > c) ? m0 : w0;
    w1 = (r > c) ? m1 : w1;
  }
  // lane c < 8 fetches S_c = W[c][c] from lane 4c + (c >> 1), component c & 1
  const int src = (4 * lane + (lane >> 1)) & 31;
  const double dw0 = __shfl_sync(0xffffffffu, w0, src), dw1 = __shfl_sync(0xffffffffu, w1, src);
  const double myS = (lane & 1) ? dw1 : dw0;
  const unsigned bad = __ballot_sync(0xffffffffu, lane < 8 && !(mydp > 0.0));
  const double tmine = lane < 8 ? rsqrt_fast(mydp * myS) : 1.0;
  const double tr = __shfl_sync(0xffffffffu, tmine, r);
  li0 = w0 * tr;
  li1 = w1 * tr;
  *reinterpret_cast<double2*>(&out[r * ldp + 2 * q]) = make_double2(li0, li1);
  lp_out = lane < 8 ? mydp * tmine : 1.0;
  info_out = bad ? col0 + __ffs(bad) : 0;
}

template <int NT, int DT, int RS>
__global__ void __launch_bounds__(K3LB<NT>::THREADS, ((NT <= 8 || RS < K3LC<NT>::SLOTS) ? 3 : 2))
k3_nlml_lc_kernel(int S, int n, int d_rt, const double* __restrict__ Xn, const double* __restrict__ Yn,
                  const double* __restrict__ theta_all, double jitter, double* __restrict__ nll,
                  int32_t* __restrict__ status, const int32_t* __restrict__ gidx, const uint32_t* __restrict__ tables) {
  using C = K3LC<NT>;
  constexpr int UW = C::UW, SLOTS = C::SLOTS, SLOTS_ALL = C::SLOTS_ALL, NP = 8 * NT, LDP = K3R_LDP;
  constexpr int THREADS = K3LB<NT>::THREADS, TABN = C::TABN;
  constexpr int UTHREADS = 32 * UW, ATHREADS = 32 * (UW + 1), PANEL_W = K3LB<NT>::W - 1;
  const int d = DT > 0 ? DT : d_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* P = reinterpret_cast<double*>(smem_raw);   // [2][NP][LDP]
  double* Li = P + 2 * NP * LDP;                     // [2][8][LDP]: Linv(kb) in slot kb & 1
  double* Dq = Li + 2 * 8 * LDP;                     // [2][8][LDP]: diagonal tile (j, j) updated through panel j-2, slot j & 1
  double* Csm = Dq + 2 * 8 * LDP;                    // [SLOTS - RS][UW][32][2]: tiles of the slots that are not kept in registers
  double* Xs = Csm + (SLOTS - RS) * UW * 64;         // [NP][d]
  double* ys = Xs + NP * d;                          // [NP]
  double* zall = ys + NP;                            // [NP]: z = L^-1 y
  double* nhw = zall + NP;                           // [16]
  double* quad_s = nhw + 16;                         // [2]: |z|^2
  double* etab = quad_s + 2;                         // [32]: 2^(j/32) (K3LC_TABLE_EXP)
  uint32_t* tOff = reinterpret_cast<uint32_t*>(etab + 32);   // [TABN]: slot s of warp u at s * UW + u: byte offset 768 I | flags, 768 J << 16
  uint32_t* stab = tOff + TABN;                      // [NT][UW]: per (step, warp) constants
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
#if K3LC_TABLE_EXP
  if (tid < 32) etab[tid] = k3lc_exp2_tab[tid];
#endif
  for (int e = tid; e < TABN + NT * UW; e += THREADS) tOff[e] = tables[e];      // tile table + step table (contiguous)
  __syncthreads();

  const int fsrc = 4 * g + (t >> 1);                                    // accumulator -> fragment shuffle source
  const bool fodd = t & 1;
  const int lbase = g * LDP + t;                                        // lane offset of a fragment element inside a panel tile

  if (w == PANEL_W) {
    // ================================ panel warp ================================
    double pm = 1.0, lp, li0, li1;
    int pe = 0, info = 0;
    K3TR(31, 0)
    k3la_bar_sync(5, ATHREADS);                                         // X'(-1): column 0 in panel buffer 0, tile (1,1) in Dq[1]
    K3TR(31, 1)
    {
      const double2 v = *reinterpret_cast<const double2*>(&P[g * LDP + 2 * t]);
      k3_diag_tile_inv_regs(v.x, v.y, Li, LDP, lane, 0, li0, li1, lp, info);
    }
    if (lane == 0) s_inf[0] = info;
    K3TR(31, 2)
    k3lb_bar_arrive(6, ATHREADS);                                       // Y(-1)
    k3_diag_tile_logdet(lp, pm, pe);
#pragma unroll 1
    for (int kb = 0; kb + 1 < NT; ++kb) {
      if (info != 0) break;
      K3TR(kb, 0)
      if (kb >= 1) k3la_bar_sync((kb - 1) & 1 ? 5 : 1, ATHREADS);       // X'(kb-1): tiles (kb+1, kb) and (kb+1, kb+1) through panel kb-1
      K3TR(kb, 1)
      // in-warp hand-over: L(kb+1, kb) = A(kb+1, kb) Linv(kb)^T, A(kb+1, kb+1) -= L L^T
      double* pk = P + (kb & 1) * NP * LDP + 8 * (kb + 1) * LDP;
      const double2 cv = *reinterpret_cast<const double2*>(&Dq[((kb + 1) & 1) * 8 * LDP + g * LDP + 2 * t]);
      double bl0, bl1, xj0, xj1, f0, f1;
      k3lb_acc_to_frag(li0, li1, fsrc, fodd, bl0, bl1);
      k3lb_solve_tile(pk + lbase, bl0, bl1, xj0, xj1);
      *reinterpret_cast<double2*>(&pk[g * LDP + 2 * t]) = make_double2(xj0, xj1);      // L(kb+1, kb) for the update warps
      k3lb_bar_arrive(kb & 1 ? 7 : 3, ATHREADS);                        // Z(kb)
      k3lb_acc_to_frag(xj0, xj1, fsrc, fodd, f0, f1);
      double c0 = cv.x, c1 = cv.y, u0 = 0.0, u1 = 0.0;
      dmma884(c0, c1, f0, -f0);
      dmma884(u0, u1, f1, -f1);
      c0 += u0; c1 += u1;
      K3TR(kb, 6)
      k3_diag_tile_inv_regs(c0, c1, Li + ((kb + 1) & 1) * 8 * LDP, LDP, lane, 8 * (kb + 1), li0, li1, lp, info);
      if (lane == 0) s_inf[(kb + 1) & 1] = info;
      K3TR(kb, 2)
      k3lb_bar_arrive(kb & 1 ? 6 : 2, ATHREADS);                        // Y(kb)
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
  K3TR(31, 0)
  double C0[RS > 0 ? RS : 1], C1[RS > 0 ? RS : 1];
  double* csm = Csm + uw * 64 + 2 * lane;            // this lane's pair in the warp's first shared-memory slot (slot stride UW * 64)
  {
    double hw[DT > 0 ? DT : 1];
    if (DT > 0) {
#pragma unroll
      for (int dd = 0; dd < (DT > 0 ? DT : 1); ++dd) hw[dd] = nhw[dd];
    }
    // slots from high to low: tile column 0 and tile (1,1) come first and are handed to the panel warp at once
#pragma unroll
    for (int s = SLOTS_ALL - 1; s >= 0; --s) {
      const uint32_t ij = tOff[s * UW + uw];
      const uint32_t oI = ij & K3LC_OFFMASK, oJ = ij >> 16;
      const int I8 = int((oI * 683u) >> 16), J8 = int((oJ * 683u) >> 16);      // 8 I, 8 J  (768 I * 683 >> 16)
      {                                                                 // (a padding entry evaluates tile (0,0) again; nothing of it is stored)
        const int i = I8 + g, j0 = J8 + 2 * t;
        const double* xi = Xs + i * d;
        const double* xj = Xs + j0 * d;
        double a0 = c0, a1 = c0;
        if (DT > 0) {
#pragma unroll
          for (int dd = 0; dd < (DT > 0 ? DT : 1); ++dd) {
            const double x = xi[dd], h = hw[dd];
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
#if K3LC_TABLE_EXP
        double k0 = k3lc_exp_tab(a0, etab), k1 = k3lc_exp_tab(a1, etab);
#else
        double k0 = k3lc_exp(a0), k1 = k3lc_exp(a1);
#endif
        if (I8 == J8 || I8 + 8 > n) {                                   // (warp-uniform) diagonal tile or padded rows / columns
          k0 += (i == j0 ? diag_add : 0.0);
          k1 += (i == j0 + 1 ? diag_add : 0.0);
          const bool in0 = i < n && j0 < n, in1 = i < n && j0 + 1 < n;
          k0 = in0 ? k0 : (i == j0 ? 1.0 : 0.0);
          k1 = in1 ? k1 : (i == j0 + 1 ? 1.0 : 0.0);
        }
        if (s < RS) { C0[s] = k0; C1[s] = k1; }
        else if (s < SLOTS) *reinterpret_cast<double2*>(csm + (s - RS) * UW * 64) = make_double2(k0, k1);
        if ((K3Deal<NT>::BY_ROWS || s >= SLOTS - 1) && oJ == 0 && (ij & K3LC_VALID))      // tile column 0 goes straight to panel buffer 0
          *reinterpret_cast<double2*>(&P[(I8 + g) * LDP + 2 * t]) = make_double2(k0, k1);
        if ((K3Deal<NT>::BY_ROWS || s == (NT * (NT + 1) / 2 - 1 - NT) / UW) && oI == 64 * LDP && oJ == 64 * LDP)      // tile (1,1): no update before the panel warp takes it
          *reinterpret_cast<double2*>(&Dq[8 * LDP + g * LDP + 2 * t]) = make_double2(k0, k1);
      }
      if (ij & K3LC_ARRIVE) {
        K3TR(31, 3)
        k3lb_bar_arrive(5, ATHREADS);                                   // X'(-1): this warp's tiles of columns 0 and 1 are out
      }
    }
  }
  K3TR(31, 1)

  const uint32_t tb32 = (uint32_t)__cvta_generic_to_shared(tOff + uw);
  bool failed = false;
#pragma unroll 1
  for (int kb = 0; kb < NT; ++kb) {
    k3la_bar_sync((kb - 1) & 1 ? 6 : 2, ATHREADS);                      // Y(kb-1): Linv(kb) ready; all updates of step kb-1 done
    if (s_inf[kb & 1] != 0) { failed = true; break; }
    const uint32_t st = stab[kb * UW + uw];
    double* buf = P + (kb & 1) * NP * LDP;
    double* nbuf = P + ((kb + 1) & 1) * NP * LDP;
    double* zk = zall + 8 * kb;
    K3TR(kb, 0)
    const double* fb = buf + lbase;                                     // fragment base of this lane in the current panel
    const double* lib = Li + (kb & 1) * 8 * LDP + lbase;
    const double bl0 = lib[0], bl1 = lib[4];                            // B fragment of the panel solve: B[k][n] = Linv[n][k]
    if (uw == UW - 1) {
      // right-hand side: z_kb = Linv y'_kb; lane (g,t) adds Linv[g][t] y[t] + Linv[g][4+t] y[4+t], the quad holds z[g]
      double p = fma(bl0, ys[8 * kb + t], bl1 * ys[8 * kb + 4 + t]);
      p += __shfl_xor_sync(0xffffffffu, p, 1);
      p += __shfl_xor_sync(0xffffffffu, p, 2);
      if (t == 0) zk[g] = p;
    }
    if (kb + 1 == NT) break;

    // (3) panel solve L(I, kb) = A(I, kb) Linv^T in place, tile rows kb+2+uw, +7, ..  (row kb+1 is the panel warp's)
    K3TR(kb, 1)
#pragma unroll 1
    for (int I = kb + 2 + uw; I < NT; I += UW) {
      double x0, x1;
      k3lb_solve_tile(fb + 8 * I * LDP, bl0, bl1, x0, x1);
      *reinterpret_cast<double2*>(&buf[(8 * I + g) * LDP + 2 * t]) = make_double2(x0, x1);
    }
    K3TR(kb, 2)
    k3la_bar_sync(kb & 1 ? 7 : 3, ATHREADS);                            // Z(kb): panel kb solved, z_kb written
    K3TR(kb, 3)

    // y'_r -= L(r, kb) z_kb for the rows below block kb, one row per thread
    {
      const int nrows = NP - 8 * (kb + 1);
      if (utid < nrows) {
        const double* rp = buf + (8 * (kb + 1) + utid) * LDP;
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

    // (4) update of every live tile of this warp (a prefix of its slots), fragments of the next tile in flight; the slots
    // from nspill on are the diag2 tile (side buffer, for the panel warp) and the tiles of column kb+1 (next panel buffer)
    {
      const int nact = st & 31, nspill = (st >> 5) & 31;
      const bool d2 = st & (1u << 10);
      const uint32_t fb32 = (uint32_t)__cvta_generic_to_shared(fb);
      const uint32_t sp32 = (uint32_t)__cvta_generic_to_shared(nbuf + g * LDP + 2 * t);      // accumulator position in tile row 0
      const uint32_t dq32 = (uint32_t)__cvta_generic_to_shared(Dq + (kb & 1) * 8 * LDP + g * LDP + 2 * t);
      uint32_t ij = k3lb_lds32(tb32);
      uint32_t offI = ij & K3LC_OFFMASK;
      uint32_t pa = fb32 + offI, pb = fb32 + (ij >> 16);
      double a0 = k3lb_lds64(pa), b0 = k3lb_lds64(pb), a1 = k3lb_lds64_32(pa), b1 = k3lb_lds64_32(pb);
      ij = k3lb_lds32(tb32 + 4 * UW);
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (s >= nact) break;
        const uint32_t offIn = ij & K3LC_OFFMASK;
        pa = fb32 + offIn;
        pb = fb32 + (ij >> 16);
        ij = k3lb_lds32(tb32 + 4 * (s + 2) * UW);
        double na0 = a0, na1 = a1;
        if (!K3Deal<NT>::BY_ROWS || offIn != offI) { na0 = k3lb_lds64(pa); na1 = k3lb_lds64_32(pa); }      // same tile row: same A fragment
        const double nb0 = k3lb_lds64(pb), nb1 = k3lb_lds64_32(pb);
        if (s < RS) {
          dmma884(C0[s < RS ? s : 0], C1[s < RS ? s : 0], a0, -b0);
          dmma884(C0[s < RS ? s : 0], C1[s < RS ? s : 0], a1, -b1);
          if (s >= nspill) {
            const uint32_t dst = (d2 && s == nspill) ? dq32 : sp32 + offI;
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(dst), "d"(C0[s < RS ? s : 0]), "d"(C1[s < RS ? s : 0]) : "memory");
          }
        } else {
          double2 cc = *reinterpret_cast<double2*>(csm + (s - RS) * UW * 64);
          dmma884(cc.x, cc.y, a0, -b0);
          dmma884(cc.x, cc.y, a1, -b1);
          if (s >= nspill) {                                            // column kb+1 / diag2 tile: its last update
            const uint32_t dst = (d2 && s == nspill) ? dq32 : sp32 + offI;
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(dst), "d"(cc.x), "d"(cc.y) : "memory");
          } else {
            *reinterpret_cast<double2*>(csm + (s - RS) * UW * 64) = cc;
          }
        }
        a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
        offI = offIn;
      }
    }
    K3TR(kb, 4)
    if (kb + 2 < NT) k3lb_bar_arrive(kb & 1 ? 5 : 1, ATHREADS);         // X'(kb): what the panel warp needs for block kb+1
  }
  if (!failed) {                                                        // (uniform: a failed factorisation left the loop on every warp)
    k3la_bar_sync(8, UTHREADS);                                         // z complete
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

// Device copy of the tables of one NT (built on first use, one per device; a few hundred bytes, never freed)
template <int NT>
const uint32_t* k3lc_tables_device() {
  static uint32_t* dev_tab[64] = {};
  static std::mutex mtx;                                // handles of several host threads may reach their first launch together
  std::lock_guard<std::mutex> lock(mtx);
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!dev_tab[dev]) {
    constexpr int N = K3LC<NT>::TABN + K3LC<NT>::STABN;
    uint32_t host[N];
    k3lc_build_tables<NT>(host);
    uint32_t* p = nullptr;
    if (cudaMalloc(&p, sizeof(host)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(p, host, sizeof(host), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(p); return nullptr; }
    dev_tab[dev] = p;
  }
  return dev_tab[dev];
}

template <int NT, int DT, int RS = K3LC<NT>::RS>
int launch_k3lc_d(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                  double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  const size_t smem = k3lc_smem_bytes<NT, RS>(d);
  const uint32_t* tables = k3lc_tables_device<NT>();
  if (!tables) return -static_cast<int>(cudaErrorMemoryAllocation);
  static PerDeviceOnce once;
  if (once.first()) {
    if (k3lc_smem_bytes<NT, RS>(K3R_MAXD) > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k3_nlml_lc_kernel<NT, DT, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           int(k3lc_smem_bytes<NT, RS>(K3R_MAXD)));
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    cudaFuncSetAttribute(k3_nlml_lc_kernel<NT, DT, RS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  k3_nlml_lc_kernel<NT, DT, RS><<<(unsigned)(B * S), K3LB<NT>::THREADS, smem, s>>>(S, n, d, Xn, Yn, theta, jitter, nll, status, gidx, tables);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

template <int NT>
int launch_k3lc(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx = nullptr) {
  switch (d) {
    case 2: return launch_k3lc_d<NT, 2>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 3: return launch_k3lc_d<NT, 3>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return launch_k3lc_d<NT, 0>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
}

#ifdef GPRTO_K3_TU
int launch_nlml_lc(int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                   double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx) {
  if (n > K3R_MAXN || d > K3R_MAXD || n <= 32) return GPRTO_EUNSUPPORTED;
  const int nt = (n + 15) / 16 * 2;
  switch (nt) {
    case 6: return launch_k3lc<6>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 8: return launch_k3lc<8>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 10: return launch_k3lc<10>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 12: return launch_k3lc<12>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 14: return launch_k3lc<14>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    case 16: return launch_k3lc<16>(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
  }
  return GPRTO_EUNSUPPORTED;
}
#endif  // GPRTO_K3_TU

}  // namespace gprto
