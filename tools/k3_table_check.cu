// Host-side consistency check of the K3 "lc" tile / step tables for every tile count (no GPU needed): each tile exactly once,
// live tiles a prefix of every warp, spill set of step kb = column kb+1 (rows >= kb+2) + the diag2 tile, one X(-1) arrival
// per warp.  Built and run by tests/test_k3_tables.py.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define GPRTO_K3_TU 1
#include "../gp_rto_ei_b200/csrc/k3_nlml_lc.cuh"
using namespace gprto;
template <int NT> int check() {
  constexpr int UW = 7, TABN = K3LC<NT>::TABN;
  uint32_t tab[TABN + NT * UW];
  k3lc_build_tables<NT>(tab);
  int bad = 0;
  // every tile exactly once
  int seen[NT][NT] = {};
  int cnt[UW] = {};
  for (int e = 0; e < TABN; ++e) if (tab[e] & K3LC_VALID) { int I = (tab[e] & K3LC_OFFMASK) / 768, J = (tab[e] >> 16) / 768; seen[I][J]++; cnt[e % UW]++; if (e / UW >= K3LC<NT>::SLOTS_ALL) bad++; if (J > 0 && e / UW >= K3LC<NT>::SLOTS) bad++; }
  for (int I = 0; I < NT; ++I) for (int J = 0; J <= I; ++J) if (seen[I][J] != 1) bad++;
  int arr = 0; for (int e = 0; e < TABN; ++e) if (tab[e] & K3LC_ARRIVE) arr++;
  if (arr != UW) bad++;
  // per step: live prefix property, spills = column kb+1 (rows >= kb+2) + diag2
  for (int kb = 0; kb + 1 < NT; ++kb) {
    int colseen = 0, d2seen = 0, upd = 0;
    for (int u = 0; u < UW; ++u) {
      uint32_t st = tab[TABN + kb * UW + u]; int nact = st & 31, nsp = (st >> 5) & 31; bool d2 = st & (1u << 10);
      for (int s = 0; s < nact; ++s) {
        uint32_t e = tab[s * UW + u]; if (!(e & K3LC_VALID)) { bad++; continue; }
        int I = (e & K3LC_OFFMASK) / 768, J = (e >> 16) / 768;
        if (J < kb + 1 || (I == kb + 1 && J == kb + 1)) bad++;
        upd++;
        if (s >= nsp) { if (d2 && s == nsp) { if (!(I == kb + 2 && J == kb + 2)) bad++; d2seen++; } else { if (J != kb + 1 || I < kb + 2) bad++; colseen++; } }
        else if (J == kb + 1 || (I == kb + 2 && J == kb + 2)) bad++;
      }
      // slots beyond nact must be dead (J <= kb) or the dead diagonal
      for (int s = nact; s < K3LC<NT>::SLOTS_ALL; ++s) { uint32_t e = tab[s * UW + u]; if (!(e & K3LC_VALID)) continue; int I = (e & K3LC_OFFMASK) / 768, J = (e >> 16) / 768; if (J > kb + 1 || (J == kb + 1 && I != kb + 1)) bad++; }
    }
    if (colseen != NT - kb - 2) bad++;
    if (d2seen != (kb + 2 < NT ? 1 : 0)) bad++;
  }
  printf("NT %2d  SLOTS_ALL %2d SLOTS %2d RS %2d  tiles per warp %d %d %d %d %d %d %d  bad %d\n", NT, K3LC<NT>::SLOTS_ALL, K3LC<NT>::SLOTS, K3LC<NT>::RS, cnt[0], cnt[1], cnt[2], cnt[3], cnt[4], cnt[5], cnt[6], bad);
  return bad;
}
int main() { int b = check<6>() + check<8>() + check<10>() + check<12>() + check<14>() + check<16>(); printf("total bad %d\n", b); return b; }
