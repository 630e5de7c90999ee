// K3 register-resident path (placeholder until implemented): reports "unsupported" so the generic path runs.
#pragma once
#include "gprto_math.cuh"
namespace gprto {
constexpr int K3R_MAXN = 128;
constexpr int K3R_MAXD = 8;
inline int launch_nlml_reg(int64_t, int, int, int, const double*, const double*, const double*, double, double*, int32_t*,
                           cudaStream_t) {
  return GPRTO_EUNSUPPORTED;
}
}  // namespace gprto
