// Translation unit holding every K3 register-resident kernel instance (see gprto_launch.h).
#define GPRTO_K3_TU 1
#include "k3_nlml_reg.cuh"
#include "k3_nlml_la.cuh"
#include "k3_nlml_lb.cuh"
#include "k3_nlml_lc.cuh"
#ifdef K3_TRACE
extern "C" int gprto_debug_k3_trace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, gprto::g_k3_trace, sizeof(long long) * size_t(n));
}
#endif
