// Translation unit holding the K4 DMMA kernel instances for NB in {10, 12} (see gprto_launch.h).
#include "gprto_launch.h"
#include "k4_posterior.cuh"

namespace gprto {
namespace {
template <int NB, int D>
int launch(const K4Params& p, cudaStream_t s) {
  constexpr size_t smem = k4_smem_bytes<NB, D>();
  static PerDeviceOnce once;        // attributes are per function AND per device
  if (once.first()) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k4_dmma_kernel<NB, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    cudaFuncSetAttribute(k4_dmma_kernel<NB, D>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  dim3 grid((unsigned)p.B, p.nchunks);      // GP on x (2^31-1 limit), candidate chunk on y
  k4_dmma_kernel<NB, D><<<grid, K4_THREADS, smem, s>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}
template <int NB>
int by_d(const K4Params& p, cudaStream_t s) {
  switch (p.d) {
    case 1: return launch<NB, 1>(p, s);
    case 2: return launch<NB, 2>(p, s);
    case 3: return launch<NB, 3>(p, s);
    case 4: return launch<NB, 4>(p, s);
    case 5: return launch<NB, 5>(p, s);
    case 6: return launch<NB, 6>(p, s);
    case 7: return launch<NB, 7>(p, s);
    case 8: return launch<NB, 8>(p, s);
  }
  return GPRTO_EUNSUPPORTED;
}
}  // namespace

int gprto_k4_launch_c(int nb, const K4Params& p, cudaStream_t s) {
  switch (nb) {
    case 9: return by_d<10>(p, s);
    case 10: return by_d<10>(p, s);
    case 11: return by_d<12>(p, s);
    case 12: return by_d<12>(p, s);
  }
  return GPRTO_EUNSUPPORTED;
}
}  // namespace gprto
