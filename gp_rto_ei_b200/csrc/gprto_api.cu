// C ABI of the GP + acquisition hot path (see include/gprto.h for the contract and the reference lines
// each entry point replaces).  Host-side dispatch only; the arithmetic is in the .cuh kernels.
#include "../../include/gprto.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>

#define GPRTO_API_TU 1
#include "gprto_launch.h"
#include "k23_chol.cuh"
#include "k4_posterior.cuh"

using namespace gprto;

struct gprto_handle {
  int device = 0;
  int sm_count = 0;
  int k4_path = GPRTO_K4_AUTO;
  int k3_path = 0;
  int k2_path = 0;
  cudaEvent_t ev_order = nullptr;      // gprto_order_after: producer-stream fence for the *_host entry points
  bool order_pending = false;
  int cov_kernel_kind = 0;
  std::atomic<int64_t> launches{0};
  // scratch (grown on demand)
  double* part_score = nullptr;
  int64_t* part_idx = nullptr;
  size_t part_cap = 0;
  unsigned long long* first_dev = nullptr;
  // host-path staging
  cudaStream_t st[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  double* dev_in[2] = {nullptr, nullptr};
  double* dev_out[2] = {nullptr, nullptr};
  size_t stage_in_cap = 0, stage_out_cap = 0;
  // small synchronous host calls (gprto_nlml_indexed_host): one pinned + one device arena
  unsigned char* small_pin = nullptr;
  unsigned char* small_dev = nullptr;
  size_t small_cap = 0;
  cudaStream_t small_stream = nullptr;
};

namespace {

#define GPRTO_CUDA(call)                                   \
  do {                                                     \
    cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) return -static_cast<int>(e_);   \
  } while (0)

constexpr int64_t kMaxGridX = 2147483647LL;      // one CTA per problem: gridDim.x limit

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

template <typename Kern>
int set_smem(Kern k, size_t bytes) {
  if (bytes > 48 * 1024) GPRTO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
  return 0;
}

int ensure_part(gprto_t* h, size_t count) {
  if (count <= h->part_cap) return 0;
  if (h->part_score) { cudaFree(h->part_score); cudaFree(h->part_idx); h->part_score = nullptr; h->part_idx = nullptr; h->part_cap = 0; }
  GPRTO_CUDA(cudaMalloc(&h->part_score, count * sizeof(double)));
  GPRTO_CUDA(cudaMalloc(&h->part_idx, count * sizeof(int64_t)));
  h->part_cap = count;
  return 0;
}

int dispatch_k4(gprto_t* h, const K4Params& p, cudaStream_t s) {
  const int nb = (p.n + 7) / 8;
  int rc;
  if (nb <= 4) rc = gprto_k4_launch_a(nb, p, s);
  else if (nb <= 8) rc = gprto_k4_launch_b(nb, p, s);
  else if (nb <= 12) rc = gprto_k4_launch_c(nb, p, s);
  else rc = gprto_k4_launch_d(nb, p, s);
  if (rc != GPRTO_EUNSUPPORTED) h->launches++;
  return rc;
}

constexpr int K4_DMMA_MAXN = 128, K4_DMMA_MAXD = 8;

int posterior_device(gprto_t* h, int64_t B, int n, int d, int64_t m, const double* Xn, const double* invK,
                     const double* alpha, const double* theta_opt, const double* x_mean, const double* x_std,
                     const double* y_mean, const double* y_std, const double* cand, const double* prior,
                     const double* f_best, int acq, double* mean, double* var, double* score, double* best_score,
                     int64_t* best_idx, cudaStream_t s, const double* centre = nullptr, const double* radius = nullptr,
                     int shape = 0, uint64_t seed = 0) {
  const bool dmma_ok = n <= K4_DMMA_MAXN && d <= K4_DMMA_MAXD;
  bool use_dmma = dmma_ok;
  if (h->k4_path == GPRTO_K4_GENERIC) use_dmma = false;
  if (h->k4_path == GPRTO_K4_DMMA && !dmma_ok) return GPRTO_EUNSUPPORTED;
  if (!use_dmma && (n > K4G_MAXN || d > K4G_MAXD)) return GPRTO_EUNSUPPORTED;
  if (B > kMaxGridX) return GPRTO_EUNSUPPORTED;
  if (!cand && !use_dmma) return GPRTO_EUNSUPPORTED;      // fused sampling lives in the DMMA kernel only
  K4Params p{};
  p.B = B; p.m = m; p.n = n; p.d = d; p.acq = acq;
  p.Xn = Xn; p.invK = invK; p.alpha = alpha; p.theta = theta_opt; p.x_mean = x_mean; p.x_std = x_std;
  p.y_mean = y_mean; p.y_std = y_std; p.cand = cand; p.prior = prior; p.f_best = f_best;
  p.mean = mean; p.var = var; p.score = score;
  p.centre = centre; p.radius = radius; p.shape = shape; p.seed = seed;
  // Chunking: one CTA per GP when the batch alone fills the machine, otherwise split the candidates of each GP
  // so that the grid has >= ~8 CTAs per SM.
  const int64_t gran = use_dmma ? K4_THREADS : K4G_THREADS;
  const int64_t want_ctas = int64_t(h->sm_count) * 8;
  int64_t nchunks = std::max<int64_t>(1, (want_ctas + B - 1) / B);
  int64_t chunk = ((m + nchunks - 1) / nchunks + gran - 1) / gran * gran;
  chunk = std::max<int64_t>(chunk, gran);
  nchunks = (m + chunk - 1) / chunk;
  if (nchunks > 65535) return GPRTO_EUNSUPPORTED;
  p.chunk = chunk;
  p.nchunks = int(nchunks);
  if (best_score) {
    if (nchunks == 1) { p.part_score = best_score; p.part_idx = best_idx; }
    else {
      int rc = ensure_part(h, size_t(B) * nchunks);
      if (rc) return rc;
      p.part_score = h->part_score; p.part_idx = h->part_idx;
    }
  }
  int rc;
  if (use_dmma) rc = dispatch_k4(h, p, s);
  else {
    dim3 grid((unsigned)B, p.nchunks);
    k4_generic_kernel<<<grid, K4G_THREADS, 0, s>>>(p);
    h->launches++;
    rc = launch_status();
  }
  if (rc) return rc;
  if (best_score && nchunks > 1) {
    const int threads = 128;
    const int64_t blocks = (B * 32 + threads - 1) / threads;
    argmin_rows_kernel<<<(unsigned)blocks, threads, 0, s>>>(B, nchunks, h->part_score, h->part_idx, best_score, best_idx);
    h->launches++;
    rc = launch_status();
  }
  return rc;
}

int ensure_stage(gprto_t* h, size_t in_bytes, size_t out_bytes) {
  if (!h->st[0]) {
    for (auto& s : h->st) GPRTO_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      GPRTO_CUDA(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
      GPRTO_CUDA(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
      GPRTO_CUDA(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
    }
  }
  if (in_bytes > h->stage_in_cap) {
    for (int i = 0; i < 2; ++i) {
      if (h->dev_in[i]) cudaFree(h->dev_in[i]);
      GPRTO_CUDA(cudaMalloc(&h->dev_in[i], in_bytes));
    }
    h->stage_in_cap = in_bytes;
  }
  if (out_bytes > h->stage_out_cap) {
    for (int i = 0; i < 2; ++i) {
      if (h->dev_out[i]) cudaFree(h->dev_out[i]);
      GPRTO_CUDA(cudaMalloc(&h->dev_out[i], out_bytes));
    }
    h->stage_out_cap = out_bytes;
  }
  return 0;
}

template <int D>
int launch_sample(gprto_t* h, int64_t B, int64_t m, const double* centre, const double* radius, int shape, uint64_t seed,
                         const int64_t* idx, double* out, cudaStream_t s) {
  const int64_t total = idx ? B : B * m;
  sample_candidates_kernel<D><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(B, m, centre, radius, shape, seed, idx, out);
  h->launches++;
  return launch_status();
}

int dispatch_sample(gprto_t* h, int64_t B, int d, int64_t m, const double* centre, const double* radius, int shape,
                           uint64_t seed, const int64_t* idx, double* out, cudaStream_t s) {
  switch (d) {
    case 1: return launch_sample<1>(h, B, m, centre, radius, shape, seed, idx, out, s);
    case 2: return launch_sample<2>(h, B, m, centre, radius, shape, seed, idx, out, s);
    case 3: return launch_sample<3>(h, B, m, centre, radius, shape, seed, idx, out, s);
    case 4: return launch_sample<4>(h, B, m, centre, radius, shape, seed, idx, out, s);
    case 5: return launch_sample<5>(h, B, m, centre, radius, shape, seed, idx, out, s);
    case 6: return launch_sample<6>(h, B, m, centre, radius, shape, seed, idx, out, s);
    case 7: return launch_sample<7>(h, B, m, centre, radius, shape, seed, idx, out, s);
    case 8: return launch_sample<8>(h, B, m, centre, radius, shape, seed, idx, out, s);
  }
  return GPRTO_EUNSUPPORTED;
}

int nlml_device(gprto_t* h, int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta, double jitter,
                double* nll, int32_t* status, cudaStream_t s, const int32_t* gidx) {
  if (B * int64_t(S) > kMaxGridX) return GPRTO_EUNSUPPORTED;
  if (h->k3_path != 1 && n <= K3R_MAXN && d <= K3R_MAXD) {
    int rc = GPRTO_EUNSUPPORTED;
    // auto: n > 32 -> "lc" (k3_nlml_lc.cuh: pivot chain in one warp, tensor-core panel solve, three CTAs per SM), measured
    // faster than every other path at every size (B200, d = 3, evals/s lc / round-1 kernels: n = 40 5.6e7 / 4.8e7,
    // 64 4.0e7 / 3.3e7, 80 2.9e7 / 2.1e7, 96 2.1e7 / 1.4e7, 112 1.6e7 / 1.1e7, 128 1.23e7 / 0.84e7); n <= 32 -> plain register
    // kernel.  Forced paths (cross-checks in the tests): 2 plain register kernel, 3 "lb", 4 round-1 look-ahead kernel, 5 "lc".
    if (h->k3_path == 5 || h->k3_path == 0) rc = launch_nlml_lc(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    if (h->k3_path == 3) rc = launch_nlml_lb(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    if (h->k3_path == 4) rc = launch_nlml_la(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    if (rc == GPRTO_EUNSUPPORTED) rc = launch_nlml_reg(B, S, n, d, Xn, Yn, theta, jitter, nll, status, s, gidx);
    if (rc != GPRTO_EUNSUPPORTED) { h->launches++; return rc; }
  }
  const size_t smem = ch_smem_bytes(n, d);
  if (smem > 227 * 1024) return GPRTO_EUNSUPPORTED;
  int rc = set_smem(nlml_generic_kernel, smem);
  if (rc) return rc;
  nlml_generic_kernel<<<(unsigned)(B * S), CH_THREADS, smem, s>>>(S, n, d, Xn, Yn, theta, jitter, nll, status, gidx);
  h->launches++;
  return launch_status();
}

}  // namespace

extern "C" {

const char* gprto_version(void) { return "gprto-b200 0.1 (sm_100a)"; }

const char* gprto_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case GPRTO_EINVAL: return "invalid argument";
    case GPRTO_EUNSUPPORTED: return "problem size outside the supported limits";
    case GPRTO_ENODEVICE: return "no sm_100 CUDA device";
    case GPRTO_ENOMEM: return "out of memory";
  }
  if (code < 0 && code > -10000) return cudaGetErrorString(static_cast<cudaError_t>(-code));
  return "unknown error";
}

int gprto_limits(int64_t limits[4]) {
  if (!limits) return GPRTO_EINVAL;
  limits[0] = K4_DMMA_MAXN;
  limits[1] = K3R_MAXN;
  int nmax = 1;
  while (ch_smem_bytes(nmax + 1, CH_MAXD) <= 227 * 1024) ++nmax;      // factor / generic NLML keep the matrix in shared memory
  limits[2] = nmax;
  limits[3] = CH_MAXD;
  return 0;
}

int gprto_create(int device, gprto_t** out) {
  if (!out) return GPRTO_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return GPRTO_ENODEVICE; }
  if (device < 0 || device >= count) return GPRTO_EINVAL;
  cudaDeviceProp prop;
  GPRTO_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return GPRTO_ENODEVICE;     // sm_100a code only; no fallback
  gprto_t* h = new (std::nothrow) gprto_handle();
  if (!h) return GPRTO_ENOMEM;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  DeviceGuard guard(device);
  if (!guard.ok) { delete h; return GPRTO_ENODEVICE; }
  cudaError_t e = cudaMalloc(&h->first_dev, sizeof(unsigned long long));
  if (e != cudaSuccess) { delete h; return -static_cast<int>(e); }
  *out = h;
  return 0;
}

int gprto_destroy(gprto_t* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaFree(h->part_score); cudaFree(h->part_idx); cudaFree(h->first_dev);
  if (h->small_pin) cudaFreeHost(h->small_pin);
  cudaFree(h->small_dev);
  if (h->small_stream) cudaStreamDestroy(h->small_stream);
  if (h->ev_order) cudaEventDestroy(h->ev_order);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->dev_in[i]); cudaFree(h->dev_out[i]);
    if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
    if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
    if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
  }
  for (auto& s : h->st) if (s) cudaStreamDestroy(s);
  delete h;
  return 0;
}

int gprto_set_option(gprto_t* h, const char* name, int64_t value) {
  if (!h || !name) return GPRTO_EINVAL;
  if (!strcmp(name, "k4_path")) { if (value < 0 || value > 2) return GPRTO_EINVAL; h->k4_path = int(value); return 0; }
  if (!strcmp(name, "cov_kernel")) { if (value < 0 || value > 2) return GPRTO_EINVAL; h->cov_kernel_kind = int(value); return 0; }
  if (!strcmp(name, "k3_path")) { if (value < 0 || value > 5) return GPRTO_EINVAL; h->k3_path = int(value); return 0; }
  if (!strcmp(name, "k2_path")) { if (value < 0 || value > 1) return GPRTO_EINVAL; h->k2_path = int(value); return 0; }
  return GPRTO_EINVAL;
}

int64_t gprto_launch_count(const gprto_t* h) { return h ? h->launches.load() : 0; }

int gprto_order_after(gprto_t* h, void* stream) {
  if (!h) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  if (!h->ev_order) GPRTO_CUDA(cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming));
  GPRTO_CUDA(cudaEventRecord(h->ev_order, (cudaStream_t)stream));
  h->order_pending = true;
  return 0;
}

int gprto_normalise(gprto_t* h, int64_t B, int n, int d, const double* X, const double* Y, double* x_mean,
                    double* x_std, double* y_mean, double* y_std, double* Xn, double* Yn, void* stream) {
  if (!h || B <= 0 || n <= 0 || d <= 0 || !X || !Y || !x_mean || !x_std || !y_mean || !y_std || !Xn || !Yn) return GPRTO_EINVAL;
  if (d > CH_MAXD) return GPRTO_EUNSUPPORTED;
  DeviceGuard guard(h->device);
  normalise_kernel<<<(unsigned)B, 128, 0, (cudaStream_t)stream>>>(n, d, X, Y, x_mean, x_std, y_mean, y_std, Xn, Yn);
  h->launches++;
  return launch_status();
}

int gprto_cov_se_ard(gprto_t* h, int64_t B, int S, int n, int d, const double* Xn, const double* theta, int add_noise,
                     double jitter, double* K, void* stream) {
  if (!h || B <= 0 || S <= 0 || n <= 0 || d <= 0 || !Xn || !theta || !K) return GPRTO_EINVAL;
  if (d > CH_MAXD || size_t(n) * d * sizeof(double) > 200 * 1024 || B * int64_t(S) > kMaxGridX) return GPRTO_EUNSUPPORTED;
  DeviceGuard guard(h->device);
  const bool tiled = cov_smem_bytes(n, d, true) <= 48 * 1024;      // n <= ~76: at least 4 CTAs per SM keep their tile
  const size_t smem = cov_smem_bytes(n, d, tiled);
  int rc = tiled ? set_smem(cov_kernel<true>, smem) : set_smem(cov_kernel<false>, smem);
  if (rc) return rc;
  if (tiled) {
    cudaFuncSetAttribute(cov_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cov_kernel<true><<<(unsigned)(B * S), CH_THREADS, smem, (cudaStream_t)stream>>>(S, n, d, Xn, theta, add_noise, jitter, h->cov_kernel_kind, K);
  } else {
    cov_kernel<false><<<(unsigned)(B * S), CH_THREADS, smem, (cudaStream_t)stream>>>(S, n, d, Xn, theta, add_noise, jitter, h->cov_kernel_kind, K);
  }
  h->launches++;
  return launch_status();
}

int gprto_nlml(gprto_t* h, int64_t B, int S, int n, int d, const double* Xn, const double* Yn, const double* theta,
               double jitter, double* nll, int32_t* status, void* stream) {
  if (!h || B <= 0 || S <= 0 || n <= 0 || d <= 0 || !Xn || !Yn || !theta || !nll || !status) return GPRTO_EINVAL;
  if (d > CH_MAXD) return GPRTO_EUNSUPPORTED;
  DeviceGuard guard(h->device);
  return nlml_device(h, B, S, n, d, Xn, Yn, theta, jitter, nll, status, (cudaStream_t)stream, nullptr);
}

int gprto_nlml_indexed_host(gprto_t* h, int64_t G, int64_t k, int n, int d, const double* Xn, const double* Yn,
                            const int32_t* gp_index_host, const double* theta_host, double jitter, double* nll_host,
                            int32_t* status_host) {
  if (!h || G <= 0 || k <= 0 || n <= 0 || d <= 0 || !Xn || !Yn || !gp_index_host || !theta_host || !nll_host || !status_host)
    return GPRTO_EINVAL;
  if (d > CH_MAXD) return GPRTO_EUNSUPPORTED;
  for (int64_t i = 0; i < k; ++i)
    if (gp_index_host[i] < 0 || gp_index_host[i] >= G) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  // arena layout (both pinned and device): theta[k,d+2] | nll[k] | status[k] (int32) | index[k] (int32)
  const size_t o_nll = size_t(k) * (d + 2) * 8, o_st = o_nll + size_t(k) * 8, o_idx = o_st + size_t(k) * 4, total = o_idx + size_t(k) * 4;
  if (total > h->small_cap) {
    if (h->small_pin) cudaFreeHost(h->small_pin);
    if (h->small_dev) cudaFree(h->small_dev);
    h->small_pin = nullptr; h->small_dev = nullptr; h->small_cap = 0;
    const size_t cap = std::max<size_t>(total * 2, 1 << 16);
    GPRTO_CUDA(cudaMallocHost(&h->small_pin, cap));
    GPRTO_CUDA(cudaMalloc(&h->small_dev, cap));
    h->small_cap = cap;
  }
  if (!h->small_stream) GPRTO_CUDA(cudaStreamCreateWithFlags(&h->small_stream, cudaStreamNonBlocking));
  cudaStream_t s = h->small_stream;
  if (h->order_pending) { GPRTO_CUDA(cudaStreamWaitEvent(s, h->ev_order, 0)); h->order_pending = false; }
  memcpy(h->small_pin, theta_host, o_nll);
  memcpy(h->small_pin + o_idx, gp_index_host, size_t(k) * 4);
  GPRTO_CUDA(cudaMemcpyAsync(h->small_dev, h->small_pin, o_nll, cudaMemcpyHostToDevice, s));
  GPRTO_CUDA(cudaMemcpyAsync(h->small_dev + o_idx, h->small_pin + o_idx, size_t(k) * 4, cudaMemcpyHostToDevice, s));
  int rc = nlml_device(h, k, 1, n, d, Xn, Yn, reinterpret_cast<const double*>(h->small_dev), jitter,
                       reinterpret_cast<double*>(h->small_dev + o_nll), reinterpret_cast<int32_t*>(h->small_dev + o_st), s,
                       reinterpret_cast<const int32_t*>(h->small_dev + o_idx));
  if (rc) { cudaStreamSynchronize(s); return rc; }
  GPRTO_CUDA(cudaMemcpyAsync(h->small_pin + o_nll, h->small_dev + o_nll, size_t(k) * 12, cudaMemcpyDeviceToHost, s));
  GPRTO_CUDA(cudaStreamSynchronize(s));
  memcpy(nll_host, h->small_pin + o_nll, size_t(k) * 8);
  memcpy(status_host, h->small_pin + o_st, size_t(k) * 4);
  return 0;
}

int gprto_argmin(gprto_t* h, int64_t B, int64_t S, const double* val, double* best_val, int64_t* best_idx, void* stream) {
  if (!h || B <= 0 || S <= 0 || !val || !best_val || !best_idx) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  const int threads = 128;
  const int64_t blocks = (B * 32 + threads - 1) / threads;
  argmin_rows_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(B, S, val, nullptr, best_val, best_idx);
  h->launches++;
  return launch_status();
}

int gprto_factor(gprto_t* h, int64_t B, int n, int d, const double* Xn, const double* Yn, const double* theta_opt,
                 double jitter, double* invK, double* alpha, double* logdet, int32_t* status, void* stream) {
  if (!h || B <= 0 || n <= 0 || d <= 0 || !Xn || !Yn || !theta_opt || !invK || !alpha || !status) return GPRTO_EINVAL;
  if (d > CH_MAXD) return GPRTO_EUNSUPPORTED;
  if (B > kMaxGridX) return GPRTO_EUNSUPPORTED;
  DeviceGuard guard(h->device);
  if (h->k2_path == 0) {                    // register-tile / DMMA kernel for 32 < n <= 128
    int rc = launch_factor_la(B, n, d, Xn, Yn, theta_opt, jitter, invK, alpha, logdet, status, (cudaStream_t)stream);
    if (rc != GPRTO_EUNSUPPORTED) { h->launches++; return rc; }
  }
  const size_t smem = ch_smem_bytes(n, d);
  if (smem > 227 * 1024) return GPRTO_EUNSUPPORTED;
  int rc = set_smem(factor_generic_kernel, smem);
  if (rc) return rc;
  factor_generic_kernel<<<(unsigned)B, CH_THREADS, smem, (cudaStream_t)stream>>>(n, d, Xn, Yn, theta_opt, jitter, invK, alpha,
                                                                             logdet, status);
  h->launches++;
  return launch_status();
}

int gprto_posterior_acq(gprto_t* h, int64_t B, int n, int d, int64_t m, const double* Xn, const double* invK,
                        const double* alpha, const double* theta_opt, const double* x_mean, const double* x_std,
                        const double* y_mean, const double* y_std, const double* cand, const double* prior,
                        const double* f_best, int acq, double* mean, double* var, double* score, double* best_score,
                        int64_t* best_idx, void* stream) {
  if (!h || B <= 0 || n <= 0 || d <= 0 || m <= 0 || !Xn || !invK || !alpha || !theta_opt || !x_mean || !x_std || !y_mean ||
      !y_std || !cand)
    return GPRTO_EINVAL;
  if (acq < 1 || acq > 3) return GPRTO_EINVAL;
  if (acq == GPRTO_ACQ_EI && !f_best) return GPRTO_EINVAL;
  if ((best_score == nullptr) != (best_idx == nullptr)) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  return posterior_device(h, B, n, d, m, Xn, invK, alpha, theta_opt, x_mean, x_std, y_mean, y_std, cand, prior, f_best, acq,
                          mean, var, score, best_score, best_idx, (cudaStream_t)stream);
}

int gprto_posterior_acq_host(gprto_t* h, int64_t B, int n, int d, int64_t m, const double* Xn, const double* invK,
                             const double* alpha, const double* theta_opt, const double* x_mean, const double* x_std,
                             const double* y_mean, const double* y_std, const double* cand_host, const double* prior_host,
                             const double* f_best, int acq, double* mean_host, double* var_host, double* score_host,
                             double* best_score_host, int64_t* best_idx_host) {
  if (!h || B <= 0 || n <= 0 || d <= 0 || m <= 0 || !Xn || !invK || !alpha || !theta_opt || !x_mean || !x_std || !y_mean ||
      !y_std || !cand_host)
    return GPRTO_EINVAL;
  if (acq < 1 || acq > 3) return GPRTO_EINVAL;
  if (acq == GPRTO_ACQ_EI && !f_best) return GPRTO_EINVAL;
  if ((best_score_host == nullptr) != (best_idx_host == nullptr)) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  // GP-chunks of ~64 MiB of candidates, double-buffered: stream 0 = H2D, 1 = compute, 2 = D2H.
  const size_t in_per_gp = size_t(m) * (d + (prior_host ? 1 : 0)) * sizeof(double);
  const int n_out = (mean_host ? 1 : 0) + (var_host ? 1 : 0) + (score_host ? 1 : 0);
  const size_t out_per_gp = size_t(m) * n_out * sizeof(double) + (best_score_host ? 16 : 0);
  int64_t gps = std::max<int64_t>(1, int64_t((size_t(64) << 20) / std::max(in_per_gp, size_t(1))));
  gps = std::min<int64_t>(gps, B);
  int rc = ensure_stage(h, in_per_gp * gps, std::max(out_per_gp, size_t(16)) * gps);
  if (rc) return rc;
  if (h->order_pending) { GPRTO_CUDA(cudaStreamWaitEvent(h->st[1], h->ev_order, 0)); h->order_pending = false; }
  // on any error below: let the copies already in flight into the caller's host buffers finish before returning
  auto drain = [h](int code) { for (auto& st : h->st) cudaStreamSynchronize(st); return code; };
#define GPRTO_CUDA_DRAIN(call)                                   \
  do {                                                           \
    cudaError_t e_ = (call);                                     \
    if (e_ != cudaSuccess) return drain(-static_cast<int>(e_));  \
  } while (0)
  // Host buffers are used in place; they are transferred at full speed only if the caller pinned them
  // (cudaHostRegister / torch pin_memory) - pageable memory still works, through the driver's staging.
  int64_t chunk_i = 0;
  for (int64_t b0 = 0; b0 < B; b0 += gps, ++chunk_i) {
    const int slot = int(chunk_i & 1);
    const int64_t bn = std::min<int64_t>(gps, B - b0);
    double* din = h->dev_in[slot];
    double* dprior = prior_host ? din + size_t(bn) * m * d : nullptr;
    double* dout = h->dev_out[slot];
    if (chunk_i >= 2) GPRTO_CUDA_DRAIN(cudaStreamWaitEvent(h->st[0], h->ev_k[slot], 0));    // kernel that read this slot is done
    GPRTO_CUDA_DRAIN(cudaMemcpyAsync(din, cand_host + size_t(b0) * m * d, size_t(bn) * m * d * sizeof(double), cudaMemcpyHostToDevice, h->st[0]));
    if (prior_host)
      GPRTO_CUDA_DRAIN(cudaMemcpyAsync(dprior, prior_host + size_t(b0) * m, size_t(bn) * m * sizeof(double), cudaMemcpyHostToDevice, h->st[0]));
    GPRTO_CUDA_DRAIN(cudaEventRecord(h->ev_h2d[slot], h->st[0]));
    GPRTO_CUDA_DRAIN(cudaStreamWaitEvent(h->st[1], h->ev_h2d[slot], 0));
    if (chunk_i >= 2) GPRTO_CUDA_DRAIN(cudaStreamWaitEvent(h->st[1], h->ev_d2h[slot], 0));  // outputs of this slot were copied out
    double* dmean = nullptr; double* dvar = nullptr; double* dscore = nullptr;
    size_t off = 0;
    if (mean_host) { dmean = dout + off; off += size_t(bn) * m; }
    if (var_host) { dvar = dout + off; off += size_t(bn) * m; }
    if (score_host) { dscore = dout + off; off += size_t(bn) * m; }
    double* dbs = nullptr; int64_t* dbi = nullptr;
    if (best_score_host) { dbs = dout + off; dbi = reinterpret_cast<int64_t*>(dout + off + bn); }
    rc = posterior_device(h, bn, n, d, m, Xn + b0 * n * d, invK + b0 * int64_t(n) * n, alpha + b0 * n, theta_opt + b0 * (d + 2),
                          x_mean + b0 * d, x_std + b0 * d, y_mean + b0, y_std + b0, din, dprior, f_best ? f_best + b0 : nullptr, acq,
                          dmean, dvar, dscore, dbs, dbi, h->st[1]);
    if (rc) return drain(rc);
    GPRTO_CUDA_DRAIN(cudaEventRecord(h->ev_k[slot], h->st[1]));
    GPRTO_CUDA_DRAIN(cudaStreamWaitEvent(h->st[2], h->ev_k[slot], 0));
    if (mean_host) GPRTO_CUDA_DRAIN(cudaMemcpyAsync(mean_host + size_t(b0) * m, dmean, size_t(bn) * m * sizeof(double), cudaMemcpyDeviceToHost, h->st[2]));
    if (var_host) GPRTO_CUDA_DRAIN(cudaMemcpyAsync(var_host + size_t(b0) * m, dvar, size_t(bn) * m * sizeof(double), cudaMemcpyDeviceToHost, h->st[2]));
    if (score_host) GPRTO_CUDA_DRAIN(cudaMemcpyAsync(score_host + size_t(b0) * m, dscore, size_t(bn) * m * sizeof(double), cudaMemcpyDeviceToHost, h->st[2]));
    if (best_score_host) {
      GPRTO_CUDA_DRAIN(cudaMemcpyAsync(best_score_host + b0, dbs, size_t(bn) * sizeof(double), cudaMemcpyDeviceToHost, h->st[2]));
      GPRTO_CUDA_DRAIN(cudaMemcpyAsync(best_idx_host + b0, dbi, size_t(bn) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->st[2]));
    }
    GPRTO_CUDA_DRAIN(cudaEventRecord(h->ev_d2h[slot], h->st[2]));
  }
  GPRTO_CUDA_DRAIN(cudaStreamSynchronize(h->st[2]));
  GPRTO_CUDA_DRAIN(cudaStreamSynchronize(h->st[1]));
  GPRTO_CUDA_DRAIN(cudaStreamSynchronize(h->st[0]));
  return 0;
#undef GPRTO_CUDA_DRAIN
}

int gprto_sample_candidates(gprto_t* h, int64_t B, int d, int64_t m, const double* centre, const double* radius, int shape,
                            uint64_t seed, const int64_t* idx, double* out, void* stream) {
  if (!h || B <= 0 || d <= 0 || m <= 0 || !centre || !radius || !out || shape < 0 || shape > 2) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  return dispatch_sample(h, B, d, m, centre, radius, shape, seed, idx, out, (cudaStream_t)stream);
}

int gprto_posterior_acq_sampled(gprto_t* h, int64_t B, int n, int d, int64_t m, const double* Xn, const double* invK,
                                const double* alpha, const double* theta_opt, const double* x_mean, const double* x_std,
                                const double* y_mean, const double* y_std, const double* centre, const double* radius, int shape,
                                uint64_t seed, const double* f_best, int acq, double* best_score, int64_t* best_idx, double* best_x,
                                void* stream) {
  if (!h || B <= 0 || n <= 0 || d <= 0 || m <= 0 || !Xn || !invK || !alpha || !theta_opt || !x_mean || !x_std || !y_mean || !y_std ||
      !centre || !radius || !best_score || !best_idx)
    return GPRTO_EINVAL;
  if (acq < 1 || acq > 3 || shape < 0 || shape > 2) return GPRTO_EINVAL;
  if (acq == GPRTO_ACQ_EI && !f_best) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = posterior_device(h, B, n, d, m, Xn, invK, alpha, theta_opt, x_mean, x_std, y_mean, y_std, nullptr, nullptr, f_best, acq,
                            nullptr, nullptr, nullptr, best_score, best_idx, s, centre, radius, shape, seed);
  if (rc || !best_x) return rc;
  return dispatch_sample(h, B, d, m, centre, radius, shape, seed, best_idx, best_x, s);
}

int gprto_pack_best(gprto_t* h, int64_t count, const double* best_score, const int64_t* best_idx, int64_t idx_offset,
                    void* out_records, void* stream) {
  if (!h || count <= 0 || !best_score || !best_idx || !out_records) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  pack_best_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(count, best_score, best_idx, idx_offset,
                                                                                 static_cast<double*>(out_records));
  h->launches++;
  return launch_status();
}

int gprto_first_failure(gprto_t* h, const int32_t* status, int64_t count, void* stream, int64_t* first) {
  if (!h || !status || count <= 0 || !first) return GPRTO_EINVAL;
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned long long init = ~0ULL;
  GPRTO_CUDA(cudaMemcpyAsync(h->first_dev, &init, sizeof(init), cudaMemcpyHostToDevice, s));
  first_failure_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(count, status, h->first_dev);
  h->launches++;
  unsigned long long out = 0;
  GPRTO_CUDA(cudaMemcpyAsync(&out, h->first_dev, sizeof(out), cudaMemcpyDeviceToHost, s));
  GPRTO_CUDA(cudaStreamSynchronize(s));
  *first = out == ~0ULL ? -1 : int64_t(out);
  return 0;
}

}  // extern "C"
