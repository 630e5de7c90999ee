/*
 * gprto.h - C ABI of the B200-native GP + acquisition hot path of GP-RTO-EI.
 *
 * The reference (panos108/GP-RTO-EI) has no FFI layer: its hot path is the plain Python class
 * GP_model (sub_uts/utilities_2.py:23-278) plus the scoring function ITR_GP_RTO.GP_obj_f
 * (sub_uts/utilities_2.py:477-507).  Each entry point below names the reference lines it replaces.
 * A maintainer binds this library with ctypes; the shipped binding is
 * gp_rto_ei_b200/_lib.py and the drop-in class is gp_rto_ei_b200/sub_uts/utilities_2.py.
 *
 * Conventions
 *   - All arrays are row-major (C order) FP64 unless stated; indices are int64, status words int32.
 *   - Pointers are DEVICE pointers owned by the caller, except in the *_host entry points.
 *   - theta = (0.5*log W_1..d, 0.5*log sf2, 0.5*log sn2), W = squared length-scales
 *     (utilities_2.py:100-102).
 *   - Calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - Return value: 0 = ok; < 0 = error: -(cudaError_t) for CUDA failures, GPRTO_E* otherwise.
 *     No C++ exception crosses the boundary.  Numerical failure (non-positive Cholesky pivot, the
 *     reference's uncaught numpy.linalg.LinAlgError at utilities_2.py:107) is reported per problem in
 *     the `status` arrays with the LAPACK dpotrf convention (0 ok, k > 0: leading minor of order k is
 *     not positive definite) and summarised by gprto_first_failure().
 *   - There is NO CPU fallback: every entry point fails with GPRTO_ENODEVICE if no sm_100 device exists.
 */
#ifndef GPRTO_H_
#define GPRTO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPRTO_EINVAL       (-10001)  /* bad argument (NULL pointer, non-positive size, unknown acq)   */
#define GPRTO_EUNSUPPORTED (-10002)  /* size outside what the kernels are built for (see gprto_limits) */
#define GPRTO_ENODEVICE    (-10003)  /* no CUDA device / not compute capability 10.x                 */
#define GPRTO_ENOMEM       (-10004)

#define GPRTO_ACQ_MEAN 1   /* prior + mean                      (utilities_2.py:485-487) */
#define GPRTO_ACQ_LCB  2   /* prior + mean - 3*sqrt(var)        (utilities_2.py:489-491) */
#define GPRTO_ACQ_EI   3   /* -(sigma*phi(Z) + Delta*Phi(Z))    (utilities_2.py:492-504) */

#define GPRTO_K4_AUTO    0  /* pick the DMMA path when n <= 128 and d <= 8, else the generic path */
#define GPRTO_K4_GENERIC 1  /* force the generic (one thread per candidate, vector DFMA) path      */
#define GPRTO_K4_DMMA    2  /* force the DMMA path (GPRTO_EUNSUPPORTED outside its limits)         */

typedef struct gprto_handle gprto_t;

/* Library / device life-cycle.  A handle binds one device and owns only scratch workspace. */
int  gprto_create(int device, gprto_t** out);
int  gprto_destroy(gprto_t* h);
const char* gprto_version(void);
const char* gprto_error_string(int code);
/* limits[0] = max n of the DMMA posterior path, [1] = max n of the register-resident NLML path,
 * [2] = max n of gprto_factor / gprto_nlml (matrix in shared memory; the generic posterior path alone goes to n = 512),
 * [3] = max d. */
int  gprto_limits(int64_t limits[4]);
int  gprto_set_option(gprto_t* h, const char* name, int64_t value);   /* "k4_path": GPRTO_K4_*; "k3_path": 0 auto, 1 generic (shared memory), 2 register kernel without look-ahead,
                                                                         3 look-ahead kernel with tensor-core panel solve ("lb"), 4 round-1 look-ahead kernel, 5 "lc" (= auto for n > 32);
                                                                         "k2_path": 0 auto (register-tile / DMMA kernel for 32 < n <= 128), 1 generic (shared memory);
                                                                         "cov_kernel": kernel of gprto_cov_se_ard only: 0 SE-ARD, 1 Matern-3/2, 2 Matern-5/2 */
/* Number of kernels this library launched through the handle since creation (bench.py's gpu_launches). */
int64_t gprto_launch_count(const gprto_t* h);

/* Stream ordering of the synchronous *_host entry points.  They run on streams PRIVATE to the handle (created
 * non-blocking, so they do not synchronise with the legacy default stream); their device-resident inputs (Xn, Yn, invK,
 * alpha, ...) are usually produced by work the caller enqueued on some other stream.  gprto_order_after(h, s) makes the
 * NEXT *_host call wait (on the device, cudaStreamWaitEvent) for everything enqueued on stream `s` up to now.  Contract:
 * either call it with the producing stream before a *_host call, or make sure that stream is idle. */
int gprto_order_after(gprto_t* h, void* stream);

/* Normalisation of the data set (replaces utilities_2.py:38-40): column mean / population std (ddof = 0).
 * X[B,n,d], Y[B,n] -> x_mean[B,d], x_std[B,d], y_mean[B], y_std[B], Xn[B,n,d], Yn[B,n]. */
int gprto_normalise(gprto_t* h, int64_t B, int n, int d, const double* X, const double* Y,
                    double* x_mean, double* x_std, double* y_mean, double* y_std,
                    double* Xn, double* Yn, void* stream);

/* K1 - SE-ARD Gram matrix (replaces GP_model.Cov_mat, RBF branch, utilities_2.py:50-59, and the diagonal
 * terms of :105 / :173).  K[b,s] = sf2*exp(-0.5*dist) + (add_noise ? sn2 : 0)*I + jitter*I.  With option
 * "cov_kernel" = 1 / 2 the Matern-3/2 / -5/2 forms of :61-66 are produced instead (never requested by the reference's
 * drivers and not used by its inference, which is hard-wired to the SE kernel, :83).
 * Xn[B,n,d], theta[B,S,d+2] -> K[B,S,n,n] (full symmetric matrix). */
int gprto_cov_se_ard(gprto_t* h, int64_t B, int S, int n, int d, const double* Xn, const double* theta,
                     int add_noise, double jitter, double* K, void* stream);

/* K3 - batched NLML over hyper-parameter multistarts (replaces GP_model.negative_loglikelihood,
 * utilities_2.py:92-114): nll[b,s] = y' (K + (sn2+jitter) I)^-1 y + log det(.)  via Cholesky; the reference
 * passes jitter = 1e-8.  status[b,s] follows dpotrf (see above); nll is NaN where status != 0.
 * Xn[B,n,d], Yn[B,n], theta[B,S,d+2] -> nll[B,S], status[B,S]. */
int gprto_nlml(gprto_t* h, int64_t B, int S, int n, int d, const double* Xn, const double* Yn,
               const double* theta, double jitter, double* nll, int32_t* status, void* stream);

/* Synchronous small-batch form for host-driven optimisers (the lock-step SLSQP rounds of the drop-in class): k problems,
 * problem i uses data set gp_index_host[i] of the G device-resident ones (Xn[G,n,d], Yn[G,n]) with the hyper-parameters
 * theta_host[i,:]; theta, index and results travel through one pinned arena and one stream owned by the handle; returns
 * after nll_host[k] / status_host[k] are filled.  Same kernels and results as gprto_nlml. */
int gprto_nlml_indexed_host(gprto_t* h, int64_t G, int64_t k, int n, int d, const double* Xn, const double* Yn,
                            const int32_t* gp_index_host, const double* theta_host, double jitter,
                            double* nll_host, int32_t* status_host);

/* Row-wise arg-min with numpy.argmin semantics (first minimum wins; a NaN wins over everything;
 * replaces utilities_2.py:166 and :991).  val[B,S] -> best_val[B], best_idx[B]. */
int gprto_argmin(gprto_t* h, int64_t B, int64_t S, const double* val, double* best_val, int64_t* best_idx,
                 void* stream);

/* K2 - final factor at the chosen hyper-parameters (replaces utilities_2.py:168-175):
 * Kopt = K + sn2*I (+ jitter*I; the reference uses jitter = 0 here), Cholesky, explicit inverse
 * invK[B,n,n] (the reference's `invKopt`), alpha = invK*y[B,n], logdet[B] (may be NULL), status[B]. */
int gprto_factor(gprto_t* h, int64_t B, int n, int d, const double* Xn, const double* Yn,
                 const double* theta_opt, double jitter, double* invK, double* alpha, double* logdet,
                 int32_t* status, void* stream);

/* K4 - fused cross-covariance -> posterior mean / variance -> acquisition score -> per-GP arg-min
 * (replaces GP_model.calc_cov_sample :74-85, GP_model.GP_inference_np :235-278 and ITR_GP_RTO.GP_obj_f
 * :477-507 for m candidates of each of B GPs).
 *   inputs : Xn[B,n,d], invK[B,n,n], alpha[B,n], theta_opt[B,d+2], x_mean[B,d], x_std[B,d], y_mean[B], y_std[B],
 *            cand[B,m,d] (raw, un-normalised inputs x = xk + d), prior[B,m] or NULL (= obj_model(x)),
 *            f_best[B] (obj_min; only read for GPRTO_ACQ_EI, may be NULL otherwise), acq in GPRTO_ACQ_*.
 *   outputs: mean[B,m], var[B,m] (de-normalised, var clamped at 0 as :268), score[B,m]  - each may be NULL;
 *            best_score[B], best_idx[B] (arg-min of score over the m candidates, numpy.argmin semantics) -
 *            both NULL or both non-NULL. */
int gprto_posterior_acq(gprto_t* h, int64_t B, int n, int d, int64_t m,
                        const double* Xn, const double* invK, const double* alpha, const double* theta_opt,
                        const double* x_mean, const double* x_std, const double* y_mean, const double* y_std,
                        const double* cand, const double* prior, const double* f_best, int acq,
                        double* mean, double* var, double* score, double* best_score, int64_t* best_idx,
                        void* stream);

/* Same computation with HOST buffers for the per-candidate streams (cand, prior in; mean, var, score,
 * best_* out); the GP-side arrays (Xn ... y_std, f_best) stay device-resident.  Candidates are processed in
 * GP-chunks of ~64 MiB, double-buffered in device staging buffers on three streams so H2D, compute and D2H overlap;
 * the host buffers are read / written IN PLACE (full PCIe speed needs them pinned by the caller - cudaHostRegister or
 * torch pin_memory -, pageable memory works through the driver's own staging); returns after the last copy has
 * completed (also on error: copies in flight are drained first).  This is the end-to-end entry bench.py times as `e2e`.
 * See gprto_order_after for the stream-ordering contract. */
int gprto_posterior_acq_host(gprto_t* h, int64_t B, int n, int d, int64_t m,
                             const double* Xn, const double* invK, const double* alpha, const double* theta_opt,
                             const double* x_mean, const double* x_std, const double* y_mean, const double* y_std,
                             const double* cand_host, const double* prior_host, const double* f_best, int acq,
                             double* mean_host, double* var_host, double* score_host,
                             double* best_score_host, int64_t* best_idx_host);

/* On-device trust-region candidates (SURVEY.md par. 8f-4; the reference draws its start points on the host,
 * utilities_2.py:678-730).  Candidate c of GP b is a pure function of (seed, b, c):
 *   cand[b,c,:] = centre[b,:] + radius[b] * u,   u uniform in [-1,1]^d (shape 0) or in the unit ball (shape 1);
 *   shape 2 = the reference's Ellipse_sampling for scaled inputs (utilities_2.py:693-730): radius is [B, d+1], per-dimension
 *   semi-axes followed by the trust-region ball radius; u uniform in the unit ball is stretched by the semi-axes and
 *   re-drawn (up to 64 times) until the point lies inside the ball (radius[b,d] <= 0: no ball test).
 * gprto_sample_candidates materialises them: out[B,m,d], or, with idx[B] != NULL, only candidate idx[b] of each GP
 * (out[B,d]).  gprto_posterior_acq_sampled is K4 with the SAME candidates generated inside the kernel (no candidate
 * array in memory at all); it returns the per-GP arg-min (best_score[B], best_idx[B]) and, if best_x != NULL, the
 * coordinates best_x[B,d] of the winning candidates.  prior is not available in this mode. */
int gprto_sample_candidates(gprto_t* h, int64_t B, int d, int64_t m, const double* centre, const double* radius, int shape,
                            uint64_t seed, const int64_t* idx, double* out, void* stream);
int gprto_posterior_acq_sampled(gprto_t* h, int64_t B, int n, int d, int64_t m,
                                const double* Xn, const double* invK, const double* alpha, const double* theta_opt,
                                const double* x_mean, const double* x_std, const double* y_mean, const double* y_std,
                                const double* centre, const double* radius, int shape, uint64_t seed,
                                const double* f_best, int acq, double* best_score, int64_t* best_idx, double* best_x,
                                void* stream);

/* Pack (score, index) pairs as 16-byte records {double score; int64 idx + idx_offset} into `out[count]`:
 * the send buffer of the final NCCL all-gather. */
int gprto_pack_best(gprto_t* h, int64_t count, const double* best_score, const int64_t* best_idx,
                    int64_t idx_offset, void* out_records, void* stream);

/* Synchronise `stream` and scan status[count]: *first = -1 if all zero, else the flat index of the first
 * non-zero word (the shim raises numpy.linalg.LinAlgError there, mirroring utilities_2.py:107). */
int gprto_first_failure(gprto_t* h, const int32_t* status, int64_t count, void* stream, int64_t* first);

#ifdef __cplusplus
}
#endif
#endif /* GPRTO_H_ */
