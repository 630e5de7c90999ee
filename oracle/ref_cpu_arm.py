"""The reference's OWN functions as the timed CPU arm.  TEST / BENCH INFRASTRUCTURE ONLY.

bench.py's `cpu_baseline` and `--impl reference` legs time these wrappers on the host cores: each worker imports the
UNMODIFIED reference module (oracle/ref_harness.py: /root/reference in the build container, the byte-for-byte staged
copy baseline/_ref on the GPU box) and loops its own methods exactly as the reference's drivers call them -

  posterior + EI : ITR_GP_RTO.GP_obj_f(d, GP, xk)  ->  GP_model.GP_inference_np(xk + d)   (utilities_2.py:477-507, 235-278)
  NLML           : GP_model.negative_loglikelihood(hyper, X_norm, Y_norm[:, i])             (utilities_2.py:92-114)

one candidate / one hyper-parameter vector per call.  The GP objects are built with every line of GP_model.__init__
(:31-40) and the final factor (:168-175, through the reference's own Cov_mat) but without the SLSQP search, whose
result (theta) is an input of the synthetic workloads.  When neither copy of the reference exists the callers fall
back to the restatement in oracle/gp_oracle.py and say so (kind "port").
"""
import os
import time

import numpy as np

from . import ref_harness


def available():
    return ref_harness.available()


def _ref():
    return ref_harness.load('utilities_2')


def bare_gp(X, Y, theta, var_out=True):
    ref = _ref()
    gp = object.__new__(ref.GP_model)
    gp.X, gp.Y, gp.kernel = X, Y, 'RBF'
    gp.n_point, gp.nx_dim = X.shape
    gp.ny_dim = Y.shape[1]
    gp.multi_hyper, gp.var_out, gp.GP_casadi = 1, var_out, False
    gp.X_mean, gp.X_std = np.mean(X, axis=0), np.std(X, axis=0)
    gp.Y_mean, gp.Y_std = np.mean(Y, axis=0), np.std(Y, axis=0)
    gp.X_norm, gp.Y_norm = (X - gp.X_mean) / gp.X_std, (Y - gp.Y_mean) / gp.Y_std
    d = gp.nx_dim
    gp.hypopt = np.asarray(theta, dtype=float).reshape(d + 2, 1).copy()
    ell, sf2, sn2 = np.exp(2. * gp.hypopt[:d, 0]), np.exp(2. * gp.hypopt[d, 0]), np.exp(2. * gp.hypopt[d + 1, 0])
    Kopt = gp.Cov_mat('RBF', gp.X_norm, ell, sf2) + sn2 * np.eye(gp.n_point)
    gp.invKopt = [np.linalg.solve(Kopt, np.eye(gp.n_point))]
    return gp


def _no_prior(x):
    return 0.


def scorer(acq, f_best, obj_model=_no_prior):
    ref = _ref()
    rto = object.__new__(ref.ITR_GP_RTO)
    rto.obj, rto.obj_model, rto.obj_min = acq, obj_model, f_best
    return rto


def ei_scores(gp, cand, f_best, acq=3):
    """score[c] = GP_obj_f(cand[c], gp, xk=0): the reference's one-candidate-at-a-time call chain."""
    rto = scorer(acq, f_best)
    xk = np.zeros(cand.shape[1])
    out = np.empty(cand.shape[0])
    for c in range(cand.shape[0]):
        out[c] = np.asarray(rto.GP_obj_f(cand[c], gp, xk)).reshape(-1)[0]
    return out


def nlml_values(gp, thetas):
    out = np.empty(len(thetas))
    y = gp.Y_norm[:, 0]
    for s, th in enumerate(thetas):
        try:
            out[s] = float(gp.negative_loglikelihood(th, gp.X_norm, y))
        except np.linalg.LinAlgError:
            out[s] = np.nan
    return out


# ------------------------------------------------------------------------------------------ pool workers
def cfg3_worker(args):
    """One process: n_gp cfg-3 GPs x m candidates through GP_obj_f (EI).  Returns (seconds, checksum)."""
    seed, n_gp, n, d, m = args
    import warnings
    from threadpoolctl import threadpool_limits
    from gp_rto_ei_b200 import workloads
    warnings.simplefilter('ignore')                      # the reference's ndim>0 -> scalar DeprecationWarnings
    with threadpool_limits(1):
        w = workloads.cfg3_numpy(seed=seed, B=n_gp, n=n, d=d, m=m)
        gps = [bare_gp(np.ascontiguousarray(w['X'][b]), w['y'][b].reshape(-1, 1).copy(), w['theta'][b]) for b in range(n_gp)]
        t0 = time.perf_counter()
        acc = 0.0
        for b in range(n_gp):
            acc += float(ei_scores(gps[b], w['cand'][b], float(w['f_best'][b])).min())
        dt = time.perf_counter() - t0
    return dt, acc


def cfg4_worker(args):
    """One process: n_gp cfg-4 GPs x S Sobol starts through negative_loglikelihood.  Returns (seconds, #finite)."""
    seed, n_gp, n, d, S = args
    import warnings
    from threadpoolctl import threadpool_limits
    from gp_rto_ei_b200 import workloads
    warnings.simplefilter('ignore')
    with threadpool_limits(1):
        w = workloads.cfg4_numpy(seed=seed, B=n_gp, n=n, d=d, S=S)
        gps = [bare_gp(np.ascontiguousarray(w['X'][b]), w['y'][b].reshape(-1, 1).copy(), w['theta'][b, 0]) for b in range(n_gp)]
        t0 = time.perf_counter()
        ok = 0
        for b in range(n_gp):
            v = nlml_values(gps[b], w['theta'][b])
            ok += int(np.isfinite(v).sum())
        dt = time.perf_counter() - t0
    return dt, ok


def make_pool(cores=None):
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    return mp.get_context('spawn').Pool(cores)


def run_pool(worker, jobs, cores=None, pool=None):
    """Map `jobs` over `cores` worker processes (one job each); a caller-owned pool is reused across steps."""
    t0 = time.perf_counter()
    if pool is not None:
        res = pool.map(worker, jobs, chunksize=1)
    else:
        with make_pool(cores) as p:
            res = p.map(worker, jobs, chunksize=1)
    return res, time.perf_counter() - t0
