"""Drop-in mirror of the reference module ``sub_uts/utilities_2.py`` on top of the B200 kernels.

Same public surface as the reference (panos108/GP-RTO-EI, sub_uts/utilities_2.py): class ``GP_model``
(:23-278: ctor, ``Cov_mat``, ``calc_cov_sample``, ``negative_loglikelihood``, ``determine_hyperparameters``,
``GP_inference_np``, attributes ``hypopt``, ``invKopt``, ``X_mean`` ...) and class ``ITR_GP_RTO`` (:338-1038:
ctor, ``RTO_routine`` and the helpers the scripts touch), so that ``run_simple/main_simple.py`` and
``run_WO/main_WO_prior_EI.py`` run unchanged with ``from sub_uts.utilities_2 import *``
(see gp_rto_ei_b200/run_main.py for the path set-up).

What is different underneath
  * every covariance / Cholesky / NLML / posterior / acquisition evaluation runs in libgprto.so (CUDA, sm_100a);
    there is no NumPy fallback - constructing a GP_model without the library or without a B200 raises;
  * the two ``scipy.optimize.minimize(method='SLSQP')`` multistart loops (:156-163 and :970-984) are advanced in
    lock step (gp_rto_ei_b200/slsqp_batch.py) so that each SLSQP round is one batched kernel launch instead of
    ``starts x (1 + dim)`` Python calls.  The SLSQP arithmetic itself is SciPy's own C routine, unchanged.
The host-side control flow (data handling, trust-region update, RNG consumption order) follows the reference
line by line where the results depend on it; those places cite the reference lines.
"""
import functools  # noqa: F401  (re-exported: the scripts star-import this module)
import random

import numpy as np
import numpy.random as rnd  # noqa: F401
import torch
from scipy.optimize import minimize  # noqa: F401

from gp_rto_ei_b200 import api, slsqp_batch
from gp_rto_ei_b200.compat import sobol_seq

__all__ = ['GP_model', 'ITR_GP_RTO', 'central_finite_diff5', 'Second_diff_fxx', 'np', 'rnd', 'random', 'minimize',
           'functools', 'sobol_seq']

_JITTER = 1e-8                       # utilities_2.py:105
_LB_MAIN, _UB_MAIN = -5.0, 5.0       # utilities_2.py:131-132
_LB_NOISE, _UB_NOISE = -8.0, -4.0


def _dev(a, handle):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=handle.device)


def _hyper_bounds(d, noise_i, y_std_i):
    """Log-hyper box of utilities_2.py:131-154; a known noise variance pins the last coordinate."""
    lb = np.array([_LB_MAIN] * (d + 1) + [_LB_NOISE])
    ub = np.array([_UB_MAIN] * (d + 1) + [_UB_NOISE])
    if noise_i is not None:
        centre = np.log(noise_i / y_std_i ** 2) * 0.5
        lb[-1], ub[-1] = centre - 0.001, centre + 0.001
    return lb, ub


def _fit_group(handle, Xn_dev, Yn_dev_rows, bounds_rows, multi_start):
    """Multistart SLSQP fit of G single-output GPs that share X (utilities_2.py:139-170), all starts of all GPs in
    lock step.  Xn_dev[n,d]; Yn_dev_rows[G,n]; bounds_rows: list of (lb, ub).  Returns theta_opt[G, d+2] and, per GP, the
    multistart record dict(localval[S], localsol[S,d+2], minindex) - the reference's `localval` / `localsol` / `minindex`."""
    n, d = Xn_dev.shape
    G = Yn_dev_rows.shape[0]
    sob = sobol_seq.i4_sobol_generate(d + 2, multi_start)
    x0s, bnds, owner = [], [], []
    for g in range(G):
        lb, ub = bounds_rows[g]
        for j in range(multi_start):
            x0s.append(lb + (ub - lb) * sob[j, :])
            bnds.append(np.stack([lb, ub], axis=1))
            owner.append(g)
    owner = np.asarray(owner)

    Xg = Xn_dev.unsqueeze(0).expand(G, n, d).contiguous()          # one copy of X per GP (they share it), device-resident

    def fun_batch(inst, P):
        # one synchronous C call per SLSQP round: theta + owner index up, NLML + status down (pinned arena in the handle)
        out, st = handle.nlml_indexed_host(Xg, Yn_dev_rows, owner[inst], P, jitter=_JITTER)
        if st.any():
            first = int(np.flatnonzero(st)[0])                           # numpy.linalg.LinAlgError as utilities_2.py:107
            raise np.linalg.LinAlgError('Matrix is not positive definite (negative_loglikelihood, problem %d of %d)' % (first, len(st)))
        return out

    res = slsqp_batch.minimize_batch(fun_batch, x0s, bnds, ftol=1e-12, maxiter=10000, inst_ids=list(range(len(x0s))))
    theta = np.empty((G, d + 2))
    info = []
    for g in range(G):
        vals = np.array([res[g * multi_start + j].fun for j in range(multi_start)])
        best = int(np.argmin(vals))                                      # first minimum wins (utilities_2.py:166)
        theta[g] = res[g * multi_start + best].x
        info.append(dict(localval=vals, localsol=np.stack([res[g * multi_start + j].x for j in range(multi_start)]), minindex=best))
    return theta, info


class GP_model:
    """Gaussian process with SE-ARD kernel, fitted by NLML multistart and evaluated on the GPU.
    Constructor arguments and attributes as the reference's GP_model (utilities_2.py:28-44)."""

    def __init__(self, X, Y, kernel, multi_hyper, var_out=True, noise=None, GP_casadi=False, _theta=None, _fit_info=None):
        self.X, self.Y, self.kernel = X, Y, kernel
        self.n_point, self.nx_dim = X.shape[0], X.shape[1]
        self.ny_dim = Y.shape[1]
        self.multi_hyper = multi_hyper
        self.var_out = var_out
        self.GP_casadi = GP_casadi
        if kernel != 'RBF':
            # the reference can build Matern Gram matrices (:61-66) but never requests them and its inference is
            # hard-wired to the SE kernel (:83); only the SE-ARD path exists on the device
            raise NotImplementedError("only kernel='RBF' (SE-ARD) is implemented on the device")
        self._h = api.default_handle()
        # normalisation on the host with the very NumPy calls of :38-40 so the public attributes are bit-identical
        self.X_mean, self.X_std = np.mean(X, axis=0), np.std(X, axis=0)
        self.Y_mean, self.Y_std = np.mean(Y, axis=0), np.std(Y, axis=0)
        self.X_norm, self.Y_norm = (X - self.X_mean) / self.X_std, (Y - self.Y_mean) / self.Y_std
        self._Xn = _dev(self.X_norm, self._h)                      # [n, d]
        self._Yn = _dev(self.Y_norm.T, self._h)                    # [ny, n]
        self._fit_info = _fit_info                                 # per output: localval / localsol / minindex of the multistart
        self.hypopt, self.invKopt = self.determine_hyperparameters(noise, _theta=_theta)
        self.meanfcn, self.varfcn = self.GP_predictor()

    # ---------------------------------------------------------------- covariance (utilities_2.py:50-85)
    def _theta_from(self, W, sf2, sn2=1.0):
        return np.concatenate([0.5 * np.log(np.asarray(W, dtype=float).reshape(-1)), [0.5 * np.log(sf2)], [0.5 * np.log(sn2)]])

    def Cov_mat(self, kernel, X_norm, W, sf2):
        kinds = {'RBF': 0, 'Matern32': 1, 'Matern52': 2}
        if kernel not in kinds:
            print('ERROR no kernel with name ', kernel)          # the reference prints and returns None (:67-68)
            return None
        Xd = self._Xn if X_norm is self.X_norm else _dev(X_norm, self._h)
        th = _dev(self._theta_from(W, sf2), self._h).reshape(1, 1, -1)
        self._h.set_option('cov_kernel', kinds[kernel])
        try:
            return self._h.cov_se_ard(Xd.unsqueeze(0).contiguous(), th)[0, 0].cpu().numpy()
        finally:
            self._h.set_option('cov_kernel', 0)

    def calc_cov_sample(self, xnorm, Xnorm, ell, sf2):
        """k(x*, X) as an (n, 1) column - one row block of the Gram matrix of [X; x*]."""
        both = np.vstack([np.asarray(Xnorm, dtype=float), np.asarray(xnorm, dtype=float).reshape(1, self.nx_dim)])
        th = _dev(self._theta_from(ell, sf2), self._h).reshape(1, 1, -1)
        K = self._h.cov_se_ard(_dev(both, self._h).unsqueeze(0), th)[0, 0]
        return K[:-1, -1:].cpu().numpy()

    # ---------------------------------------------------------------- NLML (utilities_2.py:92-114)
    def negative_loglikelihood(self, hyper, X, Y):
        Xd = (self._Xn if X is self.X_norm else _dev(X, self._h)).unsqueeze(0).contiguous()
        Yd = _dev(np.asarray(Y, dtype=float).reshape(1, -1), self._h)
        nll, st = self._h.nlml_indexed_host(Xd, Yd, np.zeros(1, dtype=np.int32), np.asarray(hyper, dtype=float).reshape(1, -1),
                                            jitter=_JITTER)
        if st[0]:
            raise np.linalg.LinAlgError('Matrix is not positive definite (leading minor of order %d)' % st[0])   # as :107
        return float(nll[0])

    # ---------------------------------------------------------------- fit (utilities_2.py:120-177)
    def determine_hyperparameters(self, noise=None, _theta=None):
        d, ny = self.nx_dim, self.ny_dim
        noise_vec = np.array([noise]).reshape((-1,))
        if _theta is None:
            bounds_rows = []
            lb = ub = None
            for i in range(ny):
                known = noise_vec[min(i, len(noise_vec) - 1)] if noise is not None else None
                if known is not None or lb is None:
                    # the reference mutates lb/ub in place, so a pinned noise bound persists for later outputs
                    lb, ub = _hyper_bounds(d, known, self.Y_std[i] if known is not None else None)
                bounds_rows.append((lb.copy(), ub.copy()))
            theta, self._fit_info = _fit_group(self._h, self._Xn, self._Yn, bounds_rows, self.multi_hyper)
        else:
            theta = np.asarray(_theta, dtype=float).reshape(ny, d + 2)
        hypopt = np.ascontiguousarray(theta.T)                      # (d+2, ny) as the reference
        # final factor: Kopt = K + sn2*I, no jitter (:173-175); invK and alpha = invK*y stay on the device too
        Xb = self._Xn.unsqueeze(0).expand(ny, self.n_point, d).contiguous()
        invK, alpha, logdet, st = self._h.factor(Xb, self._Yn, _dev(theta, self._h), jitter=0.0)
        if int(st.abs().sum()) != 0:
            # The reference inverts Kopt by LU (np.linalg.solve, :175), which survives a numerically singular but noisy Kopt
            # (near-duplicate inputs with the noise pinned around 1e-7) where a Cholesky factorisation reports a non-positive
            # pivot.  With sn2 > 0 the matrix is positive definite in exact arithmetic, so retry with a diagonal shift far below
            # the noise level and say so; an exactly singular Kopt (sn2 == 0) raises, as LAPACK does for the reference.
            sf2, sn2 = np.exp(2 * theta[:, d]), np.exp(2 * theta[:, d + 1])
            if np.all(sn2 > 0):
                for rel in (1e-12, 1e-10, 1e-8):
                    shift = float(rel * sf2.max())
                    invK, alpha, logdet, st = self._h.factor(Xb, self._Yn, _dev(theta, self._h), jitter=shift)
                    if int(st.abs().sum()) == 0:
                        import warnings
                        warnings.warn('final factor: Kopt is numerically singular; factorised with a diagonal shift of %.1e' % shift, RuntimeWarning)
                        break
        self._h.raise_on_failure(st, 'determine_hyperparameters (final factor)')
        self._gp = dict(Xn=Xb, invK=invK, alpha=alpha, theta=_dev(theta, self._h),
                        x_mean=_dev(np.tile(self.X_mean, (ny, 1)), self._h), x_std=_dev(np.tile(self.X_std, (ny, 1)), self._h),
                        y_mean=_dev(self.Y_mean, self._h), y_std=_dev(self.Y_std, self._h))
        invK_h = invK.cpu().numpy()
        return hypopt, [invK_h[i] for i in range(ny)]

    # ---------------------------------------------------------------- casadi predictor (utilities_2.py:182-232)
    def GP_predictor(self):
        """Symbolic predictor, only when casadi is importable and requested (both mains leave GP_casadi=False and
        never evaluate it); mean = k'(invK y)*stdY + meanY, var = (sf2 - k' invK k)*stdY^2 without the clamp."""
        if not self.GP_casadi:
            return None, None
        import casadi as ca
        d, n = self.nx_dim, self.n_point
        x = ca.SX.sym('x', d)
        xn = (x - ca.DM(self.X_mean)) / ca.DM(self.X_std)
        means, variances = [], []
        for i in range(self.ny_dim):
            W, sf2 = np.exp(2 * self.hypopt[:d, i]), float(np.exp(2 * self.hypopt[d, i]))
            k = ca.vertcat(*[sf2 * ca.exp(-0.5 * ca.sum1((xn - ca.DM(self.X_norm[j])) ** 2 / ca.DM(W))) for j in range(n)])
            invK = ca.DM(self.invKopt[i])
            means.append(ca.mtimes(k.T, ca.mtimes(invK, ca.DM(self.Y_norm[:, i]))))
            variances.append(sf2 - ca.mtimes(ca.mtimes(k.T, invK), k))
        mean = ca.vertcat(*means) * ca.DM(self.Y_std) + ca.DM(self.Y_mean)
        var = ca.vertcat(*variances) * ca.DM(self.Y_std) ** 2
        return ca.Function('meanfcn', [x], [mean]), ca.Function('varfcn', [x], [var])

    # ---------------------------------------------------------------- inference (utilities_2.py:235-278)
    def GP_inference_batch(self, Xq):
        """Posterior at m points at once: Xq[m, d] -> (mean[m, ny], var[m, ny]), de-normalised, var clamped at 0."""
        Xq = np.asarray(Xq, dtype=float).reshape(-1, self.nx_dim)
        # host entry of K4: one synchronous C call, no torch objects on the way (this is also the scalar path of
        # GP_inference_np, ~45 us per call)
        out = self._h.posterior_acq_host_np(self._gp, np.broadcast_to(Xq, (self.ny_dim,) + Xq.shape), None, None, api.ACQ_MEAN,
                                            want=('mean', 'var'))
        return out['mean'].T, out['var'].T

    def GP_inference_np(self, x):
        if self.GP_casadi:
            mean = self.meanfcn(x).toarray().flatten()[0]
            return (mean, self.varfcn(x).toarray().flatten()[0]) if self.var_out else mean
        mean, var = self.GP_inference_batch(np.asarray(x, dtype=float).reshape(1, self.nx_dim))
        if self.var_out:
            return mean[0], var[0]
        return mean[0].flatten()[0]


# -------------------------------------------------------------------------------------------- FD helpers (:284-335)
def central_finite_diff5(f, x):
    """Five-point central gradient with step sqrt(eps); returns an (n, 1) column."""
    h = np.sqrt(np.finfo(float).eps)
    x = np.asarray(x, dtype=float).reshape(-1)
    g = np.zeros((x.size, 1))
    for j in range(x.size):
        e = np.zeros_like(x)
        e[j] = h
        g[j] = (f(x - 2 * e) - 8 * f(x - e) + 8 * f(x + e) - f(x + 2 * e)) / (12 * h)
    return g


def Second_diff_fxx(f, x):
    """Central second differences (step 100*sqrt(eps)) -> symmetric Hessian estimate."""
    h = 1e2 * np.sqrt(np.finfo(float).eps)
    x = np.asarray(x, dtype=float).reshape(-1)
    n = x.size
    H = np.zeros((n, n))
    E = np.eye(n) * h
    for j in range(n):
        H[j, j] = (f(x + E[j]) - 2 * f(x) + f(x - E[j])) / h ** 2
        for i in range(j + 1, n):
            H[j, i] = H[i, j] = (f(x + E[j] + E[i]) - f(x + E[j] - E[i]) - f(x - E[j] + E[i]) + f(x - E[j] - E[i])) / (4 * h ** 2)
    return H


# -------------------------------------------------------------------------------------------- start-point samplers (:678-730)
def ball_sample(np_rng, py_rng, ndim, Delta0):
    """ITR_GP_RTO.Ball_sampling (utilities_2.py:678-687) on explicit generators: np_rng = numpy.random (module, global
    state) or a numpy.random.RandomState; py_rng = the stdlib random module or a random.Random - the reference draws the
    direction from NumPy's generator and the radius from the stdlib one (:682-684), in this order."""
    u = np_rng.normal(0, 1, ndim)
    length = np.sum(u ** 2) ** 0.5
    r = py_rng.random() ** (1.0 / ndim)
    return r * u / length * Delta0 * 2


def ellipse_sample(np_rng, ndim, TRmat, Delta0):
    """ITR_GP_RTO.Ellipse_sampling (utilities_2.py:693-730): 200 points uniform in the ellipsoid S = inv(TRmat) Delta0^2
    (shrunk by sqrt(0.99)), the first one inside the trust-region ball |d| <= Delta0 (TR_con >= 0, :726) is returned."""
    m_fa = 200
    S = np.linalg.inv(TRmat) * Delta0 ** 2
    pts = np_rng.normal(size=(ndim, m_fa))
    pts = pts / np.sqrt(np.sum(np.square(pts), axis=0))[None, :]
    radii = np.power(np_rng.rand(1, m_fa), 1.0 / ndim)
    ell = np.linalg.cholesky(S).T @ (radii * pts) * np.sqrt(0.99)
    for i in range(m_fa):
        d = ell[:, i].reshape(1, -1)
        if Delta0 ** 2 - d @ d.T >= 0:
            return np.array(ell[:, i]).flatten()
    raise RuntimeError('no trust-region start point found among %d samples' % m_fa)   # the reference would NameError


# -------------------------------------------------------------------------------------------- RTO driver (:338-1038)
class ITR_GP_RTO:
    """Trust-region GP real-time optimisation.  Constructor signature, attributes used by the scripts and the
    return value of RTO_routine() as the reference (utilities_2.py:344-347, 888, 1037).  Only the trust-region
    shape both mains use is implemented (TR_scaling = TR_curvature = inner_TR = False): the Broyden/BFGS shaping
    of :837-873 is experimental in the reference and its inner-TR branch cannot run there (:827 NameError)."""

    def __init__(self, obj_model, obj_system, cons_model, cons_system, x0, Delta0, Delta_max, eta0, eta1, gamma_red,
                 gamma_incr, n_iter, data, bounds, obj_setting=2, noise=None, multi_opt=5, multi_hyper=10,
                 TR_scaling=True, TR_curvature=True, store_data=True, inner_TR=False, scale_inputs=False, GP_casadi=True):
        if TR_scaling or TR_curvature or inner_TR:
            raise NotImplementedError('only the plain ball / scaled-ellipse trust region of the published runs is '
                                      'implemented (TR_scaling=TR_curvature=inner_TR=False)')
        self.GP_casadi = GP_casadi
        self.obj, self.noise = obj_setting, noise
        self.obj_model, self.n_iter = obj_model, n_iter
        self.multi_opt, self.multi_hyper = multi_opt, multi_hyper
        self.store_data, self.data, self.bounds = store_data, data, bounds
        self.x0, self.obj_system, self.cons_model = x0, obj_system, cons_model
        self.cons_system, self.TR_curvature, self.TR_scaling = cons_system, TR_curvature, TR_scaling
        self.Delta_max, self.eta0, self.eta1 = Delta_max, eta0, eta1
        self.gamma_red, self.gamma_incr = gamma_red, gamma_incr
        self.inner_TR, self.scale_inputs = inner_TR, scale_inputs
        self.TRmat = np.linalg.inv(np.diag(bounds[:, 1] - bounds[:, 0])) if scale_inputs else np.eye(len(x0))
        self.ng = len(cons_model)
        if noise is None:
            self.noise = [None] * self.ng                 # as :376-377 (the objective entry is read as noise[0])
        self.Delta0 = Delta0
        self.Xtrain, self.ytrain = self.data_handling()
        self.ndim, self.ndat = self.Xtrain.shape[1], self.Xtrain.shape[0]
        print('note: remember constraints are set as positive, so they should be set as -g(x)')

    # ---- data (utilities_2.py:391-471): plant-minus-model mismatch at the initial design, output-major order
    def data_handling(self):
        kind = self.data[0]
        if kind == 'int':
            print('- No preliminar data supplied, computing data by sobol sequence')
            return self.compute_data(self.data, np.array([]))
        if kind == 'data0':
            print('- preliminar data supplied, computing objective and constraint values')
            return self.compute_data(self.data, self.data[1])
        print('- error, data ragument ', self.data, ' is of wrong type; can be int or ')
        return None

    def compute_data(self, data, Xtrain):
        data[1] = np.array(data[1])
        sys_f = [self.obj_system] + self.cons_system
        mod_f = [self.obj_model] + self.cons_model
        if Xtrain.shape == (0,):
            box = data[1]
            unit = sobol_seq.i4_sobol_generate(box.shape[0], data[2])
            Xtrain = box[:, 0] + unit * (box[:, 1] - box[:, 0])
        Xtrain = np.array(Xtrain, dtype=float)
        ndata, ndim = Xtrain.shape
        ytrain = np.zeros((self.ng + 1, ndata))
        for ii in range(self.ng + 1):
            for i in range(ndata):
                ytrain[ii, i] = np.asarray(sys_f[ii](np.array(Xtrain[i, :])) - mod_f[ii](np.array(Xtrain[i, :]))).reshape(-1)[0]
        return Xtrain.reshape(ndata, ndim, order='F'), ytrain.reshape(self.ng + 1, ndata, order='F')

    # ---- scalar acquisition (utilities_2.py:477-507): kept for API parity; the optimiser uses the batched twin below
    def GP_obj_f(self, d, GP, xk):
        x = np.asarray(xk + d, dtype=float).reshape(1, -1)
        score = self._score_batch(GP, x)
        return score[0] if self.obj in (1, 2) else np.array([score[0]])

    def _score_batch(self, GP, Xq):
        """Acquisition score of the objective GP at the rows of Xq: K4 with the prior model evaluated on the host."""
        if self.obj not in (1, 2, 3):
            print('error, objective for GP not specified')
            return None
        h = GP._h
        prior = np.array([float(self.obj_model(x.flatten())) for x in Xq])
        cand = _dev(Xq, h).unsqueeze(0)
        fb = _dev(np.array([self.obj_min if self.obj == 3 else 0.0]), h)
        out = h.posterior_acq(GP._gp, cand, _dev(prior, h).unsqueeze(0), fb, self.obj, want=('score',))
        return out['score'][0].cpu().numpy()

    def System_Cons_check(self, x):
        return bool(np.all(np.array([c(x) for c in self.cons_system]) > 0))

    def TR_con(self, d):
        d = np.asarray(d, dtype=float)
        return self.Delta0 ** 2 - d @ d.T

    def mistmatch_con(self, xk, GP_con_i, cons_model_i, d):
        return cons_model_i((xk + d).flatten()) + GP_con_i.GP_inference_np((xk + d).flatten())

    # ---- trust-region update (utilities_2.py:637-672, 757-777)
    def Adjust_TR(self, Delta0, xk, xnew, GP_obj, i_rto):
        plant_i = self.obj_system(np.array(xk).flatten())
        plant_iplus = self.obj_system(np.array(xnew).flatten())
        m_pred, _ = GP_obj.GP_inference_batch(np.vstack([np.array(xk).flatten(), np.array(xnew).flatten()]))
        pred_k = self.obj_model(xk.flatten()) + m_pred[0, 0]
        pred_new = self.obj_model(xnew.flatten()) + m_pred[1, 0]
        rho = (plant_i - plant_iplus) / (pred_k - pred_new)
        if plant_iplus < plant_i:
            if rho >= self.eta1:
                Delta0 = min(Delta0 * self.gamma_incr, self.Delta_max)
            elif rho < self.eta0:
                Delta0 = Delta0 * self.gamma_red
        else:
            xnew = xk
            Delta0 = Delta0 * self.gamma_red
            self.backtrack_l[i_rto] = True
            print('plant_iplus<plant_i -- backtracking')
        return Delta0, xnew, xk

    def Step_constraint(self, Delta0, xk, xnew, GP_obj, i_rto):
        if not self.System_Cons_check(np.array(xnew).flatten()):
            self.backtrack_l[i_rto] = True
            print('Constraint violated -- backtracking')
            return Delta0 * self.gamma_red, xk, xk
        return self.Adjust_TR(Delta0, xk, xnew, GP_obj, i_rto)

    # ---- start points (utilities_2.py:678-730); RNG calls in the reference's order
    def Ball_sampling(self, ndim):
        return ball_sample(np.random, random, ndim, self.Delta0)

    def Ellipse_sampling(self, ndim):
        return ellipse_sample(np.random, ndim, self.TRmat, self.Delta0)

    # ---- GP construction (utilities_2.py:783-831): objective + ng constraint GPs, fitted together in lock step
    def GPs_construction(self, xk, Xtrain, ytrain, ndatplus, H_norm=0.):
        ndat = self.ndat + ndatplus
        self.obj_max = np.max(ytrain[0, :].reshape(ndat, 1))
        known = self.noise[0] is not None
        gps = _fit_shared_X(Xtrain, [ytrain[g, :].reshape(ndat, 1) for g in range(self.ng + 1)],
                            [self.noise[g] if known else None for g in range(self.ng + 1)], self.multi_hyper,
                            [True] + [False] * self.ng)
        GP_obj, GP_con = gps[0], gps[1:]
        GP_con_f = [{'type': 'ineq', 'fun': functools.partial(self.mistmatch_con, xk, GP_con[g], self.cons_model[g])}
                    for g in range(self.ng)]
        GP_con_f.append({'type': 'ineq', 'fun': self.TR_con})
        return GP_obj, GP_con_f, GP_con

    def find_min_so_far(self, funcs_system, X_opt):
        best = np.inf
        for i in range(len(X_opt)):
            y = funcs_system[0](np.array(X_opt[i]).flatten())
            if y < best:
                best = y
        return best

    # ---- acquisition multistart (utilities_2.py:964-992), all starts in lock step, one K4 launch pair per round
    def _optimise_acquisition(self, xk, d_inits, GP_obj, GP_con, TRb):
        h = GP_obj._h
        ng, ndim = self.ng, self.ndim
        group = _stack_gps([GP_obj] + list(GP_con))
        fb = _dev(np.array([self.obj_min if self.obj == 3 else 0.0]), h)
        obj_gp = {k: v[:1].contiguous() for k, v in group.items()}
        con_gp = {k: v[1:].contiguous() for k, v in group.items()} if ng else None
        cache = {}

        from . import systems as _systems
        model_batch = _systems.batch_form([self.obj_model] + list(self.cons_model))     # None unless all are known callables

        def evaluate(P):
            Xq = P + xk.reshape(1, ndim)
            if model_batch is not None:
                # same arithmetic per point as the scalar callables (bit-identical, see systems.py), all points at once
                M = model_batch(Xq)
                prior, cprior = np.ascontiguousarray(M[:, 0][None]), np.ascontiguousarray(M[:, 1:].T)
            else:
                prior = np.array([[float(self.obj_model(x)) for x in Xq]])
            sc = h.posterior_acq_host_np(obj_gp, Xq[None], prior, fb, self.obj, want=('score',))['score'][0]
            cons = np.empty((P.shape[0], ng + 1))
            if ng:
                if model_batch is None:
                    cprior = np.array([[float(np.asarray(self.cons_model[g](x)).reshape(-1)[0]) for x in Xq] for g in range(ng)])
                cm = h.posterior_acq_host_np(con_gp, np.broadcast_to(Xq, (ng,) + Xq.shape), cprior, None, api.ACQ_MEAN,
                                             want=('score',))['score']
                cons[:, :ng] = cm.T
            cons[:, ng] = self.Delta0 ** 2 - np.einsum('ij,ij->i', P, P)
            return sc, cons

        def fun_batch(inst, P):
            f, c = evaluate(P)
            cache['c'] = c
            return f

        def con_batch(inst, P):
            return cache.pop('c')

        return slsqp_batch.minimize_batch(fun_batch, d_inits, np.asarray(TRb, dtype=float), con_batch=con_batch, m_ineq=ng + 1,
                                          ftol=1e-12, maxiter=10000)

    # ---- main loop (utilities_2.py:888-1038)
    def RTO_routine(self):
        ng, ndim, n_iter, multi_opt = self.ng, self.ndim, self.n_iter, self.multi_opt
        sys_f = [self.obj_system] + self.cons_system
        mod_f = [self.obj_model] + self.cons_model
        self.backtrack_l = [False] * n_iter
        self.n_outputs = ng + 1

        def mismatch_row(x):
            row = np.zeros((1, ng + 1))
            for ii in range(ng + 1):
                row[0, ii] = np.asarray(sys_f[ii](np.array(x[:]).flatten()) - mod_f[ii](np.array(x[:]).flatten())).reshape(-1)[0]
            return row

        xnew = self.x0
        Xtrain = np.vstack((self.Xtrain, xnew))
        ytrain = np.hstack((self.ytrain, mismatch_row(xnew).T))
        GP_obj, GP_con_f, GP_con = self.GPs_construction(xnew, Xtrain, ytrain, 1)
        X_opt, y_opt = np.copy(Xtrain), np.copy(ytrain)
        self.obj_min = self.find_min_so_far(sys_f, X_opt)
        TR_l = ['error'] * (n_iter + 1)
        TR_l[0] = self.Delta0
        for i_rto in range(n_iter):
            print('============================ RTO iteration ', i_rto)
            xk = xnew
            TRb = self.bounds - xk.reshape(ndim, 1, order='F')
            sampler = self.Ellipse_sampling if self.scale_inputs else self.Ball_sampling
            d_inits = [sampler(ndim) for _ in range(multi_opt)]      # SLSQP draws no random numbers: same RNG order as :970-984
            res = self._optimise_acquisition(np.asarray(xk, dtype=float).reshape(-1), d_inits, GP_obj, GP_con, TRb)
            localval = np.array([r.fun if r.success else np.inf for r in res])
            if np.min(localval) == np.inf:
                print('warming, no feasible solution found')
                xnew = xk * (1 + 0.01 * np.random.randn())
            else:
                xnew = res[int(np.argmin(localval))].x + xk
            xnew = np.array([xnew]).reshape(1, ndim)
            ynew = mismatch_row(xnew)
            print('New Objective: ', -sys_f[0](np.array(xnew[:]).flatten()))     # consumes one plant-noise draw, as :1001
            X_opt = np.vstack((X_opt, xnew))
            y_opt = np.hstack((y_opt, ynew.T))
            self.obj_min = self.find_min_so_far(sys_f, X_opt)
            self.Delta0, xnew, xk = self.Step_constraint(self.Delta0, xk, xnew, GP_obj, i_rto)
            GP_obj, GP_con_f, GP_con = self.GPs_construction(xnew, X_opt, y_opt, 2 + i_rto)
            TR_l[i_rto + 1] = self.Delta0
        return X_opt, y_opt, TR_l, xnew, self.backtrack_l


# -------------------------------------------------------------------------------------------- group helpers
def _fit_shared_X(X, Ys, noises, multi_hyper, var_outs):
    """Fit len(Ys) single-output GPs on the same inputs with ONE lock-step multistart (objective + constraints of one
    RTO iteration, utilities_2.py:795-815), then build the GP_model objects around the result."""
    h = api.default_handle()
    X = np.asarray(X, dtype=float)
    d = X.shape[1]
    Xn = (X - np.mean(X, axis=0)) / np.std(X, axis=0)
    rows, bounds_rows = [], []
    for Y, nz in zip(Ys, noises):
        y_std = np.std(Y, axis=0)
        rows.append(((Y - np.mean(Y, axis=0)) / y_std)[:, 0])
        bounds_rows.append(_hyper_bounds(d, nz, y_std[0] if nz is not None else None))
    theta, info = _fit_group(h, _dev(Xn, h), _dev(np.stack(rows), h), bounds_rows, multi_hyper)
    return [GP_model(X, Y, 'RBF', multi_hyper=multi_hyper, var_out=vo, noise=nz, _theta=theta[g], _fit_info=[info[g]])
            for g, (Y, nz, vo) in enumerate(zip(Ys, noises, var_outs))]


def _stack_gps(gps):
    """Concatenate the device dicts of single-output GPs that share X into one batch for K4."""
    keys = ('Xn', 'invK', 'alpha', 'theta', 'x_mean', 'x_std', 'y_mean', 'y_std')
    return {k: torch.cat([g._gp[k] for g in gps], dim=0).contiguous() for k in keys}
