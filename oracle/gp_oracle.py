"""CPU oracle for the GP + acquisition hot path of GP-RTO-EI.  TEST INFRASTRUCTURE ONLY.

This is a NumPy/SciPy restatement of the arithmetic of the reference's hot path,
call for call (same library routines: scipy ``cdist(..., 'seuclidean')``, ``np.linalg.cholesky``,
general ``np.linalg.solve``, ``np.matmul``, ``scipy.stats.norm``), so that it reproduces the
reference's floating-point results and is a faithful stand-in for its CPU cost.  Every function
cites the reference lines (paths relative to /root/reference) it follows.

Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, as the CHECKER or the timed CPU baseline.  The product (gp_rto_ei_b200/) never imports it
and raises if its CUDA library is missing.

Parity pin: the reference ships NO tests, golden vectors or fixtures (SURVEY.md par. 4, 8c).  The
oracle is therefore pinned against outputs of the reference itself: tests/golden/make_golden.py
imports /root/reference/sub_uts/utilities_2.py UNMODIFIED (with inert stubs for the absent casadi /
matplotlib and the restated sobol_seq) in the build container and stores its inputs/outputs under
tests/golden/*.npz; tests/test_oracle_golden.py checks this file against those vectors bit-for-bit
where the call sequence is identical, and to 1e-12 otherwise.

Notation: theta = (0.5*log W_1..d, 0.5*log sf2, 0.5*log sn2), W = squared length-scales.
"""
import numpy as np
from scipy.spatial.distance import cdist
from scipy.stats import norm

JITTER = 1e-8          # utilities_2.py:105


# ----------------------------------------------------------------------------- normalisation
def normalise(X, Y):
    """utilities_2.py:38-40 - column mean / population std (ddof=0); no zero-std guard."""
    X_mean, X_std = np.mean(X, axis=0), np.std(X, axis=0)
    Y_mean, Y_std = np.mean(Y, axis=0), np.std(Y, axis=0)
    return X_mean, X_std, Y_mean, Y_std, (X - X_mean) / X_std, (Y - Y_mean) / Y_std


# ----------------------------------------------------------------------------- covariance
def cov_mat(kernel, X_norm, W, sf2):
    """utilities_2.py:50-68 - Gram matrix; dist = (sqrt(sum (u-v)^2/W))**2 as scipy computes it."""
    dist = cdist(X_norm, X_norm, 'seuclidean', V=W) ** 2
    r = np.sqrt(dist)
    if kernel == 'RBF':
        return sf2 * np.exp(-0.5 * dist)
    if kernel == 'Matern32':
        return sf2 * (1 + 3 ** 0.5 * r) * np.exp(-r * 3 ** 0.5)
    if kernel == 'Matern52':
        return sf2 * (1 + 5 ** 0.5 * r + 5 / 3 * r ** 2) * np.exp(-r * 5 ** 0.5)
    raise ValueError('no kernel with name %r' % (kernel,))   # the reference prints and returns None (:67-68)


def cov_sample(xnorm, X_norm, ell, sf2):
    """utilities_2.py:74-85 - cross-covariance column k(x*, X), always SE-ARD."""
    dist = cdist(X_norm, xnorm.reshape(1, X_norm.shape[1]), 'seuclidean', V=ell) ** 2
    return sf2 * np.exp(-.5 * dist)


def unpack_theta(theta, d):
    """utilities_2.py:100-102."""
    theta = np.asarray(theta, dtype=float)
    return np.exp(2 * theta[:d]), np.exp(2 * theta[d]), np.exp(2 * theta[d + 1])


# ----------------------------------------------------------------------------- NLML
def nlml(theta, X_norm, y_norm, kernel='RBF'):
    """utilities_2.py:92-114 - y'K^-1 y + log det K (no 1/2, no constant); raises LinAlgError like :107."""
    n, d = X_norm.shape
    W, sf2, sn2 = unpack_theta(theta, d)
    K = cov_mat(kernel, X_norm, W, sf2)
    K = K + (sn2 + JITTER) * np.eye(n)
    K = (K + K.T) * 0.5
    L = np.linalg.cholesky(K)
    logdetK = 2 * np.sum(np.log(np.diag(L)))
    invLY = np.linalg.solve(L, y_norm)
    alpha = np.linalg.solve(L.T, invLY)
    return np.dot(y_norm.T, alpha) + logdetK


def nlml_batch(thetas, X_norm, y_norm, kernel='RBF'):
    """NLML for a stack of hyper-parameter starts; (value, status) with LAPACK-style status
    (0 ok, k>0 = order of the first non-positive leading minor) instead of an exception."""
    thetas = np.atleast_2d(thetas)
    val = np.full(thetas.shape[0], np.nan)
    status = np.zeros(thetas.shape[0], dtype=np.int32)
    for s, th in enumerate(thetas):
        try:
            val[s] = nlml(th, X_norm, y_norm, kernel)
        except np.linalg.LinAlgError:
            status[s] = -1
    return val, status


def sobol_starts(sobol01, noise=None, y_std=None):
    """utilities_2.py:131-158 - map Sobol points in [0,1]^(d+2) onto the log-hyper box.
    Known noise variance pins the last bound to 0.5*log(noise/Y_std^2) +- 1e-3 (:149-154)."""
    dim = sobol01.shape[1]
    lb = np.array([-5.] * (dim - 1) + [-8.])
    ub = np.array([5.] * (dim - 1) + [-4.])
    if noise is not None:
        lb[-1] = np.log(noise / y_std ** 2) * 0.5 - 0.001
        ub[-1] = np.log(noise / y_std ** 2) * 0.5 + 0.001
    return lb + (ub - lb) * sobol01, lb, ub


# ----------------------------------------------------------------------------- final factor
def factor(theta_opt, X_norm, kernel='RBF'):
    """utilities_2.py:168-175 - Kopt = K + sn2*I (NO jitter, NO symmetrisation); invK = solve(Kopt, I)."""
    n, d = X_norm.shape
    ell, sf2, sn2 = unpack_theta(theta_opt, d)
    Kopt = cov_mat(kernel, X_norm, ell, sf2) + sn2 * np.eye(n)
    return np.linalg.solve(Kopt, np.eye(n))


# ----------------------------------------------------------------------------- inference
def inference(x, gp):
    """utilities_2.py:256-273 for one point.  gp: dict(X_norm, y_norm[n], theta[d+2], invK,
    X_mean, X_std, Y_mean, Y_std (scalars for the single output)).  Returns (mean, var) de-normalised."""
    d = gp['X_norm'].shape[1]
    xnorm = (np.asarray(x, dtype=float).reshape(-1) - gp['X_mean']) / gp['X_std']
    ell, sf2 = np.exp(2 * gp['theta'][:d]), np.exp(2 * gp['theta'][d])
    k = cov_sample(xnorm, gp['X_norm'], ell, sf2)
    mean = np.matmul(np.matmul(k.T, gp['invK']), gp['y_norm'])
    var = max(0., (sf2 - np.matmul(np.matmul(k.T, gp['invK']), k)).item())
    # the reference squares a 1-element ARRAY (stdY ** 2 -> x*x); a NumPy scalar ** 2 goes through pow() and can
    # differ in the last bit, so multiply explicitly
    return mean.item() * gp['Y_std'] + gp['Y_mean'], var * (gp['Y_std'] * gp['Y_std'])


def acquisition(mean_gp, var_gp, prior, f_best, acq):
    """utilities_2.py:485-504 - acq 1: prior+mean; 2: prior+mean-3*sqrt(var); 3: -EI with Z:=0 if sigma==0."""
    if acq == 1:
        return prior + mean_gp
    if acq == 2:
        return prior + mean_gp - 3 * np.sqrt(var_gp)
    if acq == 3:
        mean = prior + mean_gp
        Delta = f_best - mean
        sigma = np.sqrt(var_gp)
        Z = 0. if sigma == 0. else Delta / sigma
        return -(sigma * norm.pdf(Z) + Delta * norm.cdf(Z))
    raise ValueError('objective for GP not specified')


def posterior_acq_scalar(cands, gp, prior=None, f_best=0.0, acq=3):
    """The reference's call structure: one candidate at a time (GP_obj_f -> GP_inference_np).
    This is the faithful CPU baseline (bench.py --impl reference)."""
    m = cands.shape[0]
    mean, var, score = np.empty(m), np.empty(m), np.empty(m)
    for c in range(m):
        mean[c], var[c] = inference(cands[c], gp)
        score[c] = acquisition(mean[c], var[c], 0.0 if prior is None else prior[c], f_best, acq)
    return mean, var, score


def posterior_acq_vectorised(cands, gp, prior=None, f_best=0.0, acq=3):
    """Same algebra, all candidates of one GP at once (BLAS-3).  Not bit-identical to the scalar
    path (different BLAS kernels / summation order); used as the fast checker at large m and as the
    'best-effort NumPy' CPU figure reported beside the faithful one."""
    d = gp['X_norm'].shape[1]
    xn = (cands - gp['X_mean']) / gp['X_std']
    ell, sf2 = np.exp(2 * gp['theta'][:d]), np.exp(2 * gp['theta'][d])
    Ks = sf2 * np.exp(-.5 * cdist(xn, gp['X_norm'], 'seuclidean', V=ell) ** 2)     # (m, n)
    KsA = Ks @ gp['invK']
    mean_n = KsA @ gp['y_norm']
    var_n = np.maximum(0., sf2 - np.einsum('ij,ij->i', KsA, Ks))
    mean = mean_n * gp['Y_std'] + gp['Y_mean']
    var = var_n * (gp['Y_std'] * gp['Y_std'])
    p = 0.0 if prior is None else prior
    if acq == 1:
        return mean, var, p + mean
    if acq == 2:
        return mean, var, p + mean - 3 * np.sqrt(var)
    mu = p + mean
    Delta = f_best - mu
    sigma = np.sqrt(var)
    with np.errstate(divide='ignore', invalid='ignore'):
        Z = np.where(sigma == 0., 0., Delta / sigma)
    return mean, var, -(sigma * norm.pdf(Z) + Delta * norm.cdf(Z))


def make_gp(X, y, theta, kernel='RBF'):
    """Normalise + final factor for given hyper-parameters (utilities_2.py:38-40, 168-175)."""
    X = np.asarray(X, dtype=float)
    Y = np.asarray(y, dtype=float).reshape(-1, 1)
    X_mean, X_std, Y_mean, Y_std, X_norm, Y_norm = normalise(X, Y)
    theta = np.asarray(theta, dtype=float)
    return dict(X_norm=X_norm, y_norm=Y_norm[:, 0], theta=theta, invK=factor(theta, X_norm, kernel),
                X_mean=X_mean, X_std=X_std, Y_mean=Y_mean[0], Y_std=Y_std[0])


# ----------------------------------------------------------------------------- long-double truth
def _chol_ld(K):
    n = K.shape[0]
    L = np.array(K, dtype=np.longdouble)
    for j in range(n):
        L[j, j] = np.sqrt(L[j, j])
        L[j + 1:, j] /= L[j, j]
        L[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
    return np.tril(L)


def _trsv_lower_ld(L, B):
    n = L.shape[0]
    Z = np.array(B, dtype=np.longdouble)
    for i in range(n):
        Z[i] = (Z[i] - L[i, :i] @ Z[:i]) / L[i, i]
    return Z


def truth_ld(X_norm, y_norm, theta, xnorm_cands=None, jitter=0.0):
    """80-bit long-double arbiter (SURVEY.md par. 7.3-2): Cholesky-form NLML, mean, var for the SAME
    formulas.  Used only to decide, when GPU and NumPy disagree beyond cond(K)*eps, which is closer."""
    ld = np.longdouble
    n, d = X_norm.shape
    th = np.asarray(theta, dtype=ld)
    W, sf2, sn2 = np.exp(2 * th[:d]), np.exp(2 * th[d]), np.exp(2 * th[d + 1])
    Xl = np.asarray(X_norm, dtype=ld)
    D = ((Xl[:, None, :] - Xl[None, :, :]) ** 2 / W).sum(-1)
    K = sf2 * np.exp(-0.5 * D) + (sn2 + ld(jitter)) * np.eye(n, dtype=ld)
    L = _chol_ld(K)
    z = _trsv_lower_ld(L, np.asarray(y_norm, dtype=ld))
    out = dict(nll=float(z @ z + 2 * np.sum(np.log(np.diag(L)))))
    if xnorm_cands is not None:
        C = np.asarray(xnorm_cands, dtype=ld)
        Ds = ((Xl[:, None, :] - C[None, :, :]) ** 2 / W).sum(-1)      # (n, m)
        Ks = sf2 * np.exp(-0.5 * Ds)
        V = _trsv_lower_ld(L, Ks)
        out['mean_norm'] = np.asarray(V.T @ z, dtype=float)
        out['var_norm'] = np.asarray(np.maximum(0, sf2 - (V * V).sum(0)), dtype=float)
    return out
