"""The lock-step batched SLSQP driver must reproduce scipy.optimize.minimize(method='SLSQP') bit-for-bit when it is
given the same (CPU) function: same iterates, same exit codes.  This pins the host logic of SURVEY.md par. 8f-1 on the
two reference call sites (utilities_2.py:160-161 hyper-parameter fit, :978-979 constrained acquisition)."""
import numpy as np
from scipy.optimize import minimize

from conftest import load_golden
from gp_rto_ei_b200 import slsqp_batch, workloads
from oracle import gp_oracle as O


def test_hyperparameter_fit_matches_scipy_bitwise():
    g = load_golden('pickle_gps.npz')
    for i in (0, 3):
        X, Y = g[f'X{i}'], g[f'Y{i}']
        _, _, _, _, Xn, Yn = O.normalise(X, Y)
        d = X.shape[1]
        starts = workloads.sobol_theta_starts(d, 5)
        lb = np.array([-5.] * (d + 1) + [-8.]); ub = np.array([5.] * (d + 1) + [-4.])
        bounds = np.stack([lb, ub], axis=1)
        ref = [minimize(O.nlml, s, args=(Xn, Yn[:, 0]), method='SLSQP', options={'disp': False, 'maxiter': 10000},
                        bounds=bounds, tol=1e-12) for s in starts]

        def fun_batch(inst, P):
            return np.array([O.nlml(p, Xn, Yn[:, 0]) for p in P])

        got = slsqp_batch.minimize_batch(fun_batch, starts, bounds, ftol=1e-12, maxiter=10000)
        for r, q in zip(ref, got):
            assert np.array_equal(r.x, q.x) and r.fun == q.fun
            assert r.status == q.status and r.nit == q.nit and r.nfev == q.nfev and r.njev == q.njev


def test_constrained_problem_matches_scipy_bitwise():
    rng = np.random.default_rng(0)

    def f(x):
        return (x[0] - 0.7) ** 2 + 3 * (x[1] + 0.2) ** 2 + np.sin(3 * x[0]) * x[1]

    cons = [{'type': 'ineq', 'fun': lambda x: 0.25 - x @ x},
            {'type': 'ineq', 'fun': lambda x: x[0] + 2 * x[1] + 0.3}]
    bounds = np.array([[-0.4, 0.45], [-0.5, 0.5]])
    x0s = rng.uniform(-0.3, 0.3, (12, 2))
    x0s[0] = [0.45, 0.5]                                   # on the upper bounds: forward steps flip
    ref = [minimize(f, x0, method='SLSQP', options={'disp': False, 'maxiter': 10000}, bounds=bounds, constraints=cons, tol=1e-12)
           for x0 in x0s]

    def fun_batch(inst, P):
        return np.array([f(p) for p in P])

    def con_batch(inst, P):
        return np.array([[c['fun'](p) for c in cons] for p in P])

    got = slsqp_batch.minimize_batch(fun_batch, x0s, bounds, con_batch=con_batch, m_ineq=2, ftol=1e-12, maxiter=10000)
    for r, q in zip(ref, got):
        assert np.array_equal(r.x, q.x) and r.fun == q.fun and r.status == q.status and r.nit == q.nit
        assert r.success == q.success


def test_fd_steps_follow_scipy():
    from scipy.optimize._numdiff import approx_derivative
    lb, ub = np.array([-1.0, -1.0, 0.0]), np.array([1.0, 1.0, 1e-9])
    x0 = np.array([1.0, 0.3, 0.5e-9])
    h = slsqp_batch.fd_steps(x0, lb, ub)
    seen = []
    approx_derivative(lambda x: (seen.append(x.copy()), x.sum())[1], x0, method='2-point', abs_step=slsqp_batch._EPS, bounds=(lb, ub))
    pts = [p for p in seen if not np.array_equal(p, x0)]
    for i, p in enumerate(pts):
        assert p[i] == x0[i] + h[i]
