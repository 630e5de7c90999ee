"""Lock-step batched SLSQP: many independent SLSQP instances advanced together so that every round of
function / finite-difference evaluations becomes ONE batched GPU call (SURVEY.md par. 8f-1).

The reference calls ``scipy.optimize.minimize(method='SLSQP', tol=1e-12)`` once per multistart, one objective
evaluation at a time (reference: sub_uts/utilities_2.py:160-161 for the hyper-parameter fit, :978-979 for the
acquisition).  SciPy's SLSQP is a reverse-communication C routine (``scipy.optimize._slsqplib.slsqp``) driven by a
Python loop (``scipy/optimize/_slsqp_py.py``); all of its state lives in caller-owned arrays.  This module drives
K instances of that same routine in lock step and re-implements the surrounding bookkeeping of
``_minimize_slsqp``: clipping of x0 to the bounds, 2-point forward differences with the absolute step
``sqrt(eps)`` flipped at the upper bound (``approx_derivative(..., abs_step=eps, bounds=...)``), the evaluation
order (f and c on mode 1, gradients on mode -1) and the exit codes.  Given bit-identical function values it
reproduces SciPy's iterates bit-for-bit (tests/test_slsqp_batch.py checks this against scipy.optimize.minimize).

Callbacks
    fun_batch(inst, X) -> f[k]            objective of instance inst[i] at X[i]          (k points per call)
    con_batch(inst, X) -> c[k, m]         inequality constraints c >= 0 (optional)
Both receive every point of a round at once: the current iterates that need f/c plus all forward-difference
points of the instances that need gradients.
"""
import numpy as np

try:
    from scipy.optimize._slsqplib import slsqp as _slsqp
    from scipy.linalg.lapack import HAS_ILP64 as _ILP64
except ImportError as exc:                                    # pragma: no cover
    raise ImportError('slsqp_batch drives scipy.optimize._slsqplib.slsqp (SciPy >= 1.16); ' + str(exc))

_EPS = float(np.sqrt(np.finfo(float).eps))                     # scipy.optimize._optimize._epsilon

EXIT_MODES = {-1: 'Gradient evaluation required (g & a)', 0: 'Optimization terminated successfully',
              1: 'Function evaluation required (f & c)', 2: 'More equality constraints than independent variables',
              3: 'More than 3*n iterations in LSQ subproblem', 4: 'Inequality constraints incompatible',
              5: 'Singular matrix E in LSQ subproblem', 6: 'Singular matrix C in LSQ subproblem',
              7: 'Rank-deficient equality constraint subproblem HFTI', 8: 'Positive directional derivative for linesearch',
              9: 'Iteration limit reached'}


class Result(dict):
    __getattr__ = dict.__getitem__


def fd_steps(x0, lb, ub, eps=_EPS):
    """Forward-difference steps of approx_derivative(method='2-point', abs_step=eps, bounds=(lb, ub)); x0, lb, ub may
    carry a leading batch axis."""
    h = np.full_like(x0, eps)
    dx = (x0 + h) - x0
    sign = (x0 >= 0).astype(float) * 2 - 1
    h = np.where(dx == 0, _EPS * sign * np.maximum(1.0, np.abs(x0)), h)
    if np.all((lb == -np.inf) & (ub == np.inf)):
        return h
    lower, upper = x0 - lb, ub - x0
    x = x0 + h
    violated = (x < lb) | (x > ub)
    fitting = np.abs(h) <= np.maximum(lower, upper)
    h = np.where(violated & fitting, -h, h)
    h = np.where((upper >= lower) & ~fitting, upper, h)
    h = np.where((upper < lower) & ~fitting, -lower, h)
    return h


def _check_scipy():
    """This module drives a PRIVATE SciPy routine with the state-dict calling convention introduced in SciPy 1.16 and sizes
    its workspace with SciPy's own formula; tests/test_slsqp_batch.py (mandatory CPU test) pins bit-identity with
    scipy.optimize.minimize for the installed version.  Refuse versions older than the C port."""
    import scipy
    major, minor = (int(v) for v in scipy.__version__.split('.')[:2])
    if (major, minor) < (1, 16):
        raise ImportError('slsqp_batch needs SciPy >= 1.16 (state-dict _slsqplib.slsqp); found ' + scipy.__version__)


_check_scipy()


def minimize_arrays(fun_batch, X0, LB, UB, con_batch=None, m_ineq=0, ftol=1e-6, maxiter=100, eps=_EPS, inst_ids=None):
    """Core of the lock-step driver, struct-of-arrays: K instances of dimension n (X0, LB, UB: [K, n]) advanced together.
    All per-round bookkeeping (clipping, forward-difference points, gradients, constraint normals) is one NumPy
    operation over the instances of the round; the only per-instance Python work is the call of SciPy's C routine.
    Returns dict(x[K,n], fun[K], status[K], nit[K], nfev[K], njev[K], jac[K,n], mult[K,m])."""
    X0 = np.asarray(X0, dtype=float)
    K, n = X0.shape
    LB = np.ascontiguousarray(np.broadcast_to(np.asarray(LB, dtype=float), (K, n)))
    UB = np.ascontiguousarray(np.broadcast_to(np.asarray(UB, dtype=float), (K, n)))
    if (LB > UB).any():
        raise ValueError('SLSQP Error: lb > ub in bounds')
    m = int(m_ineq) if con_batch is not None else 0
    m1 = max(1, m)
    ids = np.arange(K) if inst_ids is None else np.asarray(inst_ids)
    XL, XU = LB.copy(), UB.copy()
    XL[~np.isfinite(LB)] = np.nan
    XU[~np.isfinite(UB)] = np.nan
    X = np.clip(X0, LB, UB)
    FX = np.zeros(K)
    G = np.zeros((K, n))
    CT = np.zeros((K, n, m1))                 # CT[k].T is the (m, n) Fortran-ordered constraint-normal matrix of instance k
    D = np.zeros((K, m1))
    MULT = np.zeros((K, max(1, m + 2 * n + 2)))
    size = n * (n + 1) // 2 + 3 * m * n + 9 * m + 8 * n * n + 35 * n + 28
    if m == 0:
        size += 2 * n * (n + 1)
    BUF = np.zeros((K, max(size, 1)))
    IND = np.zeros((K, max(m + 2 * n + 2, 1)), dtype=np.int64 if _ILP64 else np.int32)
    NFEV = np.zeros(K, dtype=np.int64)
    NJEV = np.zeros(K, dtype=np.int64)
    states, args = [], []
    for k in range(K):
        st = {'acc': ftol, 'alpha': 0.0, 'f0': 0.0, 'gs': 0.0, 'h1': 0.0, 'h2': 0.0, 'h3': 0.0, 'h4': 0.0, 't': 0.0,
              't0': 0.0, 'tol': 10.0 * ftol, 'exact': 0, 'inconsistent': 0, 'reset': 0, 'iter': 0, 'itermax': int(maxiter),
              'line': 0, 'm': m, 'meq': 0, 'mode': 0, 'n': n}
        states.append(st)
        args.append((G[k], CT[k].T, D[k], X[k], MULT[k], XL[k], XU[k], BUF[k], IND[k]))
    eye = np.eye(n)

    def evaluate(need_f, need_g):
        """One batched round: f (and c) at the iterates of need_f, forward differences for need_g (index arrays)."""
        nf, ng_ = need_f.size, need_g.size
        if nf + ng_ == 0:
            return
        blocks, pid = [], []
        if nf:
            blocks.append(np.clip(X[need_f], LB[need_f], UB[need_f]))
            pid.append(ids[need_f])
        if ng_:
            Xg = np.clip(X[need_g], LB[need_g], UB[need_g])
            H = fd_steps(Xg, LB[need_g], UB[need_g], eps)                      # [ng, n]
            # point (k, i) = x_k + h_ki e_i ; the step actually taken is (x + h) - x  (approx_derivative does the same)
            Pg = Xg[:, None, :] + H[:, :, None] * eye[None, :, :]            # [ng, n, n]
            DX = (Xg + H) - Xg
            blocks.append(Pg.reshape(ng_ * n, n))
            pid.append(np.repeat(ids[need_g], n))
        P = np.concatenate(blocks, axis=0)
        pid = np.concatenate(pid)
        f = np.asarray(fun_batch(pid, P), dtype=float).reshape(-1)
        c = np.asarray(con_batch(pid, P), dtype=float).reshape(P.shape[0], m) if m else None
        if nf:
            FX[need_f] = f[:nf]
            NFEV[need_f] += 1
            if m:
                D[need_f, :m] = c[:nf]
        if ng_:
            G[need_g] = (f[nf:].reshape(ng_, n) - FX[need_g][:, None]) / DX
            if m:
                # c(x0) is the value stored at the last f/c evaluation of the same point (deterministic: not re-evaluated)
                CT[need_g, :, :m] = (c[nf:].reshape(ng_, n, m) - D[need_g, None, :m]) / DX[:, :, None]
            NFEV[need_g] += n
            NJEV[need_g] += 1

    # mode 0 on entry: objective, constraints and both gradients at x0
    allk = np.arange(K)
    evaluate(allk, allk)
    active = allk
    calls = 0
    while active.size:
        calls += int(active.size)
        fxl = FX[active].tolist()
        modes = np.empty(active.size, dtype=np.int64)
        for j, k in enumerate(active.tolist()):
            st = states[k]
            _slsqp(st, fxl[j], *args[k])
            modes[j] = st['mode']
        evaluate(active[modes == 1], active[modes == -1])
        active = active[np.abs(modes) == 1]
    status = np.array([st['mode'] for st in states], dtype=np.int64)
    nit = np.array([st['iter'] for st in states], dtype=np.int64)
    return dict(x=X, fun=FX, status=status, nit=nit, nfev=NFEV, njev=NJEV, jac=G, mult=MULT[:, :m], slsqp_calls=calls)


def minimize_batch(fun_batch, x0s, bounds, con_batch=None, m_ineq=0, ftol=1e-6, maxiter=100, eps=_EPS, inst_ids=None):
    """Run len(x0s) SLSQP instances in lock step.  bounds: (n, 2) array shared by all instances or a list of such
    arrays (one per instance).  Returns a list of Result(x, fun, nit, nfev, njev, status, message, success)."""
    X0 = np.stack([np.asarray(x, dtype=float).reshape(-1) for x in x0s])
    K = X0.shape[0]
    if isinstance(bounds, np.ndarray) and bounds.ndim == 2:
        B = np.broadcast_to(bounds[None], (K,) + bounds.shape)
    else:
        B = np.stack([np.asarray(b, dtype=float) for b in bounds])
    r = minimize_arrays(fun_batch, X0, B[:, :, 0], B[:, :, 1], con_batch=con_batch, m_ineq=m_ineq, ftol=ftol, maxiter=maxiter, eps=eps,
                        inst_ids=inst_ids)
    out = []
    for k in range(K):
        mode = int(r['status'][k])
        out.append(Result(x=r['x'][k], fun=float(r['fun'][k]), jac=r['jac'][k], nit=int(r['nit'][k]), nfev=int(r['nfev'][k]),
                          njev=int(r['njev'][k]), status=mode, message=EXIT_MODES.get(mode, 'unknown'), success=(mode == 0),
                          multipliers=r['mult'][k]))
    return out
