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


class _Inst:
    __slots__ = ('idx', 'x', 'state', 'fx', 'g', 'C', 'd', 'mult', 'buffer', 'indices', 'nfev', 'njev', 'done')


def minimize_batch(fun_batch, x0s, bounds, con_batch=None, m_ineq=0, ftol=1e-6, maxiter=100, eps=_EPS, inst_ids=None):
    """Run len(x0s) SLSQP instances in lock step.  bounds: (n, 2) array shared by all instances or a list of such
    arrays (one per instance).  Returns a list of Result(x, fun, nit, nfev, njev, status, message, success)."""
    x0s = [np.asarray(x, dtype=float).reshape(-1) for x in x0s]
    K = len(x0s)
    n = x0s[0].size
    if isinstance(bounds, np.ndarray) and bounds.ndim == 2:
        bounds = [bounds] * K
    m = int(m_ineq) if con_batch is not None else 0
    insts = []
    lbs, ubs, xls, xus = [], [], [], []
    for k in range(K):
        b = np.asarray(bounds[k], dtype=float)
        lb, ub = b[:, 0].copy(), b[:, 1].copy()
        if (lb > ub).any():
            raise ValueError('SLSQP Error: lb > ub in bounds')
        xl, xu = lb.copy(), ub.copy()
        xl[~np.isfinite(lb)] = np.nan
        xu[~np.isfinite(ub)] = np.nan
        lbs.append(lb); ubs.append(ub); xls.append(xl); xus.append(xu)
        it = _Inst()
        it.idx = k if inst_ids is None else inst_ids[k]
        it.x = np.clip(x0s[k], lb, ub)
        it.state = {'acc': ftol, 'alpha': 0.0, 'f0': 0.0, 'gs': 0.0, 'h1': 0.0, 'h2': 0.0, 'h3': 0.0, 'h4': 0.0, 't': 0.0,
                    't0': 0.0, 'tol': 10.0 * ftol, 'exact': 0, 'inconsistent': 0, 'reset': 0, 'iter': 0, 'itermax': int(maxiter),
                    'line': 0, 'm': m, 'meq': 0, 'mode': 0, 'n': n}
        it.indices = np.zeros([max(m + 2 * n + 2, 1)], dtype=np.int64 if _ILP64 else np.int32)
        size = n * (n + 1) // 2 + 3 * m * n + 9 * m + 8 * n * n + 35 * n + 28
        if m == 0:
            size += 2 * n * (n + 1)
        it.buffer = np.zeros(max(size, 1), dtype=np.float64)
        it.mult = np.zeros([max(1, m + 2 * n + 2)], dtype=np.float64)
        it.C = np.zeros([max(1, m), n], dtype=np.float64, order='F')
        it.d = np.zeros([max(1, m)], dtype=np.float64)
        it.g = np.zeros(n)
        it.fx = 0.0
        it.nfev = it.njev = 0
        it.done = False
        insts.append(it)

    LB, UB = np.asarray(lbs), np.asarray(ubs)
    eye = np.eye(n)

    def evaluate(need_f, need_g):
        """One batched round: f (and c) at the iterates of need_f, forward differences for need_g.  All index and
        step arithmetic is vectorised over the instances of the round (one NumPy call each, not one per instance)."""
        need_f, need_g = list(need_f), list(need_g)
        nf, ng_ = len(need_f), len(need_g)
        if nf + ng_ == 0:
            return
        blocks, ids = [], []
        if nf:
            Xf = np.clip(np.stack([insts[k].x for k in need_f]), LB[need_f], UB[need_f])
            blocks.append(Xf)
            ids.extend(insts[k].idx for k in need_f)
        if ng_:
            Xg = np.clip(np.stack([insts[k].x for k in need_g]), LB[need_g], UB[need_g])
            H = fd_steps(Xg, LB[need_g], UB[need_g], eps)                      # [ng, n]
            # point (k, i) = x_k + h_ki e_i ; the step actually taken is (x + h) - x  (approx_derivative does the same)
            Pg = Xg[:, None, :] + H[:, :, None] * eye[None, :, :]            # [ng, n, n]
            DX = (Xg + H) - Xg
            blocks.append(Pg.reshape(ng_ * n, n))
            for k in need_g:
                ids.extend([insts[k].idx] * n)
        ids = np.asarray(ids)
        P = np.concatenate(blocks, axis=0)
        f = np.asarray(fun_batch(ids, P), dtype=float).reshape(-1)
        c = np.asarray(con_batch(ids, P), dtype=float).reshape(P.shape[0], m) if m else None
        for j, k in enumerate(need_f):
            it = insts[k]
            it.fx = float(f[j]); it.nfev += 1
            if m:
                it.d[:m] = c[j]
        if ng_:
            Fg = f[nf:].reshape(ng_, n)
            Cg = c[nf:].reshape(ng_, n, m) if m else None
            for j, k in enumerate(need_g):
                it = insts[k]
                it.g[:] = (Fg[j] - it.fx) / DX[j]
                if m:
                    # c(x0) is the value stored at the last f/c evaluation of the same point (deterministic: not re-evaluated)
                    it.C[:m, :] = ((Cg[j] - it.d[:m][None, :]) / DX[j][:, None]).T
                it.nfev += n; it.njev += 1

    # mode 0 on entry: objective, constraints and both gradients at x0
    evaluate(range(K), range(K))
    active = list(range(K))
    while active:
        need_f, need_g, still = [], [], []
        for k in active:
            it = insts[k]
            _slsqp(it.state, it.fx, it.g, it.C, it.d, it.x, it.mult, xls[k], xus[k], it.buffer, it.indices)
            mode = it.state['mode']
            if mode == 1:
                need_f.append(k); still.append(k)
            elif mode == -1:
                need_g.append(k); still.append(k)
            else:
                it.done = True
        evaluate(need_f, need_g)
        active = still
    out = []
    for it in insts:
        mode = it.state['mode']
        out.append(Result(x=it.x, fun=it.fx, jac=it.g, nit=it.state['iter'], nfev=it.nfev, njev=it.njev, status=mode,
                          message=EXIT_MODES.get(mode, 'unknown'), success=(mode == 0), multipliers=it.mult[:m]))
    return out
