"""Mint the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference ships no tests or known-answer vectors (SURVEY.md par. 4), so parity is pinned to the
outputs of its own code: sub_uts/utilities_2.py is imported as-is through oracle/ref_harness.py and
driven through GP_model / ITR_GP_RTO.GP_obj_f exactly as the mains drive it
(utilities_2.py:795-815, 477-507).  Outputs are stored next to their inputs as .npz.

Files written
  pickle_gps.npz    6 GPs fitted by the reference (multi_hyper=4) on (X, y) taken from the reference's
                    shipped result pickles (n = 24/25, d = 2): hypopt, invKopt, NLML at 15 Sobol starts,
                    mean/var and the three acquisition scores at 48 trust-region points.
  cfg3_small.npz    4 synthetic cfg-3 GPs (n=64, d=3, 256 candidates): Cov_mat, invK, mean/var, EI/LCB/mean.
  cfg4_small.npz    2 synthetic cfg-4 GPs (n=128, d=3) x 64 Sobol starts: negative_loglikelihood
                    (nan + flag where the reference raises LinAlgError).
  rto_simple.npz    three short seeded episodes of the reference's ITR_GP_RTO on the toy problem (trajectories, TR radii,
                    backtrack flags).
  rto_wo.npz        the same for the Williams-Otto configuration (2 constraint GPs, ellipse start sampling, model prior).
  fit_starts.npz    per-start results (x0, res.x, res.fun, nit, status) and the selected start index of the reference's
                    multistart on the six pickle data sets, free and known noise.
  ragged.npz        edge cases: n=1, n=2, n=7 (not a multiple of 8), d=1 and d=5, zero variance (candidate on a
                    noise-free data point), far-tail EI.
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_harness                      # noqa: E402
from gp_rto_ei_b200 import workloads                # noqa: E402
from gp_rto_ei_b200.compat import sobol_seq         # noqa: E402

ref = ref_harness.load()


def bare_gp(X, Y, theta, var_out=True):
    """A reference GP_model with given hyper-parameters: everything __init__ does (utilities_2.py:31-40)
    except the SLSQP search; the final factor is the reference's own lines 168-175 replayed through its
    own Cov_mat."""
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


def scorer(acq, obj_model, f_best):
    rto = object.__new__(ref.ITR_GP_RTO)
    rto.obj, rto.obj_model, rto.obj_min = acq, obj_model, f_best
    return rto


def prior_fn(x):
    x = np.asarray(x).reshape(-1)
    return float(0.3 * np.sum(x ** 2) - 0.1 * x[0])


def evaluate(gp, cand, f_best):
    """mean, var via GP_inference_np and the 3 scores via GP_obj_f with xk = 0, d = candidate."""
    m = cand.shape[0]
    mean, var = np.empty(m), np.empty(m)
    score = np.empty((3, m))
    prior = np.array([prior_fn(c) for c in cand])
    xk = np.zeros(cand.shape[1])
    for c in range(m):
        mu, v = gp.GP_inference_np(cand[c])
        mean[c], var[c] = mu[0], v[0]
        for a in (1, 2, 3):
            s = scorer(a, prior_fn, f_best).GP_obj_f(cand[c], gp, xk)
            score[a - 1, c] = np.asarray(s).reshape(-1)[0]
    return mean, var, score, prior


def nlml_ref(gp, theta):
    try:
        return float(gp.negative_loglikelihood(theta, gp.X_norm, gp.Y_norm[:, 0])), 0
    except np.linalg.LinAlgError:
        return np.nan, 1


def pickle_gps():
    rng = np.random.Generator(np.random.PCG64(7))
    srcs = [('run_simple/with_prior_with_exploration_EI.p', 0, 0), ('run_simple/with_prior_with_exploration_EI.p', 3, 1),
            ('run_simple/no_prior_with_exploration_ucb.p', 5, 0), ('run_WO/with_prior_with_exploration_ei_new2.p', 1, 0),
            ('run_WO/with_prior_with_exploration_ei_new2.p', 2, 2), ('run_WO/no_prior_with_exploration_ei.p', 4, 1)]
    out = {}
    for g, (path, ep, row) in enumerate(srcs):
        with open(os.path.join(ref_harness.REFERENCE_ROOT, path), 'rb') as fh:
            X_opt_mc, y_opt_mc = pickle.load(fh)[:2]
        X = np.ascontiguousarray(np.asarray(X_opt_mc[ep], dtype=float))
        Y = np.ascontiguousarray(np.asarray(y_opt_mc[ep], dtype=float)[row, :].reshape(-1, 1))
        np.random.seed(g)
        gp = ref.GP_model(X, Y, 'RBF', multi_hyper=4, var_out=True, noise=None)
        d = X.shape[1]
        starts = workloads.sobol_theta_starts(d, 15)
        nl = np.array([nlml_ref(gp, th) for th in starts])
        centre = X[-1]
        half = 0.25 * (X.max(0) - X.min(0))
        cand = centre + rng.uniform(-1, 1, (48, d)) * half
        cand[0] = X[3]                               # exactly on a data point
        f_best = float(Y.min())
        mean, var, score, prior = evaluate(gp, cand, f_best)
        K = gp.Cov_mat('RBF', gp.X_norm, np.exp(2 * gp.hypopt[:d, 0]), np.exp(2 * gp.hypopt[d, 0]))
        out.update({f'X{g}': X, f'Y{g}': Y, f'hypopt{g}': gp.hypopt.copy(), f'invK{g}': gp.invKopt[0],
                    f'K{g}': K, f'nlml_theta{g}': starts, f'nlml{g}': nl[:, 0], f'nlml_fail{g}': nl[:, 1],
                    f'nlml_opt{g}': nlml_ref(gp, gp.hypopt[:, 0])[0],
                    f'cand{g}': cand, f'prior{g}': prior, f'f_best{g}': f_best, f'mean{g}': mean, f'var{g}': var,
                    f'score{g}': score, f'cond{g}': np.linalg.cond(np.linalg.inv(gp.invKopt[0]))})
        print('pickle GP', g, path, 'n=%d' % X.shape[0], 'cond=%.2e' % out[f'cond{g}'], 'hyp', gp.hypopt.ravel())
    out['count'] = len(srcs)
    np.savez_compressed(os.path.join(HERE, 'pickle_gps.npz'), **out)


def fit_starts():
    """Per-start outcome of the reference's hyper-parameter multistart (utilities_2.py:156-167) on the six pickle data
    sets: res.x / res.fun of every SLSQP start and the index np.argmin selects.  Recorded by wrapping the `minimize`
    name the reference module calls (the module itself is not modified); known-noise variant included (bounds :149-154).
    The same multistart is repeated with the reference's NLML nudged by a few ulps (x (1 +- k 2^-52), k = 1..3) and by
    random relative perturbations of 1e-15: SLSQP with forward-difference gradients is chaotic with respect to the last
    bits of its objective, so individual starts - and with them the selected index - move between basins under such
    nudges.  The ensemble records where the reference's own selection is stable (only there is "identical start index"
    a defined requirement for an independent implementation)."""
    src = np.load(os.path.join(HERE, 'pickle_gps.npz'))
    out = {'count': int(src['count'])}
    orig = ref.minimize
    nll0 = ref.GP_model.negative_loglikelihood
    variants = [('x', 1.0 + k * 2.0 ** -52) for k in (1, 2, 3)] + [('x', 1.0 - k * 2.0 ** -53) for k in (1, 2, 3)] + [('r', sd) for sd in range(4)]

    def run(X, Y, noise, nll):
        log = []

        def rec(fun, x0, *a, **k):
            r = orig(fun, x0, *a, **k)
            log.append((np.array(x0, dtype=float), np.array(r.x, dtype=float), float(np.asarray(r.fun).reshape(-1)[0]), int(r.nit), int(r.status)))
            return r
        ref.minimize, ref.GP_model.negative_loglikelihood = rec, nll
        try:
            gp = ref.GP_model(X, Y, 'RBF', multi_hyper=4, var_out=True, noise=noise)
        finally:
            ref.minimize, ref.GP_model.negative_loglikelihood = orig, nll0
        return gp, log

    for g in range(int(src['count'])):
        X, Y = src[f'X{g}'], src[f'Y{g}']
        for tag, noise in (('', None), ('_kn', 1e-3)):
            gp, log = run(X, Y, noise, nll0)
            vals = np.array([l[2] for l in log])
            out.update({f'x0{tag}{g}': np.stack([l[0] for l in log]), f'localsol{tag}{g}': np.stack([l[1] for l in log]),
                        f'localval{tag}{g}': vals, f'nit{tag}{g}': np.array([l[3] for l in log]),
                        f'status{tag}{g}': np.array([l[4] for l in log]), f'minindex{tag}{g}': int(np.argmin(vals)),
                        f'hypopt{tag}{g}': gp.hypopt.copy()})
            assert np.array_equal(gp.hypopt[:, 0], log[int(np.argmin(vals))][1])
            ens_v, ens_i = [], []
            for kind, par in variants:
                if kind == 'x':
                    def nll_p(self, hyper, X_, Y_, f=par):
                        return nll0(self, hyper, X_, Y_) * f
                else:
                    rng = np.random.default_rng(par)

                    def nll_p(self, hyper, X_, Y_, rng=rng):
                        return nll0(self, hyper, X_, Y_) * (1 + 1e-15 * rng.standard_normal())
                _, lg = run(X, Y, noise, nll_p)
                v = np.array([l[2] for l in lg])
                ens_v.append(v); ens_i.append(int(np.argmin(v)))
            out[f'ens_localval{tag}{g}'] = np.stack(ens_v); out[f'ens_minindex{tag}{g}'] = np.array(ens_i)
            print('fit_starts GP', g, tag or 'free', 'localval', vals, 'minindex', int(np.argmin(vals)), '| ensemble indices', ens_i,
                  'distinct optima', np.unique(np.round(np.min(np.stack(ens_v), axis=1), 6)))
    np.savez_compressed(os.path.join(HERE, 'fit_starts.npz'), **out)


def cfg3_small():
    w = workloads.cfg3_numpy(seed=0, B=4, n=64, d=3, m=256)
    B, n, d = w['X'].shape
    m = w['cand'].shape[1]
    out = dict(w)
    out['K'] = np.empty((B, n, n)); out['invK'] = np.empty((B, n, n))
    out['mean'] = np.empty((B, m)); out['var'] = np.empty((B, m)); out['score'] = np.empty((B, 3, m))
    out['prior'] = np.empty((B, m)); out['cond'] = np.empty(B); out['nlml_at_theta'] = np.empty(B)
    for b in range(B):
        gp = bare_gp(w['X'][b], w['y'][b].reshape(-1, 1), w['theta'][b])
        out['K'][b] = gp.Cov_mat('RBF', gp.X_norm, np.exp(2 * gp.hypopt[:d, 0]), np.exp(2 * gp.hypopt[d, 0]))
        out['invK'][b] = gp.invKopt[0]
        out['mean'][b], out['var'][b], out['score'][b], out['prior'][b] = evaluate(gp, w['cand'][b], w['f_best'][b])
        out['cond'][b] = np.linalg.cond(np.linalg.inv(gp.invKopt[0]))
        out['nlml_at_theta'][b] = nlml_ref(gp, w['theta'][b])[0]
        print('cfg3 GP', b, 'cond=%.2e' % out['cond'][b])
    np.savez_compressed(os.path.join(HERE, 'cfg3_small.npz'), **out)


def cfg4_small():
    w = workloads.cfg4_numpy(seed=1, B=2, n=128, d=3, S=64)
    B, S = w['theta'].shape[:2]
    out = dict(w)
    out['nlml'] = np.empty((B, S)); out['fail'] = np.zeros((B, S), dtype=np.int32)
    for b in range(B):
        gp = bare_gp(w['X'][b], w['y'][b].reshape(-1, 1), w['theta'][b, 0])
        for s in range(S):
            out['nlml'][b, s], out['fail'][b, s] = nlml_ref(gp, w['theta'][b, s])
    print('cfg4: failures', int(out['fail'].sum()), 'of', B * S)
    np.savez_compressed(os.path.join(HERE, 'cfg4_small.npz'), **out)


def ragged():
    rng = np.random.Generator(np.random.PCG64(11))
    cases = [(2, 1), (3, 2), (7, 2), (9, 1), (13, 5), (33, 3), (100, 4)]
    out = {'count': len(cases)}
    for g, (n, d) in enumerate(cases):
        X = rng.uniform(-2, 3, (n, d))
        Y = (np.cos(X.sum(1)) + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
        theta = np.concatenate([rng.uniform(-0.3, 0.8, d), rng.uniform(-0.3, 0.3, 1), [0.5 * np.log(10.0 ** rng.uniform(-4, -2))]])
        gp = bare_gp(X, Y, theta)
        cand = X[rng.integers(0, n, 24)] + rng.uniform(-0.6, 0.6, (24, d))
        cand[0] = X[0]
        cand[1] = X[0] + 40.0                       # far away: k -> 0, var -> sf2, EI far tail
        f_best = float(Y.min()) - (5.0 if g % 2 else 0.0)    # odd cases: deep in the EI tail
        mean, var, score, prior = evaluate(gp, cand, f_best)
        starts = workloads.sobol_theta_starts(d, 8)
        nl = np.array([nlml_ref(gp, th) for th in starts])
        out.update({f'X{g}': X, f'Y{g}': Y, f'theta{g}': theta, f'invK{g}': gp.invKopt[0], f'cand{g}': cand,
                    f'prior{g}': prior, f'f_best{g}': f_best, f'mean{g}': mean, f'var{g}': var, f'score{g}': score,
                    f'nlml_theta{g}': starts, f'nlml{g}': nl[:, 0], f'nlml_fail{g}': nl[:, 1]})
    # sigma == 0 quirk (utilities_2.py:499-500): noise-free GP queried on a data point clamps var to 0
    X = np.array([[0.0], [1.0], [2.5]]); Y = np.array([[1.0], [-0.5], [0.7]])
    theta = np.array([0.0, 0.0, -40.0])
    gp = bare_gp(X, Y, theta)
    cand = np.array([[0.0], [1.0], [2.5], [0.5]])
    mean, var, score, prior = evaluate(gp, cand, 0.1)
    out.update(dict(Xz=X, Yz=Y, thetaz=theta, invKz=gp.invKopt[0], candz=cand, priorz=prior, f_bestz=0.1,
                    meanz=mean, varz=var, scorez=score))
    print('sigma==0 case: var', var, 'EI', score[2])
    np.savez_compressed(os.path.join(HERE, 'ragged.npz'), **out)


def rto_simple():
    """Short seeded episodes of the reference's own ITR_GP_RTO on the toy problem (arguments of
    run_simple/main_simple.py:179-206 with fewer iterations/starts): noiseless plant (deterministic given the seeds)
    and the noisy plant of the script (checks the RNG consumption order)."""
    import contextlib
    import io
    import random
    rsys = ref_harness.load('systems')
    out = {}
    for tag, plant, cons, acq, noise in (('noiseless_ei', rsys.Benoit_System_noiseless, rsys.con1_system_tight_noiseless, 3, None),
                                         ('noisy_ei', rsys.Benoit_System, rsys.con1_system_tight, 3, [1e-3] * 2),
                                         ('noiseless_lcb', rsys.Benoit_System_noiseless, rsys.con1_system_tight_noiseless, 2, None)):
        def episode():
            np.random.seed(0); random.seed(0)
            Xtrain = np.array([[1.2, 0.], [1.4, 0.1], [1.3, -0.1]])
            with contextlib.redirect_stdout(io.StringIO()):
                rto = ref.ITR_GP_RTO(rsys.Benoit_Model, plant, [rsys.con1_model], [cons], np.array([1.1, -0.1]), 0.25, 0.7, 0.2, 0.8,
                                     0.8, 1.2, 4, ['data0', Xtrain], np.array([[-.6, 1.5], [-1., 1.]]), obj_setting=acq, noise=noise,
                                     multi_opt=6, multi_hyper=4, TR_scaling=False, TR_curvature=False, store_data=True, inner_TR=False,
                                     scale_inputs=False)
                return rto.RTO_routine()
        X_opt, y_opt, TR_l, xnew, backtrack = episode()
        # Chaotic case (LCB, free noise, 4 data points): an ensemble of the reference under RELATIVE perturbations of its
        # NLML of 1e-15 ... 1e-12 - the trust-region sequence itself flips between two branches.
        if tag == 'noiseless_lcb':
            nll_nom = ref.GP_model.negative_loglikelihood
            ens_X, ens_TR = [], []
            for e, (rel, sd) in enumerate([(r, 99) for r in (1e-15, 1e-14, 1e-13, 1e-12)] + [(1e-13, q) for q in range(100, 108)]):
                rng = np.random.default_rng(sd)

                def nll_p(self, hyper, X, Y, rel=rel, rng=rng):
                    return nll_nom(self, hyper, X, Y) * (1 + rel * rng.standard_normal())
                ref.GP_model.negative_loglikelihood = nll_p
                try:
                    o = episode()
                finally:
                    ref.GP_model.negative_loglikelihood = nll_nom
                ens_X.append(o[0]); ens_TR.append(np.array(o[2], dtype=float))
                print('   ensemble rel=%g seed=%d TR %s max|dX| %.2e' % (rel, sd, o[2], np.abs(o[0] - X_opt).max()))
            out[tag + '_ens_X'] = np.stack(ens_X); out[tag + '_ens_TR'] = np.stack(ens_TR)
        # Self-sensitivity of the reference: the same episode with its NLML and posterior nudged by ONE ulp
        # (x (1 + 2^-52)).  Any other correct implementation differs from NumPy by at least that much, so the spread of
        # the iterates under this nudge is the floor for trajectory parity (SURVEY.md par. 7.3-1).
        nll0, inf0 = ref.GP_model.negative_loglikelihood, ref.GP_model.GP_inference_np
        ulp = 1.0 + 2.0 ** -52

        def nll1(self, hyper, X, Y):
            return nll0(self, hyper, X, Y) * ulp

        def inf1(self, x):
            r = inf0(self, x)
            return (r[0] * ulp, r[1] * ulp) if isinstance(r, tuple) else r * ulp
        ref.GP_model.negative_loglikelihood, ref.GP_model.GP_inference_np = nll1, inf1
        try:
            Xp, yp, TRp, _, btp = episode()
        finally:
            ref.GP_model.negative_loglikelihood, ref.GP_model.GP_inference_np = nll0, inf0
        out.update({tag + '_X_opt': X_opt, tag + '_y_opt': y_opt, tag + '_TR_l': np.array(TR_l, dtype=float),
                    tag + '_xnew': np.asarray(xnew), tag + '_backtrack': np.array(backtrack, dtype=bool),
                    tag + '_X_opt_ulp': Xp, tag + '_TR_l_ulp': np.array(TRp, dtype=float), tag + '_backtrack_ulp': np.array(btp, dtype=bool)})
        print('rto', tag, 'TR', TR_l, 'backtrack', backtrack, 'x', np.asarray(xnew).ravel(),
              '| one-ulp self-sensitivity max|dX| = %.3e, same TR/backtrack: %s' %
              (np.abs(Xp - X_opt).max(), np.array_equal(TRp, TR_l) and list(btp) == list(backtrack)))
    np.savez_compressed(os.path.join(HERE, 'rto_simple.npz'), **out)


def rto_wo():
    """Short seeded episodes of the reference's ITR_GP_RTO in the Williams-Otto configuration of
    run_WO/main_WO_prior_EI.py (ng = 2 constraint GPs, scale_inputs=True -> Ellipse_sampling, model prior, known noise).
    casadi is not available, so the plant / model callables are this repo's NumPy Newton restatement
    (gp_rto_ei_b200/sub_uts/systems.py) - the class under test is the reference's own driver and GP."""
    import contextlib
    import io
    import random
    from gp_rto_ei_b200.sub_uts import systems as S
    out = {}
    for tag, acq, noisy in (('wo_ei_noiseless', 3, False), ('wo_ei_noisy', 3, True), ('wo_mean_noisy', 1, True)):
        def episode():
            np.random.seed(0); random.seed(0)
            model, plant = S.WO_model(), S.WO_system()
            obj_sys = plant.WO_obj_sys_ca if noisy else plant.WO_obj_sys_ca_noise_less
            cons_sys = [plant.WO_con1_sys_ca, plant.WO_con2_sys_ca] if noisy else [plant.WO_con1_sys_ca_noise_less, plant.WO_con2_sys_ca_noise_less]
            Xtrain = np.array([[5.7, 74.], [6.35, 74.9], [6.6, 75.], [6.75, 79.]])
            with contextlib.redirect_stdout(io.StringIO()):
                rto = ref.ITR_GP_RTO(model.WO_obj_ca, obj_sys, [model.WO_con1_model_ca, model.WO_con2_model_ca], cons_sys,
                                     np.array([6.9, 83.]), 0.25, 0.7, 0.2, 0.8, 0.8, 1.2, 3, ['data0', Xtrain], np.array([[4., 7.], [70., 100.]]),
                                     obj_setting=acq, noise=[0.5 ** 2, 5e-8, 5e-8], multi_opt=5, multi_hyper=3, TR_scaling=False,
                                     TR_curvature=False, store_data=True, inner_TR=False, scale_inputs=True)
                return rto.RTO_routine()
        X_opt, y_opt, TR_l, xnew, backtrack = episode()
        nll0, inf0 = ref.GP_model.negative_loglikelihood, ref.GP_model.GP_inference_np
        ulp = 1.0 + 2.0 ** -52

        def nll1(self, hyper, X, Y):
            return nll0(self, hyper, X, Y) * ulp

        def inf1(self, x):
            r = inf0(self, x)
            return (r[0] * ulp, r[1] * ulp) if isinstance(r, tuple) else r * ulp
        ref.GP_model.negative_loglikelihood, ref.GP_model.GP_inference_np = nll1, inf1
        try:
            Xp, _, TRp, _, btp = episode()
        finally:
            ref.GP_model.negative_loglikelihood, ref.GP_model.GP_inference_np = nll0, inf0
        out.update({tag + '_X_opt': X_opt, tag + '_y_opt': y_opt, tag + '_TR_l': np.array(TR_l, dtype=float),
                    tag + '_backtrack': np.array(backtrack, dtype=bool), tag + '_X_opt_ulp': Xp,
                    tag + '_same_signature_ulp': np.array(np.array_equal(TRp, TR_l) and list(btp) == list(backtrack))})
        print('rto', tag, 'TR', TR_l, 'backtrack', backtrack, 'x', np.asarray(xnew).ravel(),
              '| one-ulp self-sensitivity max|dX| = %.3e, same TR/backtrack: %s' % (np.abs(Xp - X_opt).max(), out[tag + '_same_signature_ulp']))
    np.savez_compressed(os.path.join(HERE, 'rto_wo.npz'), **out)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        globals()[sys.argv[1]]()
        sys.exit(0)
    pickle_gps()
    fit_starts()
    cfg3_small()
    cfg4_small()
    ragged()
    rto_simple()
    rto_wo()
    print('golden vectors written to', HERE)
