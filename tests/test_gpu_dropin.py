"""The reference-facing drop-in (gp_rto_ei_b200/sub_uts/utilities_2.py) against golden runs of the reference's own
classes: GP_model fit + inference, and short seeded RTO episodes of ITR_GP_RTO."""
import contextlib
import io
import random

import numpy as np
import pytest

from conftest import load_golden
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def u2(handle):
    from gp_rto_ei_b200.sub_uts import utilities_2
    return utilities_2


def test_gp_model_fit_reaches_reference_optimum(u2):
    """Same multistart (Sobol starts, SLSQP tol 1e-12, FD gradients) on the GPU NLML: the selected optimum has the
    reference's NLML value to 1e-9 relative, and the public attributes have the reference's shapes and values."""
    g = load_golden('pickle_gps.npz')
    for i in range(int(g['count'])):
        X, Y = g[f'X{i}'], g[f'Y{i}']
        gp = u2.GP_model(X, Y, 'RBF', multi_hyper=4, var_out=True, noise=None)
        assert gp.hypopt.shape == (X.shape[1] + 2, 1) and len(gp.invKopt) == 1 and gp.invKopt[0].shape == (X.shape[0],) * 2
        _, _, _, _, Xn, Yn = O.normalise(X, Y)
        assert np.array_equal(gp.X_norm, Xn) and np.array_equal(gp.Y_norm, Yn)
        ours = O.nlml(gp.hypopt[:, 0], Xn, Yn[:, 0])
        ref = float(g[f'nlml_opt{i}'])
        assert ours <= ref + 1e-7 * max(1.0, abs(ref)), (i, ours, ref)       # never a worse optimum than the reference's
        if abs(ours - ref) <= 1e-9 * max(1.0, abs(ref)):
            # same basin: hyper-parameters agree as far as the flat NLML determines them
            assert np.allclose(gp.hypopt[:, 0], g[f'hypopt{i}'][:, 0], atol=2e-3), (i, gp.hypopt[:, 0], g[f'hypopt{i}'][:, 0])
        # NLML method itself
        th = g[f'nlml_theta{i}'][0]
        assert abs(gp.negative_loglikelihood(th, gp.X_norm, gp.Y_norm[:, 0]) - g[f'nlml{i}'][0]) <= 1e-9 * abs(g[f'nlml{i}'][0])


@pytest.mark.parametrize('tag,noise', [('', None), ('_kn', 1e-3)])
def test_gp_model_fit_selects_reference_start(u2, tag, noise):
    """north_star: "identical ... hyper-parameter-start indices wherever the score gap exceeds that tolerance".  Golden =
    res.x / res.fun of every SLSQP start of the reference's own multistart and the index its np.argmin picked
    (tests/golden/fit_starts.npz, utilities_2.py:156-167), free and known noise (:149-154), PLUS the same multistart of the
    reference with its NLML nudged by 1-3 ulps / 1e-15 relative noise (10 variants).  SLSQP with forward-difference
    gradients is chaotic in the last bits of its objective: under those nudges the reference's own starts hop between
    basins and its selected index changes on 9 of the 12 fits (it even finds a better optimum on one).  So:
      * where the reference's winner leads its runner-up by more than TOL (1e-7 relative: these are end points of an
        iterative optimiser, not the 1e-9 of the arithmetic) AND its selection is stable over the ensemble, our selected
        index must be identical;
      * everywhere, the selected optimum must be within TOL of an optimum the reference (nominal or nudged) reaches, or
        better (lower NLML, confirmed with the oracle's NLML)."""
    TOL = 1e-7
    g = load_golden('fit_starts.npz')
    src = load_golden('pickle_gps.npz')
    decisive = exact = better = same_end = total = 0
    for i in range(int(g['count'])):
        X, Y = src[f'X{i}'], src[f'Y{i}']
        gp = u2.GP_model(X, Y, 'RBF', multi_hyper=4, var_out=True, noise=noise)
        info = gp._fit_info[0]
        ours, ref_val, ref_idx = info['localval'], g[f'localval{tag}{i}'], int(g[f'minindex{tag}{i}'])
        ens_val, ens_idx = g[f'ens_localval{tag}{i}'], g[f'ens_minindex{tag}{i}']
        scale = max(1.0, abs(ref_val.min()))
        assert np.array_equal(gp.hypopt[:, 0], info['localsol'][info['minindex']])
        assert info['minindex'] == int(np.argmin(ours))
        exact += int(info['minindex'] == ref_idx)
        total += 4
        same_end += int((np.abs(ours - ref_val) <= TOL * np.maximum(scale, np.abs(ref_val))).sum())
        reached = np.concatenate([[ref_val.min()], ens_val.min(axis=1)])
        on_ref_optimum = np.abs(ours.min() - reached).min() <= TOL * scale
        if not on_ref_optimum:
            # must be a better optimum; confirm with the oracle's NLML that the value is real
            assert ours.min() < reached.min(), (i, tag, ours, ref_val)
            _, _, _, _, Xn, Yn = O.normalise(X, Y)
            assert abs(O.nlml(gp.hypopt[:, 0], Xn, Yn[:, 0]) - ours.min()) <= 1e-6 * scale
            better += 1
            continue
        srt = np.sort(ref_val)
        if srt[1] - srt[0] > TOL * scale and (ens_idx == ref_idx).all():
            decisive += 1
            assert info['minindex'] == ref_idx, (i, tag, ours, ref_val)
    print('\n[fit starts%s] %d of %d fits have a decisive reference winner (margin > %.0e and stable under ulp nudges): index identical on all of '
          'them; index identical to the nominal reference run on %d of %d fits overall (the rest are ties / chaotic starts); %d of %d single '
          'starts end at the reference value; better optimum than any reference variant on %d fits' %
          (tag or ' free', decisive, int(g['count']), TOL, exact, int(g['count']), same_end, total, better))
    if not tag:
        assert decisive >= 2


def test_batched_fit_api_matches_the_reference_fits(u2, handle):
    """Handle.fit (SURVEY.md par. 7.1-6 "batched sibling for B > 1"): B independent GPs fitted in ONE lock-step multistart.
    Against the reference's GP_model(...) on the six pickle data sets (golden fit_starts.npz, 3 of them share n = 24 and
    3 have n = 25, so two batches): selected optimum within 1e-7 relative of an optimum the reference reaches (or better),
    same selected start where the reference's winner is decisive, and per GP identical to the drop-in's one-at-a-time fit."""
    import torch
    g = load_golden('fit_starts.npz')
    src = load_golden('pickle_gps.npz')
    TOL = 1e-7
    groups = {}
    for i in range(int(g['count'])):
        groups.setdefault(src[f'X{i}'].shape, []).append(i)
    checked = 0
    for shape, idx in groups.items():
        X = np.stack([src[f'X{i}'] for i in idx]); Y = np.stack([src[f'Y{i}'][:, 0] for i in idx])
        out = handle.fit(X, Y, multi_hyper=4)
        assert int(out['status'].abs().sum()) == 0 and out['theta'].shape == (len(idx), shape[1] + 2)
        for j, i in enumerate(idx):
            ref_val, ref_idx = g[f'localval{i}'], int(g[f'minindex{i}'])
            ens_val, ens_idx = g[f'ens_localval{i}'], g[f'ens_minindex{i}']
            scale = max(1.0, abs(ref_val.min()))
            ours = out['localval'][j]
            assert out['minindex'][j] == int(np.argmin(ours))
            reached = np.concatenate([[ref_val.min()], ens_val.min(axis=1)])
            if np.abs(ours.min() - reached).min() > TOL * scale:
                assert ours.min() < reached.min()
                continue
            srt = np.sort(ref_val)
            if srt[1] - srt[0] > TOL * scale and (ens_idx == ref_idx).all():
                assert out['minindex'][j] == ref_idx
                checked += 1
            # the batched fit and the drop-in's single fit run the same kernels on (device- vs host-normalised) data: same optimum
            gp = u2.GP_model(src[f'X{i}'], src[f'Y{i}'], 'RBF', multi_hyper=4)
            assert abs(gp._fit_info[0]['localval'].min() - ours.min()) <= TOL * scale
        # the fitted batch is directly usable by K4
        cand = torch.as_tensor(np.stack([src[f'cand{i}'] for i in idx]), device=handle.device)
        fb = torch.as_tensor(np.array([float(src[f'f_best{i}']) for i in idx]), device=handle.device)
        post = handle.posterior_acq(out, cand, None, fb, 3)
        assert torch.isfinite(post['score']).all()
    assert checked >= 2
    # known noise pins the last hyper-parameter (utilities_2.py:149-154)
    i = groups[next(iter(groups))][0]
    out = handle.fit(src[f'X{i}'][None], src[f'Y{i}'][:, 0][None], multi_hyper=3, noise=1e-3)
    centre = 0.5 * np.log(1e-3 / float(out['y_std'][0]) ** 2)
    assert abs(float(out['theta'][0, -1]) - centre) <= 1e-3 + 1e-12


def test_gp_model_inference_api(u2):
    g = load_golden('pickle_gps.npz')
    i = 1
    X, Y = g[f'X{i}'], g[f'Y{i}']
    gp = u2.GP_model(X, Y, 'RBF', multi_hyper=2, var_out=True, noise=None, _theta=g[f'hypopt{i}'][:, 0])
    ogp = O.make_gp(X, Y[:, 0], g[f'hypopt{i}'][:, 0])
    for c in g[f'cand{i}'][:6]:
        mean, var = gp.GP_inference_np(c)
        mo, vo = O.inference(c, ogp)
        assert mean.shape == (1,) and var.shape == (1,)
        assert abs(mean[0] - mo) <= 1e-9 * abs(mo) + 1e-12 and abs(var[0] - vo) <= 1e-9 * abs(vo) + 1e-12
    gp.var_out = False
    assert isinstance(gp.GP_inference_np(g[f'cand{i}'][0]), float)
    K = gp.Cov_mat('RBF', gp.X_norm, np.exp(2 * gp.hypopt[:2, 0]), np.exp(2 * gp.hypopt[2, 0]))
    assert np.abs(K - g[f'K{i}']).max() <= 1e-12 * g[f'K{i}'].max()
    for kern in ('Matern32', 'Matern52'):            # selectable in the reference's Cov_mat (:61-66)
        Km = gp.Cov_mat(kern, gp.X_norm, np.exp(2 * gp.hypopt[:2, 0]), np.exp(2 * gp.hypopt[2, 0]))
        Ko = O.cov_mat(kern, gp.X_norm, np.exp(2 * gp.hypopt[:2, 0]), np.exp(2 * gp.hypopt[2, 0]))
        assert np.abs(Km - Ko).max() <= 1e-12 * Ko.max()
    assert gp.Cov_mat('nonsense', gp.X_norm, np.ones(2), 1.0) is None
    k = gp.calc_cov_sample((g[f'cand{i}'][0] - gp.X_mean) / gp.X_std, gp.X_norm, np.exp(2 * gp.hypopt[:2, 0]), np.exp(2 * gp.hypopt[2, 0]))
    assert k.shape == (X.shape[0], 1)
    assert np.allclose(k, O.cov_sample((g[f'cand{i}'][0] - gp.X_mean) / gp.X_std, gp.X_norm, np.exp(2 * gp.hypopt[:2, 0]), np.exp(2 * gp.hypopt[2, 0])), rtol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        Xd = np.vstack([X, X[:1]])                      # duplicate point, no noise, no jitter -> singular
        gp2 = u2.GP_model(Xd, np.vstack([Y, Y[:1]]), 'RBF', multi_hyper=1, _theta=np.array([0.0, 0.0, 0.0, -400.0]))


def test_final_factor_survives_a_numerically_singular_but_noisy_kernel_matrix(u2):
    """The reference inverts Kopt by LU (utilities_2.py:175), which does not fail on near-duplicate inputs with a tiny but
    positive noise; the drop-in's Cholesky factor retries with a diagonal shift far below the noise level instead of aborting
    the episode (exactly singular Kopt, sn2 == 0, still raises - see test_gp_model_inference_api)."""
    import warnings
    g = load_golden('pickle_gps.npz')
    X, Y = g['X1'], g['Y1']
    Xd = np.vstack([X, X[:2] + 1e-13])                     # two near-duplicates
    Yd = np.vstack([Y, Y[:2]])
    theta = g['hypopt1'][:, 0].copy()
    theta[-1] = 0.5 * np.log(1e-20)                         # noise variance 1e-20 > 0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        gp = u2.GP_model(Xd, Yd, 'RBF', multi_hyper=1, _theta=theta)
    assert np.isfinite(gp.invKopt[0]).all()
    mean, var = gp.GP_inference_np(X[5])
    assert np.isfinite(mean).all() and np.isfinite(var).all()


@pytest.mark.parametrize('tag,acq,noisy', [('noiseless_ei', 3, False), ('noisy_ei', 3, True), ('noiseless_lcb', 2, False)])
def test_rto_episode_matches_reference(u2, tag, acq, noisy):
    """Short seeded episode on the toy problem (arguments of run_simple/main_simple.py:179-206, fewer iterations):
    identical discrete signature (trust-region radii, backtrack flags); iterates within the reference's OWN
    sensitivity.  Bit-identity of the iterates is not attainable: SLSQP's forward differences (step 1.5e-8, tol 1e-12)
    amplify last-bit differences of any two correct GP evaluations (SURVEY.md par. 7.3-1).  The golden file therefore
    also holds the same episode of the reference with its NLML / posterior nudged by ONE ulp; the spread of its iterates
    under that nudge (5e-4 / 6e-8 / 7e-2 for the three cases) is the yardstick: ours must stay within 10x of it
    (floor 1e-6), i.e. the GPU path behaves like a last-bits perturbation of the NumPy path.  (Bit-identity is asserted
    where it is attainable: GPU vs the ordered CPU twin of the kernels, tests/test_gpu_twin.py.)"""
    from gp_rto_ei_b200.sub_uts import systems as S
    g = load_golden('rto_simple.npz')
    np.random.seed(0); random.seed(0)
    plant, cons = (S.Benoit_System, S.con1_system_tight) if noisy else (S.Benoit_System_noiseless, S.con1_system_tight_noiseless)
    Xtrain = np.array([[1.2, 0.], [1.4, 0.1], [1.3, -0.1]])
    with contextlib.redirect_stdout(io.StringIO()):
        rto = u2.ITR_GP_RTO(S.Benoit_Model, plant, [S.con1_model], [cons], np.array([1.1, -0.1]), 0.25, 0.7, 0.2, 0.8, 0.8, 1.2, 4,
                            ['data0', Xtrain], np.array([[-.6, 1.5], [-1., 1.]]), obj_setting=acq, noise=[1e-3] * 2 if noisy else None,
                            multi_opt=6, multi_hyper=4, TR_scaling=False, TR_curvature=False, store_data=True, inner_TR=False,
                            scale_inputs=False)
        X_opt, y_opt, TR_l, xnew, backtrack = rto.RTO_routine()
    print('\n[rto %s] max |dX| vs reference: %.3e' % (tag, np.abs(X_opt - g[tag + '_X_opt']).max()))
    if tag == 'noiseless_lcb':
        # chaotic case: the reference itself flips between two trust-region branches under 1e-13 relative perturbations of
        # its NLML (golden ensemble); ours must land on one of the reference's own branches, close to a member of it
        TRs = [np.asarray(g[tag + '_TR_l'])] + list(g[tag + '_ens_TR'])
        Xs = [g[tag + '_X_opt']] + list(g[tag + '_ens_X'])
        same = [i for i, t in enumerate(TRs) if np.allclose(t, np.array(TR_l, dtype=float), rtol=0, atol=1e-15)]
        assert same, ('trust-region sequence not among the reference branches', TR_l)
        nearest = min(np.abs(X_opt - Xs[i]).max() for i in same)
        print('[rto %s] branch members: %d, nearest member max|dX| = %.3e' % (tag, len(same), nearest))
        assert nearest < 0.05
        return
    assert np.allclose(np.array(TR_l, dtype=float), g[tag + '_TR_l'], rtol=0, atol=1e-15)
    assert list(backtrack) == list(g[tag + '_backtrack'])
    assert X_opt.shape == g[tag + '_X_opt'].shape and y_opt.shape == g[tag + '_y_opt'].shape
    sens = np.abs(g[tag + '_X_opt_ulp'] - g[tag + '_X_opt']).max()
    print('[rto %s] reference one-ulp self-sensitivity: %.3e' % (tag, sens))
    tol = max(10 * sens, 1e-6)
    assert np.abs(X_opt - g[tag + '_X_opt']).max() < tol
    if not noisy:
        assert np.abs(y_opt - g[tag + '_y_opt']).max() < 10 * tol


def test_run_main_executes_a_reference_style_driver(u2, tmp_path):
    """gp_rto_ei_b200/run_main.py: a script with the reference's import header (casadi / matplotlib / sobol_seq /
    plots_RTO star-imports, `from sub_uts.utilities_2 import *`) runs unchanged on this package; one full-size episode
    (20 RTO iterations, 30 acquisition starts, 15 hyper-parameter starts - the arguments of run_simple/main_simple.py:
    179-206) and the result pickle has the reference's structure.  The reference needs ~26 s per episode on 8 CPU
    cores (SURVEY.md par. 6)."""
    import os
    import pickle
    from conftest import REPO
    from gp_rto_ei_b200 import run_main
    with contextlib.redirect_stdout(io.StringIO()):
        glob, out = run_main.run(os.path.join(REPO, 'tests', 'data', 'mini_main_simple.py'), episodes=1, out=str(tmp_path))
    print('\n[run_main] one full episode (20 iterations): %.2f s' % glob['episode_seconds'])
    with open(os.path.join(out, 'with_prior_with_exploration_ei_noise.p'), 'rb') as fh:
        X_opt_mc, y_opt_mc, TR_l_mc, xnew_mc, backtrack_mc = pickle.load(fh)
    assert len(X_opt_mc) == 1 and X_opt_mc[0].shape == (24, 2) and y_opt_mc[0].shape == (2, 24)
    assert len(TR_l_mc[0]) == 21 and len(backtrack_mc[0]) == 21
    # solution quality anchor (BASELINE.md par. 3): the noise-free objective of the final iterate is in the range of the
    # reference's shipped runs for this setting (0.13 ... 1.06), and the iterate respects the input bounds
    from gp_rto_ei_b200.sub_uts import systems as S
    f = S.Benoit_System_noiseless(np.asarray(xnew_mc[0]).flatten())
    assert 0.05 < f < 1.2
    assert (X_opt_mc[0][:, 0] >= -0.6 - 1e-9).all() and (X_opt_mc[0][:, 0] <= 1.5 + 1e-9).all()


def test_replica_driver_is_reproducible_per_seed(u2):
    """SURVEY.md par. 8f-2: replica r is a function of its seed alone (both RNGs seeded), whatever ran before it."""
    from gp_rto_ei_b200 import replicas
    kw = dict(problem='simple', obj_setting=3, n_iter=3, multi_opt=5, multi_hyper=3)
    a = replicas.run_replica(3, **kw)
    replicas.run_replica(1, **kw)
    b = replicas.run_replica(3, **kw)
    assert a == b and set(a) >= {'objective', 'feasible', 'x', 'seed'}
    w = replicas.run_replica(0, problem='wo', obj_setting=3, n_iter=2, multi_opt=4, multi_hyper=3)
    assert 4.0 <= w['x'][0] <= 7.0 and 70.0 <= w['x'][1] <= 100.0 and -200 < w['objective'] < 200   # -profit; the optimum is -75.8


@pytest.mark.parametrize('tag,acq,noisy', [('wo_ei_noiseless', 3, False), ('wo_ei_noisy', 3, True), ('wo_mean_noisy', 1, True)])
def test_rto_episode_williams_otto_configuration(u2, tag, acq, noisy):
    """Config 2 (run_WO/main_WO_prior_EI.py): 1 + 2 GPs per iteration, scale_inputs=True (ellipse start sampling), model
    prior, known noise.  Golden = the reference's own ITR_GP_RTO / GP_model driven with the same NumPy plant callables.
    Same criterion as the toy problem: identical trust-region / backtrack signature, iterates within 10x the reference's
    one-ulp self-sensitivity (floor 1e-6 in inputs that span [4,7] x [70,100])."""
    from gp_rto_ei_b200.sub_uts import systems as S
    g = load_golden('rto_wo.npz')
    np.random.seed(0); random.seed(0)
    model, plant = S.WO_model(), S.WO_system()
    obj_sys = plant.WO_obj_sys_ca if noisy else plant.WO_obj_sys_ca_noise_less
    cons_sys = [plant.WO_con1_sys_ca, plant.WO_con2_sys_ca] if noisy else [plant.WO_con1_sys_ca_noise_less, plant.WO_con2_sys_ca_noise_less]
    Xtrain = np.array([[5.7, 74.], [6.35, 74.9], [6.6, 75.], [6.75, 79.]])
    with contextlib.redirect_stdout(io.StringIO()):
        rto = u2.ITR_GP_RTO(model.WO_obj_ca, obj_sys, [model.WO_con1_model_ca, model.WO_con2_model_ca], cons_sys, np.array([6.9, 83.]), 0.25,
                            0.7, 0.2, 0.8, 0.8, 1.2, 3, ['data0', Xtrain], np.array([[4., 7.], [70., 100.]]), obj_setting=acq,
                            noise=[0.5 ** 2, 5e-8, 5e-8], multi_opt=5, multi_hyper=3, TR_scaling=False, TR_curvature=False, store_data=True,
                            inner_TR=False, scale_inputs=True)
        X_opt, y_opt, TR_l, xnew, backtrack = rto.RTO_routine()
    sens = np.abs(g[tag + '_X_opt_ulp'] - g[tag + '_X_opt']).max()
    err = np.abs(X_opt - g[tag + '_X_opt']).max()
    print('\n[rto %s] max |dX| vs reference: %.3e (reference one-ulp self-sensitivity %.3e)' % (tag, err, sens))
    assert np.allclose(np.array(TR_l, dtype=float), g[tag + '_TR_l'], rtol=0, atol=1e-15)
    assert list(backtrack) == list(g[tag + '_backtrack'])
    assert err < max(10 * sens, 1e-6)
