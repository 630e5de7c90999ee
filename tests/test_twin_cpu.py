"""The ordered CPU twin (oracle/twin/) on its own, without a GPU: its deterministic scalar functions against mpmath, and the
three restated kernels against the oracle (they are the same mathematics in a different, fixed operation order, so they
agree with the NumPy restatement to rounding level - the BIT-level comparison against the B200 is tests/test_gpu_twin.py)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import gp_oracle as O
import twin_handle


@pytest.fixture(scope='module')
def twin():
    return twin_handle.Twin()


def test_deterministic_scalar_functions_are_accurate(twin):
    mp = pytest.importorskip('mpmath')
    mp.mp.dps = 40
    rng = np.random.default_rng(0)
    L = twin.lib

    def worst(fn, ref, xs):
        return max(float(abs(mp.mpf(fn(float(x))) - ref(mp.mpf(float(x)))) / abs(ref(mp.mpf(float(x))))) for x in xs)
    assert worst(L.twin_exp, mp.exp, np.concatenate([rng.uniform(-700, 0, 1500), rng.uniform(0, 20, 500)])) < 2.5e-16
    assert worst(L.twin_log, mp.log, np.exp(rng.uniform(-40, 40, 2000))) < 4e-16
    assert worst(L.twin_rsqrt, lambda x: 1 / mp.sqrt(x), np.exp(rng.uniform(-40, 40, 2000))) < 3e-16
    # erfc(x) = exp(-x^2) p(t) / (1 + 2x): the polynomial part is good to 2e-16, exp(-x^2) carries x^2 * eps
    assert worst(L.twin_erfc, mp.erfc, rng.uniform(-6, 3, 2000)) < 3e-15
    assert worst(L.twin_erfc, mp.erfc, rng.uniform(3, 26, 1000)) < 1e-13
    assert L.twin_erfc(30.0) == 0.0 and L.twin_erfc(-30.0) == 2.0 and L.twin_exp(-800.0) == 0.0 and L.twin_log(1.0) == 0.0
    assert np.isnan(L.twin_log(-1.0)) and np.isinf(L.twin_exp(710.0))


def test_twin_kernels_agree_with_the_oracle(twin):
    g = load_golden('pickle_gps.npz')
    for i in range(int(g['count'])):
        X, Y, theta = g[f'X{i}'], g[f'Y{i}'][:, 0], g[f'hypopt{i}'][:, 0]
        gp = O.make_gp(X, Y, theta)
        cond = np.linalg.cond(gp['invK'])
        # NLML at the reference's Sobol starts (the reference raises where the twin reports a status)
        for th, ref, fail in zip(g[f'nlml_theta{i}'], g[f'nlml{i}'], g[f'nlml_fail{i}']):
            v, st = twin.nlml(gp['X_norm'], gp['y_norm'], th)
            if not fail and st == 0:
                # same bound as the GPU tests: 1e-9 relative + the first-order rounding bound of the quantity, 20 eps cond(K) |y'K^-1y|
                W, sf2, sn2 = O.unpack_theta(th, X.shape[1])
                K = O.cov_mat('RBF', gp['X_norm'], W, sf2) + (sn2 + 1e-8) * np.eye(len(Y))
                quad = abs(gp['y_norm'] @ np.linalg.solve(K, gp['y_norm']))
                assert abs(v - ref) <= 1e-9 * max(1.0, abs(ref)) + 20 * np.finfo(float).eps * np.linalg.cond(K) * quad
        invK, alpha, logdet, st = twin.factor(gp['X_norm'], gp['y_norm'], theta)
        assert st == 0 and np.array_equal(invK, invK.T)
        assert np.abs(invK - gp['invK']).max() <= 50 * cond * np.finfo(float).eps * np.abs(gp['invK']).max()
        mean, var, score = twin.posterior(gp['X_norm'], gp['invK'], gp['invK'] @ gp['y_norm'], theta, gp['X_mean'], gp['X_std'], gp['Y_mean'],
                                          gp['Y_std'], g[f'cand{i}'], g[f'prior{i}'], float(g[f'f_best{i}']), 3)
        assert np.allclose(mean, g[f'mean{i}'], rtol=1e-9, atol=1e-9 * np.abs(g[f'mean{i}']).max())
        assert np.allclose(var, g[f'var{i}'], rtol=1e-7, atol=cond * 1e-15 * gp['Y_std'] ** 2)
        assert np.allclose(score, g[f'score{i}'][2], rtol=1e-6, atol=1e-9)


def test_drop_in_runs_an_episode_on_the_twin(monkeypatch):
    """The drop-in ITR_GP_RTO on the twin handle (CPU): same discrete signature as the reference's golden episode and
    iterates within 10x the reference's own one-ulp self-sensitivity."""
    import contextlib
    import io
    import random
    from gp_rto_ei_b200 import api
    from gp_rto_ei_b200.sub_uts import systems as S
    from gp_rto_ei_b200.sub_uts import utilities_2 as u2
    th = twin_handle.TwinHandle()
    monkeypatch.setattr(api, 'default_handle', lambda device=None: th)
    g = load_golden('rto_simple.npz')
    np.random.seed(0); random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        rto = u2.ITR_GP_RTO(S.Benoit_Model, S.Benoit_System, [S.con1_model], [S.con1_system_tight], np.array([1.1, -0.1]), 0.25, 0.7, 0.2, 0.8, 0.8,
                            1.2, 4, ['data0', np.array([[1.2, 0.], [1.4, 0.1], [1.3, -0.1]])], np.array([[-.6, 1.5], [-1., 1.]]), obj_setting=3,
                            noise=[1e-3] * 2, multi_opt=6, multi_hyper=4, TR_scaling=False, TR_curvature=False, store_data=True, inner_TR=False,
                            scale_inputs=False)
        X_opt, y_opt, TR_l, xnew, backtrack = rto.RTO_routine()
    assert np.allclose(np.array(TR_l, dtype=float), g['noisy_ei_TR_l'], rtol=0, atol=1e-15) and list(backtrack) == list(g['noisy_ei_backtrack'])
    sens = np.abs(g['noisy_ei_X_opt_ulp'] - g['noisy_ei_X_opt']).max()
    assert np.abs(X_opt - g['noisy_ei_X_opt']).max() < max(10 * sens, 1e-6)
