"""GPU <-> ordered CPU twin: BIT-IDENTICAL results (SURVEY.md par. 7.3-1a; north_star "bit-identical RTO iterate trajectories").

Bit-identity against the reference's NumPy path is not attainable by any independent implementation (the reference's own
episode moves by up to 7e-2 under a one-ulp nudge of its NLML, tests/golden/rto_simple.npz).  What IS attainable is a GPU
path whose arithmetic is specified: IEEE operations in a fixed order, own transcendental functions
(gp_rto_ei_b200/csrc/gprto_det_math.h), -fmad=false.  oracle/twin/gprto_twin.cpp builds the same header for the host and
restates the three kernels the RTO loop uses; these tests require equality of every bit
  * kernel by kernel on seeded inputs (NLML for n = 2 .. 32, final factor, posterior + all three acquisitions), and
  * of complete RTO episodes - the six golden configurations (toy problem and Williams-Otto) run once on the B200 and once
    on the twin through the very same drop-in code."""
import contextlib
import io
import random

import numpy as np
import pytest
import torch

import twin_handle
from gp_rto_ei_b200 import api, workloads

pytestmark = pytest.mark.gpu


def T(a, dev, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=dev)


@pytest.fixture(scope='module')
def twin():
    return twin_handle.Twin()


@pytest.mark.parametrize('n,d', [(2, 1), (4, 2), (5, 2), (8, 2), (9, 3), (16, 2), (17, 2), (24, 2), (25, 2), (32, 3), (31, 5)])
def test_nlml_kernel_is_bit_identical_to_the_twin(handle, dev, twin, n, d):
    w = workloads.cfg4_numpy(seed=n * 10 + d, B=3, n=n, d=d, S=16)
    nz = handle.normalise(T(w['X'], dev), T(w['y'], dev))
    nll, st = handle.nlml(nz['Xn'], nz['Yn'], T(w['theta'], dev))
    Xn, Yn = nz['Xn'].cpu().numpy(), nz['Yn'].cpu().numpy()
    nll, st = nll.cpu().numpy(), st.cpu().numpy()
    ok = 0
    for b in range(3):
        for s in range(16):
            v, info = twin.nlml(Xn[b], Yn[b], w['theta'][b, s])
            assert info == st[b, s]
            if info == 0:
                assert v == nll[b, s], (n, d, b, s, v, nll[b, s])
                ok += 1
    assert ok >= 24


@pytest.mark.parametrize('n,d', [(3, 1), (5, 2), (8, 2), (16, 2), (25, 2), (32, 3)])
def test_factor_kernel_is_bit_identical_to_the_twin(handle, dev, twin, n, d):
    w = workloads.cfg3_numpy(seed=n + d, B=3, n=n, d=d, m=4)
    nz = handle.normalise(T(w['X'], dev), T(w['y'], dev))
    invK, alpha, logdet, st = handle.factor(nz['Xn'], nz['Yn'], T(w['theta'], dev))
    Xn, Yn = nz['Xn'].cpu().numpy(), nz['Yn'].cpu().numpy()
    for b in range(3):
        ik, al, ld, info = twin.factor(Xn[b], Yn[b], w['theta'][b])
        assert info == int(st[b]) == 0
        assert np.array_equal(ik, invK[b].cpu().numpy()) and np.array_equal(al, alpha[b].cpu().numpy()) and ld == float(logdet[b])


@pytest.mark.parametrize('n,d,m', [(4, 2, 33), (8, 2, 64), (9, 2, 17), (25, 2, 100), (24, 3, 40), (64, 3, 257), (100, 4, 31), (128, 3, 64)])
def test_posterior_kernel_is_bit_identical_to_the_twin(handle, dev, twin, n, d, m):
    w = workloads.cfg3_numpy(seed=3 * n + m, B=2, n=n, d=d, m=m)
    gd = handle.fit_fixed(T(w['X'], dev), T(w['y'], dev), T(w['theta'], dev))
    prior = np.random.default_rng(n).normal(size=(2, m))
    h = {k: v.cpu().numpy() for k, v in gd.items()}
    for acq in (1, 2, 3):
        out = handle.posterior_acq(gd, T(w['cand'], dev), T(prior, dev), T(w['f_best'], dev), acq)
        for b in range(2):
            mean, var, score = twin.posterior(h['Xn'][b], h['invK'][b], h['alpha'][b], w['theta'][b], h['x_mean'][b], h['x_std'][b],
                                              h['y_mean'][b], h['y_std'][b], w['cand'][b], prior[b], w['f_best'][b], acq)
            assert np.array_equal(mean, out['mean'][b].cpu().numpy()), (n, d, m, acq, 'mean')
            assert np.array_equal(var, out['var'][b].cpu().numpy()), (n, d, m, acq, 'var')
            assert np.array_equal(score, out['score'][b].cpu().numpy()), (n, d, m, acq, 'score')


def _episode(u2, problem, acq, noisy, n_iter, multi_opt, multi_hyper):
    from gp_rto_ei_b200.sub_uts import systems as S
    np.random.seed(0); random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        if problem == 'simple':
            plant, cons = (S.Benoit_System, S.con1_system_tight) if noisy else (S.Benoit_System_noiseless, S.con1_system_tight_noiseless)
            rto = u2.ITR_GP_RTO(S.Benoit_Model, plant, [S.con1_model], [cons], np.array([1.1, -0.1]), 0.25, 0.7, 0.2, 0.8, 0.8, 1.2, n_iter,
                                ['data0', np.array([[1.2, 0.], [1.4, 0.1], [1.3, -0.1]])], np.array([[-.6, 1.5], [-1., 1.]]), obj_setting=acq,
                                noise=[1e-3] * 2 if noisy else None, multi_opt=multi_opt, multi_hyper=multi_hyper, TR_scaling=False,
                                TR_curvature=False, store_data=True, inner_TR=False, scale_inputs=False)
        else:
            model, plant = S.WO_model(), S.WO_system()
            obj_sys = plant.WO_obj_sys_ca if noisy else plant.WO_obj_sys_ca_noise_less
            cons_sys = [plant.WO_con1_sys_ca, plant.WO_con2_sys_ca] if noisy else [plant.WO_con1_sys_ca_noise_less, plant.WO_con2_sys_ca_noise_less]
            rto = u2.ITR_GP_RTO(model.WO_obj_ca, obj_sys, [model.WO_con1_model_ca, model.WO_con2_model_ca], cons_sys, np.array([6.9, 83.]), 0.25,
                                0.7, 0.2, 0.8, 0.8, 1.2, n_iter, ['data0', np.array([[5.7, 74.], [6.35, 74.9], [6.6, 75.], [6.75, 79.]])],
                                np.array([[4., 7.], [70., 100.]]), obj_setting=acq, noise=[0.5 ** 2, 5e-8, 5e-8], multi_opt=multi_opt,
                                multi_hyper=multi_hyper, TR_scaling=False, TR_curvature=False, store_data=True, inner_TR=False, scale_inputs=True)
        return rto.RTO_routine()


@pytest.mark.parametrize('problem,acq,noisy,n_iter,multi_opt,multi_hyper',
                         [('simple', 3, False, 4, 6, 4), ('simple', 3, True, 4, 6, 4), ('simple', 2, False, 4, 6, 4),
                          ('wo', 3, False, 3, 5, 3), ('wo', 3, True, 3, 5, 3), ('wo', 1, True, 3, 5, 3)])
def test_rto_trajectories_are_bit_identical_gpu_vs_twin(handle, monkeypatch, problem, acq, noisy, n_iter, multi_opt, multi_hyper):
    """The six golden episode configurations (tests/golden/rto_simple.npz, rto_wo.npz), once on the B200 and once on the
    ordered CPU twin, through the same drop-in ITR_GP_RTO / GP_model code: every iterate, measurement, trust-region radius and
    backtrack flag identical to the last bit - including the chaotic LCB case where the reference itself flips branches under
    1e-13 perturbations."""
    from gp_rto_ei_b200.sub_uts import utilities_2 as u2
    gpu = _episode(u2, problem, acq, noisy, n_iter, multi_opt, multi_hyper)
    th = twin_handle.TwinHandle()
    monkeypatch.setattr(api, 'default_handle', lambda device=None: th)
    cpu = _episode(u2, problem, acq, noisy, n_iter, multi_opt, multi_hyper)
    assert np.array_equal(np.asarray(gpu[2], dtype=float), np.asarray(cpu[2], dtype=float))          # TR_l
    assert list(gpu[4]) == list(cpu[4])                                                              # backtrack
    assert np.array_equal(gpu[0], cpu[0]), np.abs(gpu[0] - cpu[0]).max()                             # X_opt
    assert np.array_equal(gpu[1], cpu[1])                                                            # y_opt
    assert np.array_equal(np.asarray(gpu[3]), np.asarray(cpu[3]))                                    # xnew
