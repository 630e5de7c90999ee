"""Host logic of the lock-step replica driver without a GPU: both the sequential drop-in (ITR_GP_RTO.RTO_routine) and
gp_rto_ei_b200/lockstep.py are run on a CPU stand-in for the kernel handle (tests/fake_handle.py, oracle arithmetic) and
must give BIT-IDENTICAL episodes per seed - same SLSQP bookkeeping, same consumption order of both random streams, same
padding / scatter of the batched evaluations.  (The same property on the real kernels: tests/test_gpu_lockstep.py.)"""
import contextlib
import io
import random

import numpy as np
import pytest

from fake_handle import FakeHandle


@pytest.fixture()
def fake(monkeypatch):
    from gp_rto_ei_b200 import api
    h = FakeHandle()
    monkeypatch.setattr(api, 'default_handle', lambda device=None: h)
    return h


def sequential(seed, problem, obj, n_iter, multi_opt, multi_hyper):
    from gp_rto_ei_b200 import replicas
    return replicas.run_replica(seed, problem=problem, obj_setting=obj, n_iter=n_iter, multi_opt=multi_opt, multi_hyper=multi_hyper, trajectory=True)


@pytest.mark.parametrize('problem,seeds,n_iter,multi_opt,multi_hyper,obj', [('simple', [3, 1, 7], 3, 4, 2, 3), ('wo', [0, 4], 2, 3, 2, 3),
                                                                            ('simple', [5, 2], 2, 3, 2, 2)])
def test_lockstep_equals_sequential_on_the_cpu_stand_in(fake, problem, seeds, n_iter, multi_opt, multi_hyper, obj):
    from gp_rto_ei_b200 import lockstep
    ls = lockstep.LockstepRTO(problem, seeds, obj_setting=obj, n_iter=n_iter, multi_opt=multi_opt, multi_hyper=multi_hyper, handle=fake)
    out = ls.run()
    f, feas = ls.outcome(out)
    for j, seed in enumerate(seeds):
        seq = sequential(seed, problem, obj, n_iter, multi_opt, multi_hyper)
        assert np.array_equal(out['TR_l'][j], seq['TR_l'])
        assert np.array_equal(out['backtrack'][j], seq['backtrack'])
        assert np.array_equal(out['X_opt'][j], seq['X_opt']), np.abs(out['X_opt'][j] - seq['X_opt']).max()
        assert np.array_equal(out['y_opt'][j], seq['y_opt'])
        assert f[j] == seq['objective'] and bool(feas[j]) == seq['feasible']
    assert ls.stats['slsqp_calls_fit'] > 0 and ls.stats['slsqp_calls_acq'] > 0


def test_replica_rng_streams_match_the_global_generators():
    """RandomState(seed) / random.Random(seed) reproduce np.random.seed(seed) / random.seed(seed) call for call, including the
    mix of scalar normals, vector normals, rand and randn the episode draws."""
    rs, pr = np.random.RandomState(9), random.Random(9)
    np.random.seed(9); random.seed(9)
    a = [np.random.normal(0., 0.3) for _ in range(5)] + list(np.random.normal(0, 1, 2)) + [random.random()] + \
        list(np.random.normal(size=(2, 3)).ravel()) + list(np.random.rand(1, 3).ravel()) + [np.random.randn()] + [np.random.normal(0., 1)]
    z = rs.standard_normal(5)
    b = [0. + 0.3 * v for v in z] + list(rs.normal(0, 1, 2)) + [pr.random()] + list(rs.normal(size=(2, 3)).ravel()) + \
        list(rs.rand(1, 3).ravel()) + [rs.randn()] + [rs.standard_normal()]
    assert a == b
