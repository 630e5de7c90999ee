"""Lock-step Monte-Carlo RTO (gp_rto_ei_b200/lockstep.py, BASELINE.json configs[4]) against the sequential drop-in.

A lock-step replica must be BIT-IDENTICAL to the same seed run on its own through ITR_GP_RTO.RTO_routine
(gp_rto_ei_b200/replicas.run_replica): same kernels per problem, same NumPy arithmetic per point, same consumption order
of the replica's two random streams (utilities_2.py:646-647, 678-730, 876-883, 932, 999-1001).  This is the property
that makes 8,192 replicas in lock step the same experiment as 8,192 sequential episodes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('problem,seeds,n_iter,multi_opt,multi_hyper,obj', [('simple', [3, 1, 7, 11], 4, 6, 4, 3), ('simple', [5, 2], 3, 5, 3, 2),
                                                                            ('wo', [0, 4, 9], 3, 5, 3, 3), ('wo', [2, 6], 2, 4, 3, 1)])
def test_lockstep_replicas_are_bit_identical_to_sequential_episodes(handle, problem, seeds, n_iter, multi_opt, multi_hyper, obj):
    from gp_rto_ei_b200 import lockstep, replicas
    ls = lockstep.LockstepRTO(problem, seeds, obj_setting=obj, n_iter=n_iter, multi_opt=multi_opt, multi_hyper=multi_hyper, handle=handle)
    out = ls.run()
    f, feas = ls.outcome(out)
    assert not out['dead'].any()
    for j, seed in enumerate(seeds):
        seq = replicas.run_replica(seed, problem=problem, obj_setting=obj, n_iter=n_iter, multi_opt=multi_opt, multi_hyper=multi_hyper,
                                   trajectory=True)
        assert np.array_equal(out['TR_l'][j], seq['TR_l']), (problem, seed, out['TR_l'][j], seq['TR_l'])
        assert np.array_equal(out['backtrack'][j], seq['backtrack'])
        assert np.array_equal(out['X_opt'][j], seq['X_opt']), (problem, seed, np.abs(out['X_opt'][j] - seq['X_opt']).max())
        assert np.array_equal(out['y_opt'][j], seq['y_opt'])
        assert f[j] == seq['objective'] and bool(feas[j]) == seq['feasible']
    print('\n[lockstep %s] %d replicas x %d iterations bit-identical to the sequential drop-in; SLSQP iterations fit/acq %d/%d in %d/%d rounds' %
          (problem, len(seeds), n_iter, ls.stats['slsqp_calls_fit'], ls.stats['slsqp_calls_acq'], ls.stats['rounds_fit'], ls.stats['rounds_acq']))


def test_lockstep_is_independent_of_the_batch_composition(handle):
    """Replica r depends on its seed alone: the same seed gives the same episode whatever other replicas share the batch."""
    from gp_rto_ei_b200 import lockstep
    kw = dict(obj_setting=3, n_iter=3, multi_opt=5, multi_hyper=3, handle=handle)
    a = lockstep.LockstepRTO('simple', [4, 8, 15, 16, 23, 42], **kw).run()
    b = lockstep.LockstepRTO('simple', [42, 4], **kw).run()
    assert np.array_equal(a['X_opt'][5], b['X_opt'][0]) and np.array_equal(a['X_opt'][0], b['X_opt'][1])
    assert np.array_equal(a['TR_l'][5], b['TR_l'][0])
