"""cProfile of one full RTO episode of the drop-in (arguments of run_simple/main_simple.py:179-206)."""
import cProfile
import contextlib
import io
import os
import pstats
import random
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
from gp_rto_ei_b200.sub_uts import utilities_2 as u2, systems as S  # noqa: E402


def episode(seed=0, n_iter=20):
    np.random.seed(seed); random.seed(seed)
    Xtrain = np.array([[1.2, 0.], [1.4, 0.1], [1.3, -0.1]])
    with contextlib.redirect_stdout(io.StringIO()):
        rto = u2.ITR_GP_RTO(S.Benoit_Model, S.Benoit_System, [S.con1_model], [S.con1_system_tight], np.array([1.1, -0.1]), 0.25, 0.7, 0.2,
                            0.8, 0.8, 1.2, n_iter, ['data0', Xtrain], np.array([[-.6, 1.5], [-1., 1.]]), obj_setting=3, noise=[1e-3] * 2,
                            multi_opt=30, multi_hyper=15, TR_scaling=False, TR_curvature=False, store_data=True, inner_TR=False,
                            scale_inputs=False)
        return rto.RTO_routine()


episode(1, 2)      # warm-up (library load, kernels)
t0 = time.time(); episode(0); print('episode wall %.2f s' % (time.time() - t0))
pr = cProfile.Profile(); pr.enable(); episode(0); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)

# scalar API latency (the reference needs ~50 us per GP_inference_np call, SURVEY.md par. 6)
rng = np.random.default_rng(0)
X = rng.uniform(-1, 1, (24, 2)); Y = np.sin(X.sum(1, keepdims=True))
gp = u2.GP_model(X, Y, 'RBF', multi_hyper=2, var_out=True, noise=None)
x = np.array([0.1, 0.2])
for _ in range(50):
    gp.GP_inference_np(x)
t0 = time.perf_counter()
for _ in range(2000):
    gp.GP_inference_np(x)
print('GP_inference_np scalar call: %.1f us' % ((time.perf_counter() - t0) / 2000 * 1e6))
t0 = time.perf_counter()
for _ in range(500):
    gp.negative_loglikelihood(gp.hypopt[:, 0], gp.X_norm, gp.Y_norm[:, 0])
print('negative_loglikelihood scalar call: %.1f us' % ((time.perf_counter() - t0) / 500 * 1e6))
