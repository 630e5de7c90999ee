# A miniature driver in the style of the reference's run_simple/main_simple.py (same import header, one experiment
# block with the arguments of its "EI - prior - known noise" block), used to test gp_rto_ei_b200/run_main.py on the
# GPU box, where the reference tree itself is not available.
import time
import random
import numpy as np
import numpy.random as rnd
from scipy.spatial.distance import cdist
import sobol_seq
from scipy.optimize import minimize
import scipy
import matplotlib.pyplot as plt
import functools
from matplotlib.patches import Ellipse
from casadi import *
from sub_uts.utilities_2 import *
from sub_uts.systems import *
import pickle
from plots_RTO import plot_obj, Plot_simple, compute_obj_simple

if not (os.path.exists('figs')):
    os.mkdir('figs')

np.random.seed(0)
random.seed(0)
X_opt_mc, y_opt_mc, TR_l_mc, xnew_mc, backtrack_1_mc = [], [], [], [], []
t_start = time.time()
for i in range(30):
    n_iter = 20
    bounds = [[-.6, 1.5], [-1., 1.]]
    Xtrain = np.array([[1.2, 0.], [1.4, 0.1], [1.3, -0.1]])
    data = ['data0', Xtrain]
    u0 = np.array([1.1, -0.1])
    ITR_GP_opt = ITR_GP_RTO(Benoit_Model, Benoit_System, [con1_model], [con1_system_tight], u0, 0.25,
                            0.7, 0.2, 0.8, 0.8, 1.2,
                            n_iter, data, np.array(bounds), obj_setting=3, noise=[1e-3] * 2, multi_opt=30,
                            multi_hyper=15, TR_scaling=False, TR_curvature=False,
                            store_data=True, inner_TR=False, scale_inputs=False)
    print('Episode: ', i)
    X_opt, y_opt, TR_l, xnew, backtrack_l = ITR_GP_opt.RTO_routine()
    X_opt_mc += [X_opt]
    y_opt_mc += [y_opt]
    TR_l_mc += [TR_l]
    xnew_mc += [xnew]
    backtrack_1_mc += [[False, *backtrack_l]]
episode_seconds = (time.time() - t_start) / len(X_opt_mc)
pickle.dump([X_opt_mc, y_opt_mc, TR_l_mc, xnew_mc, backtrack_1_mc], open('with_prior_with_exploration_ei_noise.p', 'wb'))
Plot_simple('with_prior_with_exploration_ei_noise')
