"""Plant / model callables (gp_rto_ei_b200/sub_uts/systems.py) against the constants the reference publishes.

The reference solves the Williams-Otto balances with casadi's Newton rootfinder (sub_uts/systems.py:101-131, 240-256);
casadi is absent here, so systems.py restates the same equations with a NumPy Newton iteration.  These tests pin that
restatement to the reference's own published anchors: the "Real Optimum" 275.811 = 200 - f* of plots_RTO.py:462 at
u* = (4.389, 80.49) with both plant constraints active, and the toy optimum 0.145 at (0.36730946, -0.39393939) of
plots_RTO.py:337 / run_simple/plots_RTO_simple.py:462.  CPU only.
"""
import numpy as np
from scipy.optimize import minimize

from gp_rto_ei_b200.sub_uts import systems as S


def test_williams_otto_plant_optimum_matches_published_constant():
    plant = S.WO_system()
    cons = [{'type': 'ineq', 'fun': plant.WO_con1_sys_ca_noise_less}, {'type': 'ineq', 'fun': plant.WO_con2_sys_ca_noise_less}]
    best = None
    for u0 in ([6.9, 83.], [5.0, 80.], [4.5, 85.]):
        r = minimize(lambda u: float(plant.WO_obj_sys_ca_noise_less(u)), np.array(u0), method='SLSQP', bounds=[(4., 7.), (70., 100.)],
                     constraints=cons, tol=1e-12)
        if r.success and (best is None or r.fun < best.fun):
            best = r
    assert best is not None
    # the published constant is the plant objective AT the published (rounded) optimum (marker plots_RTO.py:189, line :462)
    u_pub = np.array([4.389, 80.49])
    assert abs((200. - float(plant.WO_obj_sys_ca_noise_less(u_pub))) - 275.811) < 5e-4
    assert abs(float(plant.WO_con1_sys_ca_noise_less(u_pub))) < 5e-5 and abs(float(plant.WO_con2_sys_ca_noise_less(u_pub))) < 5e-5
    # and the exact constrained optimum of the restated plant is that point to the published digits
    assert np.allclose(best.x, u_pub, atol=6e-3), best.x
    assert abs((200. - best.fun) - 275.811) < 1.5e-2, 200. - best.fun
    # both constraints are active at the optimum (the optimum sits in the corner of g1 = Xa <= 0.12 and g2 = Xg <= 0.08)
    assert abs(float(plant.WO_con1_sys_ca_noise_less(best.x))) < 1e-7 and abs(float(plant.WO_con2_sys_ca_noise_less(best.x))) < 1e-7


def test_williams_otto_balances_close_and_fractions_sum_to_one():
    plant, model = S.WO_system(), S.WO_model()
    rng = np.random.default_rng(0)
    for _ in range(20):
        u = np.array([rng.uniform(4., 7.), rng.uniform(70., 100.)])
        w = plant.eval(plant.W0, u).toarray().reshape(-1)
        F, _ = plant._balances(w, u)
        assert np.abs(F).max() < 1e-13 and abs(w.sum() - 1.0) < 1e-10 and (w > 0).all()
        wm = model.eval(model.W0, u).toarray().reshape(-1)
        Fm, _ = model._balances(wm, u)
        assert np.abs(Fm).max() < 1e-10 and abs(wm.sum() - 1.0) < 1e-10 and (wm > 0).all()


def test_williams_otto_jacobians_match_finite_differences():
    plant, model = S.WO_system(), S.WO_model()
    u = np.array([5.3, 84.])
    for obj in (plant, model):
        w = np.array(obj.W0) * 1.1
        F, J = obj._balances(w, u)
        Jfd = np.empty_like(J)
        for j in range(len(w)):
            e = np.zeros_like(w); e[j] = 1e-7
            Jfd[:, j] = (obj._balances(w + e, u)[0] - obj._balances(w - e, u)[0]) / 2e-7
        assert np.abs(J - Jfd).max() <= 1e-6 * np.abs(J).max()


def test_model_has_structural_mismatch_but_similar_scale():
    plant, model = S.WO_system(), S.WO_model()
    u = np.array([4.389, 80.49])
    fp, fm = float(plant.WO_obj_sys_ca_noise_less(u)), float(model.WO_obj_ca(u))
    assert abs(fp - fm) > 1.0 and abs(fp - fm) < 150.0


def test_benoit_optimum_matches_published_constant():
    cons = [{'type': 'ineq', 'fun': S.con1_system_tight_noiseless}]
    r = minimize(S.Benoit_System_noiseless, np.array([1.1, -0.1]), method='SLSQP', bounds=[(-.6, 1.5), (-1., 1.)], constraints=cons, tol=1e-13)
    assert r.success
    assert abs(r.fun - 0.145) < 5e-4                                            # plots_RTO.py:337
    assert np.allclose(r.x, [0.36730946, -0.39393939], atol=2e-3), r.x           # run_simple/plots_RTO_simple.py:462
    assert abs(S.con1_system_tight_noiseless(r.x)) < 1e-8


def test_noisy_plants_draw_exactly_one_normal_per_call():
    """The RTO trajectories depend on the RNG consumption order (reference systems.py:48-60, 149-187)."""
    plant = S.WO_system()
    u = np.array([5.0, 80.0])
    for f, scale in ((S.Benoit_System, np.sqrt(1e-3)), (S.con1_system_tight, -np.sqrt(1e-3)), (plant.WO_obj_sys_ca, 0.5),
                     (plant.WO_con1_sys_ca, -5e-4), (plant.WO_con2_sys_ca, -5e-4)):
        np.random.seed(5)
        v = float(np.asarray(f(u)).reshape(-1)[0])
        state_after = np.random.get_state()[1].copy()
        np.random.seed(5)
        z = np.random.normal(0., 1.)
        assert np.array_equal(np.random.get_state()[1], state_after)
        np.random.seed(5)
        v2 = float(np.asarray(f(u)).reshape(-1)[0])
        assert v == v2 and np.isfinite(v) and z == z and scale != 0
