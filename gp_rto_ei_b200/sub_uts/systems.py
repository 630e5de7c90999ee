"""Plant / model callables of the two benchmark problems - mirror of the reference's ``sub_uts/systems.py``.

Host-side test-problem definitions (SURVEY.md par. 2.1 rows 17-19): plain callables ``f(u) -> float``.
  * Benoit toy problem (reference systems.py:23-60): quadratic objective and one constraint; the "system" versions
    add N(0, 1e-3) measurement noise with ONE ``np.random.normal`` draw per call, which the RTO trajectories depend on.
  * Williams-Otto reactor (reference systems.py:63-297): steady-state mass balances, 6 species for the plant and a
    5-species, 2-reaction structural-mismatch model.  The reference solves them with casadi's Newton rootfinder from
    a fixed initial guess; casadi is not available here, so the same equations are solved by a dense Newton
    iteration with the analytic Jacobian (numpy) from the same initial guesses.  Results agree with the published
    optimum (profit 75.81 at u = (4.39, 80.5), reference plots_RTO.py:462) - see tests/test_systems.py.
"""
import numpy as np

_NOISE_STD = np.sqrt(1e-3)


# ------------------------------------------------------------------------------------------ Benoit toy problem
def _benoit_plant(u):
    return u[0] ** 2 + u[1] ** 2 + u[0] * u[1]


def Benoit_Model(u):
    return u[0] ** 2 + u[1] ** 2


def con1_model(u):
    return -(1. - u[0] + u[1] ** 2)


def Benoit_System(u):
    return _benoit_plant(u) + np.random.normal(0., _NOISE_STD)


def con1_system(u):
    return -(1. - u[0] + u[1] ** 2 + 2. * u[1] - 2. + np.random.normal(0., _NOISE_STD))


def con1_system_tight(u):
    return -(1. - u[0] + u[1] ** 2 + 2. * u[1] + np.random.normal(0., _NOISE_STD))


def Benoit_System_noiseless(u):
    return _benoit_plant(u)


def con1_system_noiseless(u):
    return -(1. - u[0] + u[1] ** 2 + 2. * u[1] - 2.)


def con1_system_tight_noiseless(u):
    return -(1. - u[0] + u[1] ** 2 + 2. * u[1])


def con_empty(u):
    return 0.


def obj_empty(u):
    return 0.


# Batch forms used by the lock-step replica driver (gp_rto_ei_b200/lockstep.py); each column is bit-identical to the scalar
# callable above on the same point (same operations in the same order, elementwise).
def benoit_model_batch(U):
    """[Benoit_Model, con1_model] at U[k, 2]."""
    U = np.asarray(U, dtype=float).reshape(-1, 2)
    u0, u1 = U[:, 0], U[:, 1]
    return np.stack([_sq(u0) + _sq(u1), -(1. - u0 + _sq(u1))], axis=1)


def benoit_plant_batch(U, Z=None):
    """[Benoit_System, con1_system_tight] at U[k, 2]; Z[k, 2] standard-normal draws (None: noise-free)."""
    U = np.asarray(U, dtype=float).reshape(-1, 2)
    u0, u1 = U[:, 0], U[:, 1]
    f = _sq(u0) + _sq(u1) + u0 * u1
    if Z is None:
        return np.stack([f, -(1. - u0 + _sq(u1) + 2. * u1)], axis=1)
    return np.stack([f + (0. + _NOISE_STD * Z[:, 0]), -(1. - u0 + _sq(u1) + 2. * u1 + (0. + _NOISE_STD * Z[:, 1]))], axis=1)


def _sq(a):
    """Elementwise `** 2` with the rounding of the scalar callables: a NumPy scalar ** 2 goes through pow(), an array ** 2
    through multiply - np.float_power keeps the pow() path for arrays."""
    return np.float_power(a, 2)


# ------------------------------------------------------------------------------------------ Williams-Otto
class _Vec:
    """Tiny stand-in for the casadi DM the reference's callables return: supports x[i] and .toarray()."""

    def __init__(self, a):
        self.a = np.asarray(a, dtype=float).reshape(-1, 1)

    def __getitem__(self, i):
        return _Vec(self.a[i])

    def toarray(self):
        return self.a

    def __float__(self):
        return float(self.a.reshape(-1)[0])


def _newton(residual_and_jacobian, w0, u, tol=1e-13, max_iter=200):
    w = np.array(w0, dtype=float)
    for _ in range(max_iter):
        F, J = residual_and_jacobian(w, u)
        step = np.linalg.solve(J, F)
        w = w - step
        if np.max(np.abs(step)) <= tol * max(1.0, np.max(np.abs(w))):
            break
    return w


def _newton_batch(residual_and_jacobian_batch, w0, U, tol=1e-13, max_iter=200):
    """The same Newton iteration for k input points at once (U[k, 2]): identical arithmetic per point - elementwise
    residuals / Jacobians, one LAPACK gesv per point through the batched np.linalg.solve, and each point stops at its OWN
    convergence test (converged points are frozen) - so w[i] is bit-identical to _newton(.., U[i]).  This is what lets the
    lock-step replica driver evaluate the model at all SLSQP points of all replicas in one call."""
    U = np.asarray(U, dtype=float).reshape(-1, 2)
    W = np.tile(np.asarray(w0, dtype=float), (U.shape[0], 1))
    active = np.arange(U.shape[0])
    for _ in range(max_iter):
        if active.size == 0:
            break
        F, J = residual_and_jacobian_batch(W[active], U[active])
        step = np.linalg.solve(J, F[:, :, None])[:, :, 0]
        Wn = W[active] - step
        W[active] = Wn
        done = np.max(np.abs(step), axis=1) <= tol * np.maximum(1.0, np.max(np.abs(Wn), axis=1))
        active = active[~done]
    return W


class WO_system:
    """Plant: A + B -> C, B + C -> P + E, C + P -> G in a CSTR (6 mass fractions Xa, Xb, Xc, Xe, Xp, Xg)."""
    Fa = 1.8275
    Mt = 2105.2
    W0 = np.array([0.114805, 0.525604, 0.0260265, 0.207296, 0.0923376, 0.0339309])

    def __init__(self):
        self.eval = self.integrator_system()

    def _balances(self, w, u):
        Xa, Xb, Xc, Xe, Xp, Xg = w
        Fb, Tr = float(u[0]), float(u[1])
        Fa, Mt = self.Fa, self.Mt
        k1 = 1.6599e6 * np.exp(-6666.7 / (Tr + 273.15))
        k2 = 7.2117e8 * np.exp(-8333.3 / (Tr + 273.15))
        k3 = 2.6745e12 * np.exp(-11111. / (Tr + 273.15))
        Fr = Fa + Fb
        r1, r2, r3 = k1 * Xa * Xb * Mt, k2 * Xb * Xc * Mt, k3 * Xc * Xp * Mt
        F = np.array([Fa - r1 - Fr * Xa, Fb - r1 - r2 - Fr * Xb, 2 * r1 - 2 * r2 - r3 - Fr * Xc,
                      2 * r2 - Fr * Xe, r2 - 0.5 * r3 - Fr * Xp, 1.5 * r3 - Fr * Xg]) / Mt
        # d r / d w
        dr1 = np.array([k1 * Xb * Mt, k1 * Xa * Mt, 0, 0, 0, 0])
        dr2 = np.array([0, k2 * Xc * Mt, k2 * Xb * Mt, 0, 0, 0])
        dr3 = np.array([0, 0, k3 * Xp * Mt, 0, k3 * Xc * Mt, 0])
        J = np.stack([-dr1, -dr1 - dr2, 2 * dr1 - 2 * dr2 - dr3, 2 * dr2, dr2 - 0.5 * dr3, 1.5 * dr3])
        J = (J - Fr * np.eye(6)) / Mt
        return F, J

    def _balances_batch(self, W, U):
        """_balances for k points: W[k, 6], U[k, 2] -> F[k, 6], J[k, 6, 6]; same operations in the same order per point."""
        Xa, Xb, Xc, Xe, Xp, Xg = (W[:, i] for i in range(6))
        Fb, Tr = U[:, 0], U[:, 1]
        Fa, Mt = self.Fa, self.Mt
        k1 = 1.6599e6 * np.exp(-6666.7 / (Tr + 273.15))
        k2 = 7.2117e8 * np.exp(-8333.3 / (Tr + 273.15))
        k3 = 2.6745e12 * np.exp(-11111. / (Tr + 273.15))
        Fr = Fa + Fb
        r1, r2, r3 = k1 * Xa * Xb * Mt, k2 * Xb * Xc * Mt, k3 * Xc * Xp * Mt
        F = np.stack([Fa - r1 - Fr * Xa, Fb - r1 - r2 - Fr * Xb, 2 * r1 - 2 * r2 - r3 - Fr * Xc,
                      2 * r2 - Fr * Xe, r2 - 0.5 * r3 - Fr * Xp, 1.5 * r3 - Fr * Xg], axis=1) / Mt
        z = np.zeros_like(Xa)
        dr1 = np.stack([k1 * Xb * Mt, k1 * Xa * Mt, z, z, z, z], axis=1)
        dr2 = np.stack([z, k2 * Xc * Mt, k2 * Xb * Mt, z, z, z], axis=1)
        dr3 = np.stack([z, z, k3 * Xp * Mt, z, k3 * Xc * Mt, z], axis=1)
        J = np.stack([-dr1, -dr1 - dr2, 2 * dr1 - 2 * dr2 - dr3, 2 * dr2, dr2 - 0.5 * dr3, 1.5 * dr3], axis=1)
        J = (J - Fr[:, None, None] * np.eye(6)[None]) / Mt
        return F, J

    def states_batch(self, U):
        return _newton_batch(self._balances_batch, self.W0, U)

    def outputs_batch(self, U, Z=None):
        """[objective, constraint 1, constraint 2] of the plant at the k points U[k, 2]; Z[k, 3] = the standard-normal draws
        the three noisy callables would consume (None: noise-free).  Column j equals the scalar callable on U[i]."""
        U = np.asarray(U, dtype=float).reshape(-1, 2)
        X = self.states_batch(U)
        Fb = U[:, 0]
        Fr = self.Fa + Fb
        profit = 1043.38 * X[:, 4] * Fr + 20.92 * X[:, 3] * Fr - 79.23 * self.Fa - 118.34 * Fb
        if Z is None:
            return np.stack([-profit, -(X[:, 0] - 0.12), -(X[:, 5] - 0.08)], axis=1)
        return np.stack([-profit + 0.5 * Z[:, 0], -(X[:, 0] - 0.12 + 5e-4 * Z[:, 1]), -(X[:, 5] - 0.08 + 5e-4 * Z[:, 2])], axis=1)

    def integrator_system(self):
        return lambda w0, u: _Vec(_newton(self._balances, w0, u))

    def _profit(self, x, u):
        Fb = u[0]
        Fr = self.Fa + Fb
        return 1043.38 * float(x[4]) * Fr + 20.92 * float(x[3]) * Fr - 79.23 * self.Fa - 118.34 * Fb

    def WO_obj_sys_ca(self, u):
        return -self._profit(self.eval(self.W0, u), u) + 0.5 * np.random.normal(0., 1)

    def WO_obj_sys_ca_noise_less(self, u):
        return -self._profit(self.eval(self.W0, u), u)

    def WO_con1_sys_ca(self, u):
        x = self.eval(self.W0, u)
        return -(x[0].toarray()[0] - 0.12 + 5e-4 * np.random.normal(0., 1))

    def WO_con2_sys_ca(self, u):
        x = self.eval(self.W0, u)
        return -(x[5].toarray()[0] - 0.08 + 5e-4 * np.random.normal(0., 1))

    def WO_con1_sys_ca_noise_less(self, u):
        return -(self.eval(self.W0, u)[0].toarray()[0] - 0.12)

    def WO_con2_sys_ca_noise_less(self, u):
        return -(self.eval(self.W0, u)[5].toarray()[0] - 0.08)


class WO_model:
    """Two-reaction approximate model: A + 2B -> P + E, A + B + P -> G (5 mass fractions Xa, Xb, Xe, Xp, Xg)."""
    Fa = 1.8275
    Mt = 2105.2
    phi1, psi1, phi2, psi2 = -3., -17., -4., -29.
    Tref = 110. + 273.15
    W0 = np.array([0.114805, 0.525604, 0.207296, 0.0923376, 0.0339309])

    def __init__(self):
        self.eval = self.integrator_model()

    def _balances(self, w, u):
        Xa, Xb, Xe, Xp, Xg = w
        Fb, Tr = float(u[0]), float(u[1])
        Fa, Mt = self.Fa, self.Mt
        k1 = np.exp(self.phi1) * np.exp((self.Tref / (Tr + 273.15) - 1) * self.psi1)
        k2 = np.exp(self.phi2) * np.exp((self.Tref / (Tr + 273.15) - 1) * self.psi2)
        Fr = Fa + Fb
        r1, r2 = k1 * Xa * Xb * Xb * Mt, k2 * Xa * Xb * Xp * Mt
        F = np.array([Fa - r1 - r2 - Fr * Xa, Fb - 2 * r1 - r2 - Fr * Xb, 2 * r1 - Fr * Xe, r1 - r2 - Fr * Xp, 3 * r2 - Fr * Xg])
        dr1 = np.array([k1 * Xb * Xb * Mt, 2 * k1 * Xa * Xb * Mt, 0, 0, 0])
        dr2 = np.array([k2 * Xb * Xp * Mt, k2 * Xa * Xp * Mt, 0, k2 * Xa * Xb * Mt, 0])
        J = np.stack([-dr1 - dr2, -2 * dr1 - dr2, 2 * dr1, dr1 - dr2, 3 * dr2]) - Fr * np.eye(5)
        return F, J

    def _balances_batch(self, W, U):
        Xa, Xb, Xe, Xp, Xg = (W[:, i] for i in range(5))
        Fb, Tr = U[:, 0], U[:, 1]
        Fa, Mt = self.Fa, self.Mt
        k1 = np.exp(self.phi1) * np.exp((self.Tref / (Tr + 273.15) - 1) * self.psi1)
        k2 = np.exp(self.phi2) * np.exp((self.Tref / (Tr + 273.15) - 1) * self.psi2)
        Fr = Fa + Fb
        r1, r2 = k1 * Xa * Xb * Xb * Mt, k2 * Xa * Xb * Xp * Mt
        F = np.stack([Fa - r1 - r2 - Fr * Xa, Fb - 2 * r1 - r2 - Fr * Xb, 2 * r1 - Fr * Xe, r1 - r2 - Fr * Xp, 3 * r2 - Fr * Xg], axis=1)
        z = np.zeros_like(Xa)
        dr1 = np.stack([k1 * Xb * Xb * Mt, 2 * k1 * Xa * Xb * Mt, z, z, z], axis=1)
        dr2 = np.stack([k2 * Xb * Xp * Mt, k2 * Xa * Xp * Mt, z, k2 * Xa * Xb * Mt, z], axis=1)
        J = np.stack([-dr1 - dr2, -2 * dr1 - dr2, 2 * dr1, dr1 - dr2, 3 * dr2], axis=1) - Fr[:, None, None] * np.eye(5)[None]
        return F, J

    def outputs_batch(self, U):
        """[objective, constraint 1, constraint 2] of the model at the k points U[k, 2] (bit-identical to the scalar callables)."""
        U = np.asarray(U, dtype=float).reshape(-1, 2)
        X = _newton_batch(self._balances_batch, self.W0, U)
        Fb = U[:, 0]
        Fr = self.Fa + Fb
        obj = -(1043.38 * X[:, 3] * Fr + 20.92 * X[:, 2] * Fr - 79.23 * self.Fa - 118.34 * Fb)
        return np.stack([obj, -(X[:, 0] - 0.12), -(X[:, 4] - 0.08)], axis=1)

    def integrator_model(self):
        return lambda w0, u: _Vec(_newton(self._balances, w0, u))

    def WO_obj_ca(self, u):
        x = self.eval(self.W0, u)
        Fb = u[0]
        Fr = self.Fa + Fb
        return -(1043.38 * float(x[3]) * Fr + 20.92 * float(x[2]) * Fr - 79.23 * self.Fa - 118.34 * Fb)

    def WO_con1_model_ca(self, u):
        return -(self.eval(self.W0, u)[0].toarray()[0] - 0.12)

    def WO_con2_model_ca(self, u):
        return -(self.eval(self.W0, u)[4].toarray()[0] - 0.08)


# ------------------------------------------------------------------------------------------ batch forms of known callables
def batch_form(callables):
    """Vectorised twin of a list of model callables [obj, con_1, ..]: a function U[k, 2] -> values[k, len(callables)] whose
    column j is bit-identical to callables[j](U[i]) - or None if any callable is not one of this module's (arbitrary user
    callables are then evaluated one point at a time, as the reference does)."""
    def col(f):
        if f is Benoit_Model:
            return lambda U, cache: benoit_model_batch(U)[:, 0]
        if f is con1_model:
            return lambda U, cache: benoit_model_batch(U)[:, 1]
        if f is obj_empty or f is con_empty:
            return lambda U, cache: np.zeros(np.asarray(U).reshape(-1, 2).shape[0])
        owner = getattr(f, '__self__', None)
        if isinstance(owner, WO_model):
            j = {'WO_obj_ca': 0, 'WO_con1_model_ca': 1, 'WO_con2_model_ca': 2}.get(f.__name__)
            if j is not None:
                def g(U, cache, owner=owner, j=j):
                    key = id(owner)
                    if key not in cache:
                        cache[key] = owner.outputs_batch(U)          # one Newton solve serves objective and both constraints
                    return cache[key][:, j]
                return g
        return None
    cols = [col(f) for f in callables]
    if any(c is None for c in cols):
        return None

    def run(U):
        cache = {}
        return np.stack([c(U, cache) for c in cols], axis=1)
    return run
