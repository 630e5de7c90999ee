"""Lock-step Monte-Carlo RTO: R seeded replicas of one trust-region GP-RTO episode advanced TOGETHER, so that every
SLSQP round of every replica's hyper-parameter fit / acquisition optimisation is one batched kernel launch
(BASELINE.json configs[4]: "Monte-Carlo noisy Williams-Otto RTO: 8,192 replicas sharded over 8xB200"; SURVEY.md par. 8f-1/-2).

The reference runs its episodes one after the other, one GP evaluation per Python call (run_simple/main_simple.py:174,
run_WO/main_WO_prior_EI.py:105; loop body sub_uts/utilities_2.py:888-1038).  Here the episode is a state machine over
struct-of-arrays replica state:

  per RTO iteration    K3  NLML of R x (1+ng) GPs x 15 Sobol starts x (1 + d+2) forward-difference points   (utilities_2.py:156-163)
                       K2  final factor of the R x (1+ng) fitted GPs                                        (:168-175)
                       K4  objective / constraint GPs at the SLSQP points of R x 30 acquisition starts      (:970-984)
                       K4  posterior mean at (xk, xnew) for the trust-region ratio                          (:648-650)

and the host keeps only what is inherently sequential: SciPy's SLSQP C routine per instance (gp_rto_ei_b200/slsqp_batch.py)
and each replica's two random streams.  Replica r is a function of its seed alone: it owns numpy.random.RandomState(r) and
random.Random(r) and consumes them in exactly the order the sequential drop-in (and the reference) does - plant noise is
drawn one normal per noisy plant call, in call order (utilities_2.py:646-647, 876-883, 932, 999-1001), the start points by
Ball_sampling / Ellipse_sampling (:678-730).  All arithmetic per replica is the sequential drop-in's (same kernels per
problem, same NumPy operations per point), so a lock-step replica is BIT-IDENTICAL to
gp_rto_ei_b200.replicas.run_replica(seed) (tests/test_gpu_lockstep.py).
"""
import random as _pyrandom

import numpy as np
import torch

from . import api, slsqp_batch
from .compat import sobol_seq
from .sub_uts import systems as S
from .sub_uts import utilities_2 as u2


class Problem:
    """Batch definition of a benchmark problem: model(U) -> [k, 1+ng] (objective model, constraint models),
    plant(U, Z) -> [k, 1+ng] (Z[k, 1+ng] standard normals of the noisy plant callables, None = noise-free)."""

    def __init__(self, name):
        self.name = name
        if name == 'simple':            # run_simple/main_simple.py:179-206 (EI / prior / known noise block)
            self.x0 = np.array([1.1, -0.1])
            self.Xtrain = np.array([[1.2, 0.], [1.4, 0.1], [1.3, -0.1]])
            self.bounds = np.array([[-.6, 1.5], [-1., 1.]])
            self.noise = [1e-3] * 2
            self.scale_inputs = False
            self.ng = 1
            self.model = S.benoit_model_batch
            self.plant = S.benoit_plant_batch
        elif name == 'wo':              # run_WO/main_WO_prior_EI.py (EI, model prior, known noise)
            self._model, self._plant = S.WO_model(), S.WO_system()
            self.x0 = np.array([6.9, 83.])
            self.Xtrain = np.array([[5.7, 74.], [6.35, 74.9], [6.6, 75.], [6.75, 79.]])
            self.bounds = np.array([[4., 7.], [70., 100.]])
            self.noise = [0.5 ** 2, 5e-8, 5e-8]
            self.scale_inputs = True
            self.ng = 2
            self.model = self._model.outputs_batch
            self.plant = self._plant.outputs_batch
        else:
            raise ValueError(name)
        self.d = self.x0.size


class LockstepRTO:
    """R replicas of ITR_GP_RTO.RTO_routine (TR_scaling = TR_curvature = inner_TR = False) in lock step."""

    def __init__(self, problem, seeds, obj_setting=3, n_iter=20, multi_opt=30, multi_hyper=15, Delta0=0.25, Delta_max=0.7,
                 eta0=0.2, eta1=0.8, gamma_red=0.8, gamma_incr=1.2, handle=None):
        self.p = problem if isinstance(problem, Problem) else Problem(problem)
        self.seeds = [int(s) for s in seeds]
        self.R = len(self.seeds)
        self.obj, self.n_iter, self.multi_opt, self.multi_hyper = obj_setting, n_iter, multi_opt, multi_hyper
        self.Delta_max, self.eta0, self.eta1, self.gamma_red, self.gamma_incr = Delta_max, eta0, eta1, gamma_red, gamma_incr
        self.Delta0 = np.full(self.R, float(Delta0))
        self.h = handle or api.default_handle()
        self.rs = [np.random.RandomState(s) for s in self.seeds]       # == np.random.seed(s) + the module-level functions
        self.pr = [_pyrandom.Random(s) for s in self.seeds]            # == random.seed(s) + random.random()
        p = self.p
        self.ng, self.d = p.ng, p.d
        self.TRmat = np.linalg.inv(np.diag(p.bounds[:, 1] - p.bounds[:, 0])) if p.scale_inputs else np.eye(p.d)
        self.dead = np.zeros(self.R, dtype=bool)                       # replicas whose episode the reference would have aborted (LinAlgError)
        self.stats = dict(slsqp_calls_fit=0, slsqp_calls_acq=0, rounds_fit=0, rounds_acq=0, nlml_evals=0, k4_points=0)

    # ---------------------------------------------------------------- noisy plant, one normal per call, in call order
    def _draw(self, count):
        """Z[R, count]: the next `count` standard normals of every replica's NumPy stream."""
        return np.stack([rs.standard_normal(count) for rs in self.rs])

    def _plant_all_outputs(self, U, reps):
        """Noisy [obj, cons..] at U[k] for the replicas reps[k]: 1+ng draws per point, in output order (mismatch rows)."""
        Z = np.stack([self.rs[r].standard_normal(self.ng + 1) for r in reps])
        return self.p.plant(U, Z)

    # ---------------------------------------------------------------- GP fit of all replicas (utilities_2.py:795-815 -> 120-177)
    def _fit(self, X, Y):
        """X[R, n, d], Y[R, 1+ng, n] -> device dict of the R x (1+ng) fitted GPs (replica-major, objective first)."""
        R, n, d = X.shape
        G = self.ng + 1
        S_ = self.multi_hyper
        h = self.h
        Xn = np.empty((R, G, n, d)); Yn = np.empty((R, G, n))
        xm = np.empty((R, G, d)); xs = np.empty((R, G, d)); ym = np.empty((R, G)); ys = np.empty((R, G))
        LB = np.empty((R, G, d + 2)); UB = np.empty((R, G, d + 2))
        known = self.p.noise is not None and self.p.noise[0] is not None
        for r in range(R):
            # the very NumPy calls of GP_model.__init__ / _fit_shared_X, per replica (bit-identical public attributes)
            Xr = X[r]
            mu, sd = np.mean(Xr, axis=0), np.std(Xr, axis=0)
            Xnr = (Xr - mu) / sd
            for g in range(G):
                Yg = Y[r, g, :].reshape(n, 1)
                y_mu, y_sd = np.mean(Yg, axis=0), np.std(Yg, axis=0)
                Xn[r, g], Yn[r, g] = Xnr, ((Yg - y_mu) / y_sd)[:, 0]
                xm[r, g], xs[r, g], ym[r, g], ys[r, g] = mu, sd, y_mu[0], y_sd[0]
                nz = self.p.noise[g] if known else None
                LB[r, g], UB[r, g] = u2._hyper_bounds(d, nz, y_sd[0] if nz is not None else None)
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=h.device)  # noqa: E731
        Xn_d, Yn_d = dev(Xn.reshape(R * G, n, d)), dev(Yn.reshape(R * G, n))
        sob = sobol_seq.i4_sobol_generate(d + 2, S_)
        lb = np.repeat(LB.reshape(R * G, 1, d + 2), S_, axis=1).reshape(-1, d + 2)
        ub = np.repeat(UB.reshape(R * G, 1, d + 2), S_, axis=1).reshape(-1, d + 2)
        x0 = lb + (ub - lb) * np.tile(sob, (R * G, 1))                      # instance (dataset q, start j) = q * S + j
        owner = np.repeat(np.arange(R * G, dtype=np.int32), S_)
        failed = np.zeros(R * G, dtype=bool)

        def fun_batch(inst, P):
            out, st = h.nlml_indexed_host(Xn_d, Yn_d, owner[inst], P, jitter=u2._JITTER)
            self.stats['nlml_evals'] += len(inst)
            self.stats['rounds_fit'] += 1
            if st.any():            # the reference raises LinAlgError here and the episode dies (utilities_2.py:107): mark the replica
                bad = np.flatnonzero(st)
                failed[owner[inst][bad]] = True
                out[bad] = 1e300
            return out

        res = slsqp_batch.minimize_arrays(fun_batch, x0, lb, ub, ftol=1e-12, maxiter=10000)
        self.stats['slsqp_calls_fit'] += res['slsqp_calls']
        vals = res['fun'].reshape(R * G, S_)
        best = np.argmin(vals, axis=1)                                       # first minimum wins (utilities_2.py:166)
        theta = res['x'].reshape(R * G, S_, d + 2)[np.arange(R * G), best]
        theta_d = dev(theta)
        invK, alpha, logdet, st = h.factor(Xn_d, Yn_d, theta_d, jitter=0.0)
        failed |= (st.cpu().numpy() != 0)
        self.dead |= failed.reshape(R, G).any(axis=1)
        gp = dict(Xn=Xn_d, invK=invK, alpha=alpha, theta=theta_d, x_mean=dev(xm.reshape(R * G, d)), x_std=dev(xs.reshape(R * G, d)),
                  y_mean=dev(ym.reshape(R * G)), y_std=dev(ys.reshape(R * G)))
        io = torch.arange(0, R * G, G, device=h.device)
        self.gp_obj = {k: v[io].contiguous() for k, v in gp.items()}
        if self.ng:
            ic = torch.as_tensor(np.setdiff1d(np.arange(R * G), np.arange(0, R * G, G)), device=h.device)
            self.gp_con = {k: v[ic].contiguous() for k, v in gp.items()}      # index r * ng + (g - 1)
        self.theta = theta.reshape(R, G, d + 2)

    # ---------------------------------------------------------------- acquisition multistart of all replicas (:964-992)
    def _acquire(self, xk, d_inits, obj_min):
        """xk[R, d]; d_inits[R, multi_opt, d] -> (x[R*multi_opt, d], fun, success) of the lock-step SLSQP."""
        R, d, ng, mo, h = self.R, self.d, self.ng, self.multi_opt, self.h
        bounds = self.p.bounds
        lb = np.repeat((bounds[:, 0][None, :] - xk)[:, None, :], mo, axis=1).reshape(-1, d)
        ub = np.repeat((bounds[:, 1][None, :] - xk)[:, None, :], mo, axis=1).reshape(-1, d)
        fb = torch.as_tensor(np.ascontiguousarray(obj_min if self.obj == 3 else np.zeros(R)), device=h.device)
        cache = {}
        dead_inst = np.repeat(self.dead, mo)

        def evaluate(inst, P):
            rep = inst // mo
            Xq = P + xk[rep]
            M = self.p.model(Xq)                                             # [k, 1+ng] priors
            # group the points by replica, padded to the longest group (a GP evaluates its own candidates only)
            order = np.argsort(rep, kind='stable')
            cnt = np.bincount(rep, minlength=R)
            m_max = int(cnt.max())
            start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
            pos = np.empty(len(rep), dtype=np.int64)
            pos[order] = np.arange(len(rep)) - start[rep[order]]
            cand = np.repeat(xk[:, None, :], m_max, axis=1)
            cand[rep, pos] = Xq
            prior = np.zeros((R, m_max))
            prior[rep, pos] = M[:, 0]
            sc = h.posterior_acq_host_np(self.gp_obj, cand, prior, fb, self.obj, want=('score',))['score'][rep, pos]
            cons = np.empty((len(rep), ng + 1))
            if ng:
                cprior = np.zeros((R, ng, m_max))
                cprior[rep, :, pos] = M[:, 1:]
                cm = h.posterior_acq_host_np(self.gp_con, np.repeat(cand[:, None], ng, axis=1).reshape(R * ng, m_max, d),
                                             cprior.reshape(R * ng, m_max), None, api.ACQ_MEAN, want=('score',))['score']
                cons[:, :ng] = cm.reshape(R, ng, m_max)[rep, :, pos]
            cons[:, ng] = S._sq(self.Delta0[rep]) - np.einsum('ij,ij->i', P, P)     # pow(), as the scalar Delta0 ** 2 of TR_con
            self.stats['k4_points'] += len(rep) * (ng + 1)
            self.stats['rounds_acq'] += 1
            dd = dead_inst[inst]
            if dd.any():
                sc = np.where(dd, 0.0, sc)
                cons = np.where(dd[:, None], 1.0, cons)
            return sc, cons

        def fun_batch(inst, P):
            f, c = evaluate(inst, P)
            cache['c'] = c
            return f

        def con_batch(inst, P):
            return cache.pop('c')

        res = slsqp_batch.minimize_arrays(fun_batch, d_inits.reshape(-1, d), lb, ub, con_batch=con_batch, m_ineq=ng + 1, ftol=1e-12,
                                          maxiter=10000)
        self.stats['slsqp_calls_acq'] += res['slsqp_calls']
        return res

    # ---------------------------------------------------------------- the episode (utilities_2.py:888-1038)
    def run(self, progress=None):
        p, R, d, ng, G = self.p, self.R, self.d, self.ng, self.ng + 1
        nd = p.Xtrain.shape[0]
        reps = np.arange(R)
        # compute_data (:452-471): plant-minus-model mismatch at the initial design; draws output-major, point-minor
        Z0 = np.stack([rs.standard_normal(G * nd).reshape(G, nd) for rs in self.rs])          # [R, G, nd]
        U0 = np.tile(p.Xtrain, (R, 1))
        Y0 = p.plant(U0, Z0.transpose(0, 2, 1).reshape(R * nd, G)) - p.model(U0)              # [R*nd, G]
        ytrain = Y0.reshape(R, nd, G).transpose(0, 2, 1)                                    # [R, G, nd]
        # RTO_routine: evaluate the initial point
        xnew = np.tile(p.x0, (R, 1))
        ynew = self._plant_all_outputs(xnew, reps) - p.model(xnew)                          # [R, G]
        X_opt = np.concatenate([np.tile(p.Xtrain[None], (R, 1, 1)), xnew[:, None, :]], axis=1)       # [R, n, d]
        y_opt = np.concatenate([ytrain, ynew[:, :, None]], axis=2)                          # [R, G, n]
        self._fit(X_opt, y_opt)
        obj_min = self._min_so_far(X_opt)
        TR_l = np.empty((R, self.n_iter + 1)); TR_l[:, 0] = self.Delta0
        backtrack = np.zeros((R, self.n_iter), dtype=bool)
        for it in range(self.n_iter):
            xk = xnew.copy()
            if p.scale_inputs:
                d_inits = np.stack([np.stack([u2.ellipse_sample(self.rs[r], d, self.TRmat, self.Delta0[r]) for _ in range(self.multi_opt)])
                                    for r in range(R)])
            else:
                d_inits = np.stack([np.stack([u2.ball_sample(self.rs[r], self.pr[r], d, self.Delta0[r]) for _ in range(self.multi_opt)])
                                    for r in range(R)])
            res = self._acquire(xk, d_inits, obj_min)
            val = np.where(res['status'] == 0, res['fun'], np.inf).reshape(R, self.multi_opt)
            sol = res['x'].reshape(R, self.multi_opt, d)
            for r in range(R):
                if np.min(val[r]) == np.inf:                                 # 'warming, no feasible solution found' (:985-988)
                    xnew[r] = xk[r] * (1 + 0.01 * self.rs[r].randn())
                else:
                    xnew[r] = sol[r, int(np.argmin(val[r]))] + xk[r]
            ynew = self._plant_all_outputs(xnew, reps) - p.model(xnew)
            self._draw(1)                                                    # print('New Objective: ', -obj_system(xnew)) consumes one draw (:1001)
            X_opt = np.concatenate([X_opt, xnew[:, None, :]], axis=1)
            y_opt = np.concatenate([y_opt, ynew[:, :, None]], axis=2)
            obj_min = self._min_so_far(X_opt)
            xnew = self._step_constraint(xk, xnew, backtrack, it)
            self._fit(X_opt, y_opt)
            TR_l[:, it + 1] = self.Delta0
            if progress:
                progress(it)
        return dict(X_opt=X_opt, y_opt=y_opt, TR_l=TR_l, xnew=xnew, backtrack=backtrack, dead=self.dead.copy())

    def _min_so_far(self, X_opt):
        """find_min_so_far (:876-883): fresh noisy objective at every stored point, n draws per replica."""
        R, n, d = X_opt.shape
        Z = np.zeros((R * n, self.ng + 1))
        Z[:, 0] = self._draw(n).reshape(-1)
        return self.p.plant(X_opt.reshape(R * n, d), Z)[:, 0].reshape(R, n).min(axis=1)

    def _step_constraint(self, xk, xnew, backtrack, it):
        """Step_constraint + Adjust_TR (:757-777, 637-672)."""
        R, ng, h = self.R, self.ng, self.h
        # System_Cons_check: every constraint callable is evaluated (ng draws)
        Z = np.zeros((R, ng + 1))
        if ng:
            Z[:, 1:] = self._draw(ng)
        cons = self.p.plant(xnew, Z)[:, 1:]
        ok = np.all(cons > 0, axis=1)
        # Adjust_TR for the replicas that passed: plant at xk and xnew (2 draws, in this order), GP mean at both
        Zk = np.zeros((R, ng + 1)); Zn = np.zeros((R, ng + 1))
        for r in np.flatnonzero(ok):
            z = self.rs[r].standard_normal(2)
            Zk[r, 0], Zn[r, 0] = z[0], z[1]
        plant_i = self.p.plant(xk, Zk)[:, 0]
        plant_ip = self.p.plant(xnew, Zn)[:, 0]
        both = np.stack([xk, xnew], axis=1)                                  # [R, 2, d]
        m_pred = h.posterior_acq_host_np(self.gp_obj, both, None, None, api.ACQ_MEAN, want=('mean',))['mean']      # [R, 2]
        mod = self.p.model(both.reshape(2 * R, self.d))[:, 0].reshape(R, 2)
        with np.errstate(divide='ignore', invalid='ignore'):
            rho = (plant_i - plant_ip) / ((mod[:, 0] + m_pred[:, 0]) - (mod[:, 1] + m_pred[:, 1]))
        out = xnew.copy()
        for r in range(R):
            if not ok[r]:
                backtrack[r, it] = True
                self.Delta0[r] = self.Delta0[r] * self.gamma_red
                out[r] = xk[r]
            elif plant_ip[r] < plant_i[r]:
                if rho[r] >= self.eta1:
                    self.Delta0[r] = min(self.Delta0[r] * self.gamma_incr, self.Delta_max)
                elif rho[r] < self.eta0:
                    self.Delta0[r] = self.Delta0[r] * self.gamma_red
            else:
                out[r] = xk[r]
                self.Delta0[r] = self.Delta0[r] * self.gamma_red
                backtrack[r, it] = True
        return out

    # ---------------------------------------------------------------- outcome
    def outcome(self, result):
        """Noise-free objective and feasibility of every replica's final iterate."""
        Y = self.p.plant(result['xnew'], None)
        return Y[:, 0], np.all(Y[:, 1:] > 0, axis=1) & ~result['dead']
