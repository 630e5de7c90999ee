"""Batched device API over the C ABI (torch tensors are only the owners of device memory / streams).

Shapes follow include/gprto.h: B independent GPs, S hyper-parameter starts, n data points, d inputs,
m candidates per GP.  All tensors are float64, contiguous, on the handle's device.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import ACQ_EI, ACQ_LCB, ACQ_MEAN  # noqa: F401

JITTER_NLML = 1e-8     # reference: sub_uts/utilities_2.py:105


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Handle:
    """One per process and device.  Raises (never falls back) when the library or a B200 is missing."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError('gp_rto_ei_b200 needs a CUDA device (sm_100); there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        h = ctypes.c_void_p()
        _lib.check(self.lib.gprto_create(self.device.index, ctypes.byref(h)), 'gprto_create')
        self._h = h

    def close(self):
        if getattr(self, '_h', None):
            self.lib.gprto_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t, shape, name, dtype=torch.float64):
        if t.dtype != dtype or not t.is_contiguous() or t.device != self.device or tuple(t.shape) != tuple(shape):
            raise ValueError('%s must be a contiguous %s tensor of shape %s on %s (got %s %s on %s)' %
                             (name, dtype, tuple(shape), self.device, t.dtype, tuple(t.shape), t.device))
        return t

    def _new(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def _order_host_call(self):
        """The *_host entry points run on handle-private streams: fence them behind torch's current stream, which
        produced their device-resident inputs (include/gprto.h: gprto_order_after)."""
        _lib.check(self.lib.gprto_order_after(self._h, self._stream()), 'gprto_order_after')

    def set_option(self, name, value):
        _lib.check(self.lib.gprto_set_option(self._h, name.encode(), int(value)), 'gprto_set_option')

    @property
    def launch_count(self):
        return int(self.lib.gprto_launch_count(self._h))

    def limits(self):
        a = (ctypes.c_int64 * 4)()
        _lib.check(self.lib.gprto_limits(a), 'gprto_limits')
        return dict(k4_dmma_max_n=a[0], k3_reg_max_n=a[1], max_n=a[2], max_d=a[3])

    # -- entry points
    def normalise(self, X, Y):
        B, n, d = X.shape
        self._chk(X, (B, n, d), 'X'); self._chk(Y, (B, n), 'Y')
        out = dict(x_mean=self._new(B, d), x_std=self._new(B, d), y_mean=self._new(B), y_std=self._new(B),
                   Xn=self._new(B, n, d), Yn=self._new(B, n))
        _lib.check(self.lib.gprto_normalise(self._h, B, n, d, _ptr(X), _ptr(Y), _ptr(out['x_mean']), _ptr(out['x_std']),
                                            _ptr(out['y_mean']), _ptr(out['y_std']), _ptr(out['Xn']), _ptr(out['Yn']),
                                            self._stream()), 'gprto_normalise')
        return out

    def cov_se_ard(self, Xn, theta, add_noise=False, jitter=0.0):
        B, n, d = Xn.shape
        S = theta.shape[1]
        self._chk(Xn, (B, n, d), 'Xn'); self._chk(theta, (B, S, d + 2), 'theta')
        K = self._new(B, S, n, n)
        _lib.check(self.lib.gprto_cov_se_ard(self._h, B, S, n, d, _ptr(Xn), _ptr(theta), int(bool(add_noise)), float(jitter),
                                             _ptr(K), self._stream()), 'gprto_cov_se_ard')
        return K

    def nlml(self, Xn, Yn, theta, jitter=JITTER_NLML, out=None, status=None):
        B, n, d = Xn.shape
        S = theta.shape[1]
        self._chk(Xn, (B, n, d), 'Xn'); self._chk(Yn, (B, n), 'Yn'); self._chk(theta, (B, S, d + 2), 'theta')
        nll = self._new(B, S) if out is None else self._chk(out, (B, S), 'out')
        st = self._new(B, S, dtype=torch.int32) if status is None else self._chk(status, (B, S), 'status', torch.int32)
        _lib.check(self.lib.gprto_nlml(self._h, B, S, n, d, _ptr(Xn), _ptr(Yn), _ptr(theta), float(jitter), _ptr(nll), _ptr(st),
                                       self._stream()), 'gprto_nlml')
        return nll, st

    def nlml_indexed_host(self, Xn, Yn, gp_index, theta, jitter=JITTER_NLML):
        """Small synchronous batches driven from the host: Xn[G,n,d], Yn[G,n] on the device; gp_index[k] (int32) and
        theta[k,d+2] NumPy arrays -> (nll[k], status[k]) NumPy arrays.  No torch objects are created on this path."""
        G, n, d = Xn.shape
        gp_index = np.ascontiguousarray(gp_index, dtype=np.int32)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        k = gp_index.shape[0]
        nll = np.empty(k, dtype=np.float64)
        st = np.empty(k, dtype=np.int32)
        self._order_host_call()
        _lib.check(self.lib.gprto_nlml_indexed_host(self._h, G, k, n, d, _ptr(Xn), _ptr(Yn), gp_index.ctypes.data, theta.ctypes.data,
                                                    float(jitter), nll.ctypes.data, st.ctypes.data), 'gprto_nlml_indexed_host')
        return nll, st

    def posterior_acq_host_np(self, gp, cand, prior=None, f_best=None, acq=ACQ_EI, want=('score',)):
        """gprto_posterior_acq_host with NumPy arrays for the per-candidate streams: cand[B,m,d] (+ prior[B,m]) in,
        dict of NumPy arrays out (keys of `want` among mean, var, score, best)."""
        B, n, d = gp['Xn'].shape
        cand = np.ascontiguousarray(cand, dtype=np.float64)
        m = cand.shape[1]
        prior = None if prior is None else np.ascontiguousarray(prior, dtype=np.float64)
        out = {k: np.empty((B, m)) for k in ('mean', 'var', 'score') if k in want}
        if 'best' in want:
            out['best_score'], out['best_idx'] = np.empty(B), np.empty(B, dtype=np.int64)
        g = lambda k: out[k].ctypes.data if k in out else None  # noqa: E731
        self._order_host_call()
        _lib.check(self.lib.gprto_posterior_acq_host(
            self._h, B, n, d, m, _ptr(gp['Xn']), _ptr(gp['invK']), _ptr(gp['alpha']), _ptr(gp['theta']), _ptr(gp['x_mean']),
            _ptr(gp['x_std']), _ptr(gp['y_mean']), _ptr(gp['y_std']), cand.ctypes.data, None if prior is None else prior.ctypes.data,
            _ptr(f_best), int(acq), g('mean'), g('var'), g('score'), g('best_score'), g('best_idx')), 'gprto_posterior_acq_host')
        return out

    def argmin(self, val):
        B, S = val.shape
        self._chk(val, (B, S), 'val')
        bv, bi = self._new(B), self._new(B, dtype=torch.int64)
        _lib.check(self.lib.gprto_argmin(self._h, B, S, _ptr(val), _ptr(bv), _ptr(bi), self._stream()), 'gprto_argmin')
        return bv, bi

    def factor(self, Xn, Yn, theta_opt, jitter=0.0):
        B, n, d = Xn.shape
        self._chk(Xn, (B, n, d), 'Xn'); self._chk(Yn, (B, n), 'Yn'); self._chk(theta_opt, (B, d + 2), 'theta_opt')
        invK, alpha, logdet = self._new(B, n, n), self._new(B, n), self._new(B)
        st = self._new(B, dtype=torch.int32)
        _lib.check(self.lib.gprto_factor(self._h, B, n, d, _ptr(Xn), _ptr(Yn), _ptr(theta_opt), float(jitter), _ptr(invK),
                                         _ptr(alpha), _ptr(logdet), _ptr(st), self._stream()), 'gprto_factor')
        return invK, alpha, logdet, st

    def first_failure(self, status):
        first = ctypes.c_int64(-1)
        _lib.check(self.lib.gprto_first_failure(self._h, _ptr(status), status.numel(), self._stream(), ctypes.byref(first)),
                   'gprto_first_failure')
        return first.value

    def raise_on_failure(self, status, what='Cholesky'):
        """Mirror of the reference's uncaught numpy.linalg.LinAlgError (utilities_2.py:107)."""
        first = self.first_failure(status)
        if first >= 0:
            raise np.linalg.LinAlgError('Matrix is not positive definite (%s, problem %d of %d)' %
                                        (what, first, status.numel()))

    def posterior_acq(self, gp, cand, prior=None, f_best=None, acq=ACQ_EI, want=('mean', 'var', 'score', 'best')):
        """gp: dict with Xn[B,n,d], invK[B,n,n], alpha[B,n], theta[B,d+2], x_mean[B,d], x_std[B,d], y_mean[B], y_std[B]."""
        B, n, d = gp['Xn'].shape
        m = cand.shape[1]
        self._chk(gp['Xn'], (B, n, d), 'Xn'); self._chk(gp['invK'], (B, n, n), 'invK'); self._chk(gp['alpha'], (B, n), 'alpha')
        self._chk(gp['theta'], (B, d + 2), 'theta'); self._chk(gp['x_mean'], (B, d), 'x_mean'); self._chk(gp['x_std'], (B, d), 'x_std')
        self._chk(gp['y_mean'], (B,), 'y_mean'); self._chk(gp['y_std'], (B,), 'y_std'); self._chk(cand, (B, m, d), 'cand')
        if prior is not None:
            self._chk(prior, (B, m), 'prior')
        if f_best is not None:
            self._chk(f_best, (B,), 'f_best')
        out = {}
        for k in ('mean', 'var', 'score'):
            out[k] = self._new(B, m) if k in want else None
        if 'best' in want:
            out['best_score'], out['best_idx'] = self._new(B), self._new(B, dtype=torch.int64)
        else:
            out['best_score'] = out['best_idx'] = None
        _lib.check(self.lib.gprto_posterior_acq(
            self._h, B, n, d, m, _ptr(gp['Xn']), _ptr(gp['invK']), _ptr(gp['alpha']), _ptr(gp['theta']), _ptr(gp['x_mean']),
            _ptr(gp['x_std']), _ptr(gp['y_mean']), _ptr(gp['y_std']), _ptr(cand), _ptr(prior), _ptr(f_best), int(acq),
            _ptr(out['mean']), _ptr(out['var']), _ptr(out['score']), _ptr(out['best_score']), _ptr(out['best_idx']),
            self._stream()), 'gprto_posterior_acq')
        return out

    def posterior_acq_host(self, gp, cand_host, prior_host=None, f_best=None, acq=ACQ_EI, out_host=None):
        """End-to-end entry: candidates (and prior) in HOST tensors (pin them for full PCIe speed), results written
        to the HOST tensors given in out_host (keys among mean, var, score, best_score, best_idx)."""
        B, n, d = gp['Xn'].shape
        m = cand_host.shape[1]
        assert cand_host.device.type == 'cpu' and cand_host.dtype == torch.float64 and cand_host.is_contiguous()
        assert tuple(cand_host.shape) == (B, m, d)
        out_host = out_host or {}
        for k, v in out_host.items():
            assert v.device.type == 'cpu' and v.is_contiguous(), k
        g = lambda k: _ptr(out_host.get(k))  # noqa: E731
        self._order_host_call()
        _lib.check(self.lib.gprto_posterior_acq_host(
            self._h, B, n, d, m, _ptr(gp['Xn']), _ptr(gp['invK']), _ptr(gp['alpha']), _ptr(gp['theta']), _ptr(gp['x_mean']),
            _ptr(gp['x_std']), _ptr(gp['y_mean']), _ptr(gp['y_std']), _ptr(cand_host), _ptr(prior_host), _ptr(f_best), int(acq),
            g('mean'), g('var'), g('score'), g('best_score'), g('best_idx')), 'gprto_posterior_acq_host')
        return out_host

    def sample_candidates(self, centre, radius, m, shape=0, seed=0, idx=None):
        """centre[B,d], radius[B] (shape 2: radius[B,d+1] = semi-axes + ball radius) -> cand[B,m,d] (or [B,d] for the
        candidates idx[B])."""
        B, d = centre.shape
        self._chk(centre, (B, d), 'centre'); self._chk(radius, (B, d + 1) if shape == 2 else (B,), 'radius')
        out = self._new(B, d) if idx is not None else self._new(B, m, d)
        _lib.check(self.lib.gprto_sample_candidates(self._h, B, d, m, _ptr(centre), _ptr(radius), int(shape), int(seed), _ptr(idx),
                                                    _ptr(out), self._stream()), 'gprto_sample_candidates')
        return out

    def posterior_acq_sampled(self, gp, centre, radius, m, shape=0, seed=0, f_best=None, acq=ACQ_EI, out=None):
        """K4 with the m candidates of each GP generated inside the kernel; returns best_score[B], best_idx[B], best_x[B,d]."""
        B, n, d = gp['Xn'].shape
        self._chk(centre, (B, d), 'centre'); self._chk(radius, (B, d + 1) if shape == 2 else (B,), 'radius')
        out = out or dict(best_score=self._new(B), best_idx=self._new(B, dtype=torch.int64), best_x=self._new(B, d))
        _lib.check(self.lib.gprto_posterior_acq_sampled(
            self._h, B, n, d, int(m), _ptr(gp['Xn']), _ptr(gp['invK']), _ptr(gp['alpha']), _ptr(gp['theta']), _ptr(gp['x_mean']),
            _ptr(gp['x_std']), _ptr(gp['y_mean']), _ptr(gp['y_std']), _ptr(centre), _ptr(radius), int(shape), int(seed), _ptr(f_best),
            int(acq), _ptr(out['best_score']), _ptr(out['best_idx']), _ptr(out.get('best_x')), self._stream()),
            'gprto_posterior_acq_sampled')
        return out

    def pack_best(self, best_score, best_idx, idx_offset=0, out=None):
        count = best_score.numel()
        rec = self._new(count, 2) if out is None else out
        _lib.check(self.lib.gprto_pack_best(self._h, count, _ptr(best_score), _ptr(best_idx), int(idx_offset), _ptr(rec),
                                            self._stream()), 'gprto_pack_best')
        return rec

    def fit_fixed(self, X, Y, theta):
        """normalise + factor at given hyper-parameters -> the `gp` dict posterior_acq consumes."""
        nz = self.normalise(X, Y)
        invK, alpha, logdet, st = self.factor(nz['Xn'], nz['Yn'], theta)
        return dict(Xn=nz['Xn'], Yn=nz['Yn'], invK=invK, alpha=alpha, theta=theta, x_mean=nz['x_mean'], x_std=nz['x_std'],
                    y_mean=nz['y_mean'], y_std=nz['y_std'], logdet=logdet, status=st)


    def fit(self, X, Y, multi_hyper=15, noise=None, jitter=JITTER_NLML, ftol=1e-12, maxiter=10000):
        """Batched sibling of GP_model(X, Y, 'RBF', multi_hyper, noise=...) for B independent single-output GPs
        (reference: sub_uts/utilities_2.py:28-44, 120-177): X[B,n,d], Y[B,n] (device tensors or NumPy arrays).

        normalise (kernel) -> B x multi_hyper Sobol starts in the reference's log-hyper box (:131-158; a known noise
        variance `noise` (scalar or [B]) pins the last bound to 0.5 log(noise / y_std^2) +- 1e-3, :149-154) -> all
        B x S SLSQP instances in lock step over gprto_nlml_indexed_host (one launch per round; SciPy's SLSQP routine,
        tol 1e-12, forward differences, as :160-161) -> gprto_argmin over the starts (first minimum wins, :166) ->
        gprto_factor at the selected hyper-parameters (:168-175).
        Returns the gp dict posterior_acq consumes plus theta[B,d+2], localval[B,S] (NumPy), minindex[B] (NumPy)."""
        from . import slsqp_batch
        from .compat import sobol_seq
        T = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=self.device)  # noqa: E731
        X, Y = T(X).contiguous(), T(Y).contiguous()
        B, n, d = X.shape
        S = int(multi_hyper)
        nz = self.normalise(X, Y)
        lb = np.tile(np.array([-5.0] * (d + 1) + [-8.0]), (B, 1))
        ub = np.tile(np.array([5.0] * (d + 1) + [-4.0]), (B, 1))
        if noise is not None:
            centre = 0.5 * np.log(np.broadcast_to(np.asarray(noise, dtype=float), (B,)) / nz['y_std'].cpu().numpy() ** 2)
            lb[:, -1], ub[:, -1] = centre - 0.001, centre + 0.001
        sob = sobol_seq.i4_sobol_generate(d + 2, S)
        LB, UB = np.repeat(lb, S, axis=0), np.repeat(ub, S, axis=0)
        x0 = LB + (UB - LB) * np.tile(sob, (B, 1))                          # instance (b, j) = b * S + j
        owner = np.repeat(np.arange(B, dtype=np.int32), S)
        failed = np.zeros(B, dtype=bool)

        def fun_batch(inst, P):
            out, st = self.nlml_indexed_host(nz['Xn'], nz['Yn'], owner[inst], P, jitter=jitter)
            if st.any():            # the reference raises numpy.linalg.LinAlgError here (:107); reported per GP in `status`
                bad = np.flatnonzero(st)
                failed[owner[inst][bad]] = True
                out[bad] = 1e300
            return out

        res = slsqp_batch.minimize_arrays(fun_batch, x0, LB, UB, ftol=ftol, maxiter=maxiter)
        localval = res['fun'].reshape(B, S)
        _, best = self.argmin(T(localval))
        best_h = best.cpu().numpy()
        theta = T(res['x'].reshape(B, S, d + 2)[np.arange(B), best_h])
        invK, alpha, logdet, st = self.factor(nz['Xn'], nz['Yn'], theta)
        st = st | T(failed.astype(np.int32)).to(torch.int32)
        return dict(Xn=nz['Xn'], Yn=nz['Yn'], invK=invK, alpha=alpha, theta=theta, x_mean=nz['x_mean'], x_std=nz['x_std'],
                    y_mean=nz['y_mean'], y_std=nz['y_std'], logdet=logdet, status=st, localval=localval, localsol=res['x'].reshape(B, S, d + 2),
                    minindex=best_h, slsqp_calls=res['slsqp_calls'])


_default = {}


def default_handle(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError('gp_rto_ei_b200 needs a CUDA device (sm_100); there is no CPU fallback')
    idx = torch.cuda.current_device() if device is None else device
    if idx not in _default:
        _default[idx] = Handle(idx)
    return _default[idx]
