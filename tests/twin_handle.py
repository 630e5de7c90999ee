"""ctypes front end of the ordered CPU twin (oracle/twin/gprto_twin.cpp).  TEST INFRASTRUCTURE ONLY.

TwinHandle offers the Handle methods the drop-in's RTO loop calls, computed by the twin - the host build of the kernels'
own specified arithmetic (gp_rto_ei_b200/csrc/gprto_det_math.h + the kernels' operation order).  On a GPU box the tests run
the same episode once on the B200 and once on the twin and require BIT-IDENTICAL trajectories."""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
TWIN_DIR = os.path.join(REPO, 'oracle', 'twin')
TWIN_SRC = os.path.join(TWIN_DIR, 'gprto_twin.cpp')
TWIN_LIB = os.path.join(TWIN_DIR, 'libgprto_twin.so')
_D = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    deps = [TWIN_SRC, os.path.join(REPO, 'gp_rto_ei_b200', 'csrc', 'gprto_det_math.h')]
    if force or not os.path.isfile(TWIN_LIB) or any(os.path.getmtime(p) > os.path.getmtime(TWIN_LIB) for p in deps):
        subprocess.run(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', TWIN_LIB, TWIN_SRC], check=True)
    return TWIN_LIB


def load():
    lib = ctypes.CDLL(build())
    for f in ('twin_exp', 'twin_log', 'twin_rsqrt', 'twin_erfc'):
        getattr(lib, f).restype, getattr(lib, f).argtypes = ctypes.c_double, [ctypes.c_double]
    lib.twin_nlml.argtypes = [ctypes.c_int, ctypes.c_int, _D, _D, _D, ctypes.c_double, _D, ctypes.POINTER(ctypes.c_int)]
    lib.twin_factor.argtypes = [ctypes.c_int, ctypes.c_int, _D, _D, _D, ctypes.c_double, _D, _D, _D, ctypes.POINTER(ctypes.c_int)]
    lib.twin_posterior.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _D, _D, _D, _D, _D, _D, ctypes.c_double, ctypes.c_double, _D, _D,
                                   ctypes.c_double, ctypes.c_int, _D, _D, _D]
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_D)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Twin:
    def __init__(self):
        self.lib = load()

    def nlml(self, Xn, Yn, theta, jitter=1e-8):
        Xn, Yn, theta = _c(Xn), _c(Yn), _c(theta)
        out, st = ctypes.c_double(), ctypes.c_int()
        self.lib.twin_nlml(Xn.shape[0], Xn.shape[1], _p(Xn), _p(Yn), _p(theta), jitter, ctypes.byref(out), ctypes.byref(st))
        return out.value, st.value

    def factor(self, Xn, Yn, theta, jitter=0.0):
        Xn, Yn, theta = _c(Xn), _c(Yn), _c(theta)
        n = Xn.shape[0]
        invK, alpha = np.empty((n, n)), np.empty(n)
        ld, st = ctypes.c_double(), ctypes.c_int()
        self.lib.twin_factor(n, Xn.shape[1], _p(Xn), _p(Yn), _p(theta), jitter, _p(invK), _p(alpha), ctypes.byref(ld), ctypes.byref(st))
        return invK, alpha, ld.value, st.value

    def posterior(self, Xn, invK, alpha, theta, x_mean, x_std, y_mean, y_std, cand, prior, f_best, acq):
        Xn, invK, alpha, theta, x_mean, x_std, cand = (_c(a) for a in (Xn, invK, alpha, theta, x_mean, x_std, cand))
        prior = None if prior is None else _c(prior)
        m = cand.shape[0]
        mean, var, score = np.empty(m), np.empty(m), np.empty(m)
        self.lib.twin_posterior(Xn.shape[0], Xn.shape[1], m, _p(Xn), _p(invK), _p(alpha), _p(theta), _p(x_mean), _p(x_std), float(y_mean),
                                float(y_std), _p(cand), _p(prior), 0.0 if f_best is None else float(f_best), int(acq), _p(mean), _p(var), _p(score))
        return mean, var, score


class TwinHandle:
    """Drop-in for api.Handle in the RTO loop (same call surface as tests/fake_handle.FakeHandle), twin arithmetic."""
    device = torch.device('cpu')
    launch_count = 0

    def __init__(self):
        self.t = Twin()

    def nlml_indexed_host(self, Xn, Yn, gp_index, theta, jitter=1e-8):
        Xn, Yn = Xn.numpy(), Yn.numpy()
        theta = np.asarray(theta, dtype=float)
        out = np.empty(len(gp_index)); st = np.zeros(len(gp_index), dtype=np.int32)
        for i, q in enumerate(gp_index):
            out[i], st[i] = self.t.nlml(Xn[q], Yn[q], theta[i], jitter)
        return out, st

    def factor(self, Xn, Yn, theta, jitter=0.0):
        B, n, d = Xn.shape
        invK = np.empty((B, n, n)); alpha = np.empty((B, n)); logdet = np.zeros(B); st = np.zeros(B, dtype=np.int32)
        for b in range(B):
            invK[b], alpha[b], logdet[b], st[b] = self.t.factor(Xn[b].numpy(), Yn[b].numpy(), theta[b].numpy(), jitter)
        return torch.from_numpy(invK), torch.from_numpy(alpha), torch.from_numpy(logdet), torch.from_numpy(st)

    def raise_on_failure(self, status, what=''):
        if int(status.abs().sum()):
            raise np.linalg.LinAlgError(what)

    def posterior_acq_host_np(self, gp, cand, prior=None, f_best=None, acq=3, want=('score',)):
        B, n, d = gp['Xn'].shape
        cand = np.asarray(cand, dtype=float)
        m = cand.shape[1]
        out = {k: np.empty((B, m)) for k in ('mean', 'var', 'score') if k in want}
        for b in range(B):
            mean, var, score = self.t.posterior(gp['Xn'][b].numpy(), gp['invK'][b].numpy(), gp['alpha'][b].numpy(), gp['theta'][b].numpy(),
                                                gp['x_mean'][b].numpy(), gp['x_std'][b].numpy(), float(gp['y_mean'][b]), float(gp['y_std'][b]),
                                                cand[b], None if prior is None else prior[b], None if f_best is None else float(f_best[b]), acq)
            for k, v in (('mean', mean), ('var', var), ('score', score)):
                if k in out:
                    out[k][b] = v
        return out
