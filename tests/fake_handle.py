"""CPU stand-in for api.Handle, built on the oracle.  TEST INFRASTRUCTURE ONLY (host-logic tests without a GPU).

Implements the handful of Handle methods the drop-in's RTO loop and the lock-step driver call, problem by problem through
oracle/gp_oracle.py, so that the HOST logic (SLSQP lock-step bookkeeping, RNG discipline, trust-region updates, batching /
padding / scatter of the lock-step replica driver) can be exercised in the `-m "not gpu"` suite.  Nothing in the product
imports this; the product's Handle raises without a B200."""
import numpy as np
import torch

from oracle import gp_oracle as O


class FakeHandle:
    device = torch.device('cpu')
    launch_count = 0

    def nlml_indexed_host(self, Xn, Yn, gp_index, theta, jitter=1e-8):
        Xn, Yn = Xn.numpy(), Yn.numpy()
        theta = np.asarray(theta, dtype=float)
        out = np.empty(len(gp_index)); st = np.zeros(len(gp_index), dtype=np.int32)
        for i, q in enumerate(gp_index):
            try:
                out[i] = O.nlml(theta[i], Xn[q], Yn[q])
            except np.linalg.LinAlgError:
                out[i], st[i] = np.nan, 1
        return out, st

    def factor(self, Xn, Yn, theta, jitter=0.0):
        B, n, d = Xn.shape
        invK = np.empty((B, n, n)); alpha = np.empty((B, n)); logdet = np.zeros(B); st = np.zeros(B, dtype=np.int32)
        for b in range(B):
            invK[b] = O.factor(theta[b].numpy(), Xn[b].numpy())
            alpha[b] = invK[b] @ Yn[b].numpy()
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
            g = dict(X_norm=gp['Xn'][b].numpy(), y_norm=None, theta=gp['theta'][b].numpy(), invK=gp['invK'][b].numpy(),
                     X_mean=gp['x_mean'][b].numpy(), X_std=gp['x_std'][b].numpy(), Y_mean=float(gp['y_mean'][b]), Y_std=float(gp['y_std'][b]))
            g['y_norm'] = np.linalg.solve(g['invK'], gp['alpha'][b].numpy())
            # candidate by candidate (results independent of the batch composition, like one GPU thread group per candidate)
            for c in range(m):
                mv, vv = O.inference(cand[b, c], g)
                if 'mean' in out:
                    out['mean'][b, c] = mv
                if 'var' in out:
                    out['var'][b, c] = vv
                if 'score' in out:
                    out['score'][b, c] = O.acquisition(mv, vv, 0.0 if prior is None else prior[b, c], None if f_best is None else float(f_best[b]), acq)
        return out
