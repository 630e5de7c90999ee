"""Synthetic workloads of BASELINE.json / SURVEY.md par. 8(d), shared by bench.py, the tests and the
golden-vector generator.  NumPy (host, PCG64) for the parity-sized cases; torch (device, Philox) for
the full-size bench inputs, same distributions.

cfg 3  "synthetic batched GP inference": B independent GPs, n data points, d inputs, m EI
       candidates each inside a trust region of half-width `delta` around one data point.
cfg 4  "batched NLML hyper-parameter fit": B GPs x S Sobol multistarts in the reference's log-hyper box
       (reference: sub_uts/utilities_2.py:131-132, 158).
Hyper-parameters for cfg 3 are drawn so that cond(K) <~ 1e5 (SURVEY.md par. 7.3-2): W in [0.5, 4],
sf2 in [0.5, 2], sn2 in [1e-3, 1e-2]; theta = (0.5 log W_1..d, 0.5 log sf2, 0.5 log sn2).
"""
import math

import numpy as np

from .compat import sobol_seq

LB_THETA = (-5.0, -8.0)      # (lower bound of the first d+1 log-hypers, of the noise log-hyper)
UB_THETA = (5.0, -4.0)


def cfg3_numpy(seed=0, B=4, n=64, d=3, m=256, delta=0.25):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.uniform(-1.0, 1.0, (B, n, d))
    phase = rng.uniform(0.0, 2 * math.pi, (B, 1))
    y = np.sin(X.sum(-1) + phase) + 0.05 * rng.standard_normal((B, n))
    theta = np.empty((B, d + 2))
    theta[:, :d] = rng.uniform(-0.35, 0.7, (B, d))
    theta[:, d] = rng.uniform(-0.35, 0.35, B)
    theta[:, d + 1] = rng.uniform(0.5 * math.log(1e-3), 0.5 * math.log(1e-2), B)
    centre = X[np.arange(B), rng.integers(0, n, B)]                       # (B, d)
    cand = centre[:, None, :] + rng.uniform(-delta, delta, (B, m, d))
    f_best = y.min(axis=1)
    return dict(X=X, y=y, theta=theta, cand=cand, f_best=f_best)


def cfg3_torch(device, seed=0, B=65536, n=64, d=3, m=4096, delta=0.25):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device, generator=g)

    def unif(lo, hi, *shape):
        return torch.rand(*shape, **f64) * (hi - lo) + lo

    X = unif(-1.0, 1.0, B, n, d)
    phase = unif(0.0, 2 * math.pi, B, 1)
    y = torch.sin(X.sum(-1) + phase) + 0.05 * torch.randn(B, n, **f64)
    theta = torch.empty(B, d + 2, dtype=torch.float64, device=device)
    theta[:, :d] = unif(-0.35, 0.7, B, d)
    theta[:, d] = unif(-0.35, 0.35, B)
    theta[:, d + 1] = unif(0.5 * math.log(1e-3), 0.5 * math.log(1e-2), B)
    idx = torch.randint(0, n, (B,), device=device, generator=g)
    centre = X[torch.arange(B, device=device), idx]
    cand = torch.rand(B, m, d, **f64)
    cand.mul_(2 * delta).sub_(delta).add_(centre[:, None, :])
    f_best = y.min(dim=1).values
    return dict(X=X, y=y, theta=theta, cand=cand, f_best=f_best)


def sobol_theta_starts(d, S):
    """theta[S, d+2] = lb + (ub - lb) * sobol(d+2, S)  (utilities_2.py:131-139, 158)."""
    lb = np.array([LB_THETA[0]] * (d + 1) + [LB_THETA[1]])
    ub = np.array([UB_THETA[0]] * (d + 1) + [UB_THETA[1]])
    return lb + (ub - lb) * sobol_seq.i4_sobol_generate(d + 2, S)


def cfg4_numpy(seed=1, B=2, n=128, d=3, S=64):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.uniform(-1.0, 1.0, (B, n, d))
    phase = rng.uniform(0.0, 2 * math.pi, (B, 1))
    y = np.sin(X.sum(-1) + phase) + 0.05 * rng.standard_normal((B, n))
    theta = np.broadcast_to(sobol_theta_starts(d, S), (B, S, d + 2)).copy()
    return dict(X=X, y=y, theta=theta)


def cfg4_torch(device, seed=1, B=16384, n=128, d=3, S=64):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device, generator=g)
    X = torch.rand(B, n, d, **f64) * 2.0 - 1.0
    phase = torch.rand(B, 1, **f64) * (2 * math.pi)
    y = torch.sin(X.sum(-1) + phase) + 0.05 * torch.randn(B, n, **f64)
    th = torch.from_numpy(sobol_theta_starts(d, S)).to(device)
    theta = th.unsqueeze(0).expand(B, S, d + 2).contiguous()
    return dict(X=X, y=y, theta=theta)
