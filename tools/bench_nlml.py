#!/usr/bin/env python
"""Secondary metric: batched NLML evaluations / s (BASELINE.json configs[3]: 16,384 GPs x 64 Sobol multistarts,
n=128, d=3).  python tools/bench_nlml.py [--B 16384] [--S 64] [--n 128] [--reps 5] [--path 0|1]   -> one JSON line."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from gp_rto_ei_b200 import api, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=16384)
    ap.add_argument('--S', type=int, default=64)
    ap.add_argument('--n', type=int, default=128)
    ap.add_argument('--d', type=int, default=3)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--path', type=int, default=0)
    a = ap.parse_args()
    h = api.Handle(0)
    h.set_option('k3_path', a.path)
    w = workloads.cfg4_torch(h.device, seed=1, B=a.B, n=a.n, d=a.d, S=a.S)
    nz = h.normalise(w['X'], w['y'])
    nll = torch.empty(a.B, a.S, dtype=torch.float64, device=h.device)
    st = torch.empty(a.B, a.S, dtype=torch.int32, device=h.device)
    for _ in range(a.warmup):
        h.nlml(nz['Xn'], nz['Yn'], w['theta'], out=nll, status=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        h.nlml(nz['Xn'], nz['Yn'], w['theta'], out=nll, status=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    bv, bi = h.argmin(nll)
    check = None
    if a.path not in (0, 1):            # cross-check against the shared-memory kernel on a slice
        h.set_option('k3_path', 1)
        nb = min(a.B, 64)
        ref, rst = h.nlml(nz['Xn'][:nb].contiguous(), nz['Yn'][:nb].contiguous(), w['theta'][:nb].contiguous())
        ok = (rst == 0) & (st[:nb] == 0)
        check = {'max_rel_diff_vs_generic': float(((nll[:nb] - ref).abs() / ref.abs().clamp(min=1.0))[ok].max()),
                 'status_mismatch': int(((rst == 0) != (st[:nb] == 0)).sum())}
    F = a.n ** 3 / 3 + 2 * a.n ** 2 + (a.n ** 2 / 2) * (3 * a.d + 2) + 4 * a.n
    evals = a.B * a.S
    print(json.dumps({'metric': 'NLML evals/s (FP64, batched)', 'value': evals / (ms * 1e-3), 'ms_per_sweep': ms, 'B': a.B, 'S': a.S, 'n': a.n,
                      'd': a.d, 'path': {0: 'auto', 1: 'generic', 2: 'reg', 3: 'lb (look-ahead + DMMA panel solve)', 4: 'la (round-1 look-ahead)', 5: 'lc (pivot chain in one warp)'}[a.path], 'flops_per_eval': F, 'achieved_tflops': evals * F / (ms * 1e-3) / 1e12,
                      'failed': int((st != 0).sum()), 'nan': int(torch.isnan(nll).sum()), 'argmin_hist_first8': bi[:8].tolist(), 'check': check}))


if __name__ == '__main__':
    main()
