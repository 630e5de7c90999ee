#!/usr/bin/env python
"""Auxiliary kernels of the path: K1 (Gram matrix to HBM - the only HBM-bound kernel) and K2 (final factor: Cholesky,
explicit inverse, alpha).  python tools/bench_aux.py [--B 65536] [--n 64] [--d 3]  -> one JSON line."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from gp_rto_ei_b200 import api, workloads  # noqa: E402


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=65536)
    ap.add_argument('--n', type=int, default=64)
    ap.add_argument('--d', type=int, default=3)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--k2-path', type=int, default=0, help='0 auto (register-tile kernel for 32 < n <= 128), 1 shared-memory kernel')
    a = ap.parse_args()
    h = api.Handle(0)
    h.set_option('k2_path', a.k2_path)
    w = workloads.cfg3_torch(h.device, seed=0, B=a.B, n=a.n, d=a.d, m=8)
    nz = h.normalise(w['X'], w['y'])
    theta = w['theta']
    ms_f = timed(lambda: h.factor(nz['Xn'], nz['Yn'], theta), a.reps)
    ms_k = timed(lambda: h.cov_se_ard(nz['Xn'], theta.unsqueeze(1)), a.reps)
    ms_n = timed(lambda: h.normalise(w['X'], w['y']), a.reps)
    try:
        hbm = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        hbm = 6650.0
    kbytes = a.B * a.n * a.n * 8.0
    n = a.n
    f_flops = n ** 3 / 3 + n ** 3 / 3 + n ** 3 / 3 + (n * n / 2) * (3 * a.d + 2)      # Cholesky + triangular inverse + L^-T L^-1 + K
    print(json.dumps({'B': a.B, 'n': a.n, 'd': a.d, 'k2_path': a.k2_path,
                      'k2_factor': {'ms': ms_f, 'factors_per_s': a.B / (ms_f * 1e-3), 'tflops': a.B * f_flops / (ms_f * 1e-3) / 1e12,
                                    'bytes_out_gbs': (kbytes + a.B * n * 8) / (ms_f * 1e-3) / 1e9},
                      'k1_cov': {'ms': ms_k, 'matrices_per_s': a.B / (ms_k * 1e-3), 'achieved_gbs': kbytes / (ms_k * 1e-3) / 1e9, 'hbm_peak_gbs': hbm,
                                 'frac': kbytes / (ms_k * 1e-3) / 1e9 / hbm},
                      'normalise': {'ms': ms_n}}))


if __name__ == '__main__':
    main()
