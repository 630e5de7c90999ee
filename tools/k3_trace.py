#!/usr/bin/env python
"""Timeline of ONE CTA of the K3 kernels (debug build with -DK3_TRACE): absolute clock64 stamps per warp / block step / event.

  GPRTO_LIB_OUT=build/libgprto_trace.so python -m gp_rto_ei_b200.build --force -DK3_TRACE
  GPRTO_LIB=build/libgprto_trace.so python tools/k3_trace.py --n 128 --path 0|3 [--out file.json]

Events, update warps (0..6): la: 0 step start, 1 panel rows solved, 2 after Z, 3 arrive X, 4 trailing update done, 5 after Y;
lb: 0 step start, 1 z / Linv fragments, 2 column kb+1 done (before Z), 3 after Z, 4 trailing update done, 5 after Y.
Panel warp (7): 0 loop top, 1 after X, 2 diagonal tile done, 3 after Y.  Step 31 = prologue (0 entry, 1 K built, 2 after Y(-1);
panel: 0 entry, 1 after X(-1), 2 diag(0) done, 3 after Y(-1))."""
import argparse
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from gp_rto_ei_b200 import api, workloads, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=128)
    ap.add_argument('--path', type=int, default=0)
    ap.add_argument('--B', type=int, default=1024)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    h = api.Handle(0)
    h.set_option('k3_path', a.path)
    w = workloads.cfg4_torch(h.device, seed=1, B=a.B, n=a.n, d=3, S=64)
    nz = h.normalise(w['X'], w['y'])
    for _ in range(2):
        nll, st = h.nlml(nz['Xn'], nz['Yn'], w['theta'])
    torch.cuda.synchronize()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    buf = np.zeros(8 * 32 * 8, dtype=np.int64)
    rc = lib.gprto_debug_k3_trace(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(buf.size))
    assert rc == 0, rc
    tr = buf.reshape(8, 32, 8)
    NT = (a.n + 15) // 16 * 2
    t0 = tr[tr > 0].min()
    rel = np.where(tr > 0, tr - t0, -1)
    out = {'n': a.n, 'path': a.path, 'NT': NT, 'total_cycles': int(rel.max()), 'prologue': rel[:, 31, :4].tolist(),
           'steps': rel[:, :NT, :8].tolist()}
    # summary per step: when each event happened, min/max over update warps
    rows = []
    for kb in range(NT):
        u = rel[:7, kb, :6]
        p = rel[7, kb, :4]
        rows.append({'kb': kb, 'upd_ev_min': u.min(axis=0).tolist(), 'upd_ev_max': u.max(axis=0).tolist(), 'panel': p.tolist()})
    out['summary'] = rows
    if a.out:
        with open(a.out, 'w') as fh:
            json.dump(out, fh)
    print('total', out['total_cycles'], 'prologue upd(w0)', rel[0, 31, :3].tolist(), 'panel', rel[7, 31, :4].tolist())
    for r in rows:
        print(r['kb'], 'upd max', r['upd_ev_max'], '| min', r['upd_ev_min'], '| panel', r['panel'])
    # compact view: per step, relative to the step start of the panel warp: hand-over (event 6 of any warp, lb only), panel done, last update warp done
    print('kb  step  handover  diag_done  upd_done(max ev4)')
    for kb in range(NT - 1):
        s0 = rel[7, kb, 0]
        nxt = rel[7, kb + 1, 0] if kb + 2 < NT else rel[:7, kb + 1, 0].max()
        ho = rel[:7, kb, 6].max()
        print('%2d %6d %8s %9d %9d' % (kb, nxt - s0, (ho - s0) if ho > 0 else '-', rel[7, kb, 2] - s0, rel[:7, kb, 4].max() - s0))


if __name__ == '__main__':
    main()
