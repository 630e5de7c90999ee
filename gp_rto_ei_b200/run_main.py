"""Run one of the reference's driver scripts UNCHANGED on top of this package.

    python -m gp_rto_ei_b200.run_main /path/to/GP-RTO-EI/run_simple/main_simple.py [--episodes N] [--iters K] [--out DIR]

The script's ``from sub_uts.utilities_2 import *`` / ``from sub_uts.systems import *`` resolve to
gp_rto_ei_b200/sub_uts (this package's directory is put first on sys.path); casadi, matplotlib, sobol_seq and the
plotting module are replaced by the stand-ins in gp_rto_ei_b200/compat only if they are not importable.
--episodes / --iters shorten the Monte-Carlo loops without editing the script (its ``range(30)`` episode loops and its
``n_iter = 20`` are intercepted through the globals the script is executed with); results are pickled by the script
itself into the working directory (--out, default: a temporary directory).
"""
import argparse
import builtins
import importlib
import os
import runpy
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def prepare_path():
    for p in (REPO, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    shims = os.path.join(HERE, 'compat', 'shims')
    need = False
    for mod in ('casadi', 'matplotlib', 'plots_RTO'):
        try:
            importlib.import_module(mod)
        except Exception:
            need = True
    if need and shims not in sys.path:
        sys.path.append(shims)
    try:
        importlib.import_module('sobol_seq')
    except ImportError:
        from gp_rto_ei_b200.compat import sobol_seq
        sys.modules['sobol_seq'] = sobol_seq


def run(script, episodes=None, iters=None, out=None):
    prepare_path()
    out = out or tempfile.mkdtemp(prefix='gprto_run_')
    os.makedirs(out, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(out)
    init = {}
    if episodes is not None:
        def short_range(*a):
            if len(a) == 1 and a[0] == 30:
                return builtins.range(episodes)
            return builtins.range(*a)
        init['range'] = short_range
    if iters is not None:
        import gp_rto_ei_b200.sub_uts.utilities_2 as u2   # noqa: F401
        import sub_uts.utilities_2 as su2
        orig = su2.ITR_GP_RTO.__init__

        def patched(self, *a, **k):
            a = list(a)
            if len(a) > 11:
                a[11] = iters
            elif 'n_iter' in k:
                k['n_iter'] = iters
            orig(self, *a, **k)
        su2.ITR_GP_RTO.__init__ = patched
    try:
        return runpy.run_path(script, init_globals=init, run_name='__main__'), out
    finally:
        os.chdir(cwd)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('script')
    ap.add_argument('--episodes', type=int, default=None)
    ap.add_argument('--iters', type=int, default=None)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    _, where = run(a.script, a.episodes, a.iters, a.out)
    print('results written to', where)
