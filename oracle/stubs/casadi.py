"""Inert stand-in for casadi so the reference module imports in this image (casadi is absent).

utilities_2.py:20 star-imports casadi and GP_model.__init__ (:44) builds a symbolic predictor that
neither main ever evaluates (GP_casadi=False, utilities_2.py:28).  Every object here swallows any
call / operator and returns another inert object.  `os` is re-exported because main_simple.py:21
reaches `os` through `from casadi import *`.  Test infrastructure only.
"""
import os  # noqa: F401  (re-exported on purpose)


class _Inert:
    shape = (0, 0)
    T = property(lambda self: self)

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Inert()

    def __getitem__(self, key):
        return _Inert()

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter(())

    def _op(self, *a, **k):
        return _Inert()

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _op
    __pow__ = __rpow__ = __neg__ = __pos__ = __matmul__ = __rmatmul__ = _op
    __array_priority__ = 1e6


class _Factory(_Inert):
    sym = staticmethod(lambda *a, **k: _Inert())
    zeros = staticmethod(lambda *a, **k: _Inert())
    ones = staticmethod(lambda *a, **k: _Inert())
    eye = staticmethod(lambda *a, **k: _Inert())


SX = MX = DM = _Factory
Function = rootfinder = nlpsol = _Inert


def _fn(*a, **k):
    return _Inert()


exp = log = sqrt = mtimes = sum1 = sum2 = vertcat = horzcat = fmax = fmin = jacobian = hessian = _fn

__all__ = ['SX', 'MX', 'DM', 'Function', 'rootfinder', 'nlpsol', 'exp', 'log', 'sqrt', 'mtimes',
           'sum1', 'sum2', 'vertcat', 'horzcat', 'fmax', 'fmin', 'jacobian', 'hessian', 'os']
