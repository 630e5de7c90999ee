"""ctypes binding of libgprto.so (C ABI: include/gprto.h).  No fallback: a missing library is an error."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_uint64, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GPRTO_LIB', os.path.join(HERE, 'libgprto.so'))   # GPRTO_LIB: A/B builds of the same ABI

EINVAL, EUNSUPPORTED, ENODEVICE, ENOMEM = -10001, -10002, -10003, -10004
ACQ_MEAN, ACQ_LCB, ACQ_EI = 1, 2, 3
K4_AUTO, K4_GENERIC, K4_DMMA = 0, 1, 2

_P = c_void_p          # device / host pointers travel as integers
_lib = None

# name -> (restype, argtypes); kept in the order of include/gprto.h
PROTOTYPES = {
    'gprto_create': (c_int, [c_int, POINTER(c_void_p)]),
    'gprto_destroy': (c_int, [c_void_p]),
    'gprto_version': (c_char_p, []),
    'gprto_error_string': (c_char_p, [c_int]),
    'gprto_limits': (c_int, [POINTER(c_int64)]),
    'gprto_set_option': (c_int, [c_void_p, c_char_p, c_int64]),
    'gprto_launch_count': (c_int64, [c_void_p]),
    'gprto_order_after': (c_int, [c_void_p, _P]),
    'gprto_normalise': (c_int, [c_void_p, c_int64, c_int, c_int] + [_P] * 8 + [_P]),
    'gprto_cov_se_ard': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, _P, _P, c_int, c_double, _P, _P]),
    'gprto_nlml': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, _P, _P, _P, c_double, _P, _P, _P]),
    'gprto_nlml_indexed_host': (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P, c_double, _P, _P]),
    'gprto_argmin': (c_int, [c_void_p, c_int64, c_int64, _P, _P, _P, _P]),
    'gprto_factor': (c_int, [c_void_p, c_int64, c_int, c_int, _P, _P, _P, c_double, _P, _P, _P, _P, _P]),
    'gprto_posterior_acq': (c_int, [c_void_p, c_int64, c_int, c_int, c_int64] + [_P] * 11 + [c_int] + [_P] * 5 + [_P]),
    'gprto_posterior_acq_host': (c_int, [c_void_p, c_int64, c_int, c_int, c_int64] + [_P] * 11 + [c_int] + [_P] * 5),
    'gprto_sample_candidates': (c_int, [c_void_p, c_int64, c_int, c_int64, _P, _P, c_int, c_uint64, _P, _P, _P]),
    'gprto_posterior_acq_sampled': (c_int, [c_void_p, c_int64, c_int, c_int, c_int64] + [_P] * 10 + [c_int, c_uint64, _P, c_int, _P, _P, _P, _P]),
    'gprto_pack_best': (c_int, [c_void_p, c_int64, _P, _P, c_int64, _P, _P]),
    'gprto_first_failure': (c_int, [c_void_p, _P, c_int64, _P, POINTER(c_int64)]),
}


def load():
    """dlopen the in-tree library and attach prototypes.  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError('%s not found - build it with `python -m gp_rto_ei_b200.build` '
                           '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class GprtoError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = load().gprto_error_string(code).decode()
        super().__init__('%s failed: %s (%d)' % (where, msg, code))


def check(code, where):
    if code != 0:
        raise GprtoError(code, where)
