"""CPU checks of the drop-in boundary: libgprto.so builds, loads, and exports exactly what include/gprto.h declares.
No compute call is made (there is no GPU on the CI box)."""
import ctypes
import os
import re
import subprocess

from conftest import REPO
from gp_rto_ei_b200 import _lib, build


def _header_functions():
    text = open(os.path.join(REPO, 'include', 'gprto.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gprto_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = build.build_lib()
    assert os.path.isfile(lib_path)
    declared = _header_functions()
    assert len(declared) >= 15
    out = subprocess.run(['nm', '-D', '--defined-only', lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r' T (gprto_[a-z0-9_]+)', out))
    assert exported == set(declared), (sorted(exported ^ set(declared)))
    assert set(_lib.PROTOTYPES) == set(declared)


def test_binding_loads_and_reports_no_device_without_fallback():
    lib = _lib.load()
    assert b'sm_100a' in lib.gprto_version()
    h = ctypes.c_void_p()
    import torch
    if not torch.cuda.is_available():
        assert lib.gprto_create(0, ctypes.byref(h)) == _lib.ENODEVICE
        assert lib.gprto_error_string(_lib.ENODEVICE) == b'no sm_100 CUDA device'
    lim = (ctypes.c_int64 * 4)()
    assert lib.gprto_limits(lim) == 0 and lim[0] == 128 and lim[3] >= 8


def test_sass_is_sm100_and_uses_fp64_tensor_path():
    sass = subprocess.run(['cuobjdump', '-sass', build.LIB], capture_output=True, text=True, check=True).stdout
    assert 'sm_100a' in sass
    assert 'DMMA.8x8x4' in sass          # FP64 tensor-core path of K4
    assert len(set(re.findall(r'arch = (sm_\w+)', sass))) == 1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, 'gp_rto_ei_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(root, f)
