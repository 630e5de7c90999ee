"""Import the UNMODIFIED reference module in the build container.  TEST INFRASTRUCTURE ONLY.

/root/reference exists only in the build container (never on the GPU box), so this is used by
tests/golden/make_golden.py to mint the committed golden vectors and by the CPU tests that are
skipped when the reference is absent.  The three modules the reference imports but this image
lacks (casadi, matplotlib, sobol_seq) are provided from oracle/stubs.
"""
import importlib
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get('GPRTO_REFERENCE_ROOT', '/root/reference')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'sub_uts', 'utilities_2.py'))


def load(name='utilities_2'):
    """Return the reference's sub_uts.<name> module object (imported as `_ref_<name>`); name = utilities_2 | systems."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    modname = '_ref_' + name
    if modname in sys.modules:
        return sys.modules[modname]
    saved = list(sys.path)
    try:
        for dep in ('casadi', 'matplotlib', 'sobol_seq'):
            try:
                importlib.import_module(dep)
            except ImportError:
                if _STUBS not in sys.path:
                    sys.path.insert(0, _STUBS)
        if _REPO not in sys.path:
            sys.path.insert(0, _REPO)
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, 'sub_uts', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path[:] = saved
