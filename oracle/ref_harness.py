"""Import the UNMODIFIED reference module.  TEST INFRASTRUCTURE ONLY.

/root/reference exists only in the build container.  `stage()` (called by __graft_entry__.build()) copies the
reference's Python sources, its two driver scripts and its result pickles, byte for byte, into the git-ignored
`baseline/_ref/` - that directory is NOT gpurun-ignored, so it travels to the GPU box, where the `-m gpu` tests run
the reference's own mains from it and bench.py's `cpu_baseline` / `--impl reference` legs time the reference's own
functions (kind "reference").  Nothing under baseline/_ref is ever committed (no reference sources in the history).
The three modules the reference imports but this image lacks (casadi, matplotlib, sobol_seq) are provided from
oracle/stubs.
"""
import importlib
import importlib.util
import os
import sys

import shutil

_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE_ROOT = '/root/reference'                                  # build container only
STAGED_ROOT = os.path.join(_REPO, 'baseline', '_ref')            # git-ignored, ships with the tree


def _has_ref(root):
    return os.path.isfile(os.path.join(root, 'sub_uts', 'utilities_2.py'))


def _resolve_root():
    env = os.environ.get('GPRTO_REFERENCE_ROOT')
    if env:
        return env
    return SOURCE_ROOT if _has_ref(SOURCE_ROOT) else STAGED_ROOT


REFERENCE_ROOT = _resolve_root()


def stage(force=False):
    """Copy the reference's *.py and *.p files (sub_uts/, run_simple/, run_WO/, plots_RTO.py; ~1.6 MB, no figures) from
    /root/reference into baseline/_ref, unmodified.  No-op when /root/reference is absent (GPU box: already staged)."""
    if not _has_ref(SOURCE_ROOT):
        return STAGED_ROOT if _has_ref(STAGED_ROOT) else None
    n = 0
    for sub in ('', 'sub_uts', 'run_simple', 'run_WO'):
        src_dir = os.path.join(SOURCE_ROOT, sub)
        dst_dir = os.path.join(STAGED_ROOT, sub)
        os.makedirs(dst_dir, exist_ok=True)
        for f in sorted(os.listdir(src_dir)):
            src, dst = os.path.join(src_dir, f), os.path.join(dst_dir, f)
            if not os.path.isfile(src) or not f.endswith(('.py', '.p', '.md')):
                continue
            if force or not os.path.isfile(dst) or os.path.getsize(dst) != os.path.getsize(src) or \
                    os.path.getmtime(dst) < os.path.getmtime(src):
                shutil.copy2(src, dst)
                n += 1
    return STAGED_ROOT


def available():
    return _has_ref(REFERENCE_ROOT)


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
