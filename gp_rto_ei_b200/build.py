"""Build the CUDA library in-tree:  python -m gp_rto_ei_b200.build  [--force]

One nvcc invocation, sm_100a only (no fat binary, no fallback arch).  The resulting
gp_rto_ei_b200/libgprto.so is git-ignored but travels to the GPU box with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgprto.so')
SOURCES = ['gprto_api.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '--threads', '0']


def _deps():
    out = [os.path.join(HERE, '..', 'include', 'gprto.h')]
    for f in os.listdir(CSRC):
        if f.endswith(('.cu', '.cuh', '.h')):
            out.append(os.path.join(CSRC, f))
    return out


def up_to_date():
    if not os.path.isfile(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in _deps())


def build_lib(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == '__main__':
    print(build_lib(force='--force' in sys.argv, verbose='-v' in sys.argv))
