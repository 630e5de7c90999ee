"""Build the CUDA library in-tree:  python -m gp_rto_ei_b200.build  [--force] [-v]

sm_100a only (no fat binary, no fallback arch).  The template instantiations are spread over several translation
units that nvcc compiles in parallel; the objects are linked into gp_rto_ei_b200/libgprto.so, which is git-ignored
but travels to the GPU box with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_obj')
LIB = os.environ.get('GPRTO_LIB_OUT', os.path.join(HERE, 'libgprto.so'))
SOURCES = ['gprto_api.cu', 'k3_inst.cu', 'k2_inst.cu', 'k4_inst_a.cu', 'k4_inst_b.cu', 'k4_inst_c.cu', 'k4_inst_d.cu']
# -fmad=false: a*b+c stays two roundings unless fma() is written - the arithmetic is SPECIFIED (gprto_det_math.h) so that the
# ordered CPU twin (oracle/twin/) reproduces it bit for bit; every hot loop uses explicit fma / DMMA, so this costs nothing
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-fmad=false', '-Xcompiler', '-fPIC']


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


def build_lib(force=False, verbose=False, extra_flags=()):
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    os.makedirs(OBJ, exist_ok=True)
    flags = NVCC_FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if verbose else [])
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ, src.replace('.cu', '.o'))
        cmd = [nvcc] + flags + ['-c', '-o', obj, os.path.join(CSRC, src)]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    for cmd, obj, pr in procs:
        out, _ = pr.communicate()
        log.append(out)
        if pr.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + out)
        objs.append(obj)
    link = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs
    pr = subprocess.run(link, capture_output=True, text=True)
    if pr.returncode != 0:
        raise RuntimeError('link failed:\n' + ' '.join(link) + '\n' + pr.stdout + pr.stderr)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    flags = [a for a in sys.argv[1:] if a.startswith('-D')]
    print(build_lib(force='--force' in sys.argv, verbose='-v' in sys.argv, extra_flags=flags))
