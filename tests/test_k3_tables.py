"""K3 'lc' kernel: the tile-to-warp deal and the per-(step, warp) table the kernel runs on (k3_nlml_lc.cuh, built on the host,
copied to the device once).  The check is host code inside the kernel header's translation unit, so it is compiled with nvcc
for sm_100a but needs no GPU to run."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc not available')
def test_k3_lc_tables_are_consistent_for_every_tile_count(tmp_path):
    exe = str(tmp_path / 'k3_table_check')
    src = os.path.join(REPO, 'tools', 'k3_table_check.cu')
    subprocess.run(['nvcc', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', exe, src], check=True,
                   capture_output=True, text=True, timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    print(out.stdout)
    assert out.returncode == 0 and 'total bad 0' in out.stdout, out.stdout + out.stderr
