import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope='session')
def handle():
    """The CUDA library handle.  No skip and no fallback: on a GPU box a missing library is a failure."""
    from gp_rto_ei_b200 import api, build
    build.build_lib()                      # no-op when libgprto.so is newer than its sources
    return api.Handle(0)


@pytest.fixture(scope='session')
def dev():
    import torch
    return torch.device('cuda', 0)
