import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope='session')
def cfg():
    with open(os.path.join(ROOT, 'config', 'faghri.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def golden():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden.npz'))


@pytest.fixture(scope='session')
def golden_params():
    with open(os.path.join(ROOT, 'tests', 'golden', 'reference_params.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def built_lib():
    """libhpsolver.so, (re)built with nvcc when sources are newer (cross-compiles without a GPU)."""
    from heat_pipe_solver_b200 import build
    return build.build()


@pytest.fixture(scope='session')
def hostcheck(tmp_path_factory):
    """The device property header compiled for the host (tests/hostcheck)."""
    import ctypes as C
    out = tmp_path_factory.mktemp('hostcheck') / 'libhostcheck.so'
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-x', 'c++',
                           '-I' + os.path.join(ROOT, 'include'), '-I' + os.path.join(ROOT, 'heat_pipe_solver_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'hostcheck', 'hostcheck.cpp'), '-o', str(out)])
    return C.CDLL(str(out))
