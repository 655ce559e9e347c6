import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_sessionstart(session):
    """the built library is git-ignored: a fresh checkout builds it once (nvcc cross-compiles sm_100a without a GPU)"""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, 'bihome_b200', 'libbihome_b200.so')
    if not os.path.isfile(lib) and shutil.which('nvcc') and shutil.which('make'):
        subprocess.run(['make', '-C', os.path.join(ROOT, 'bihome_b200', 'csrc'), '-j8'], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(autouse=True)
def _cudnn_flags_do_not_leak():
    """train.main() turns torch.backends.cudnn.benchmark on for its process (fixed shapes); parity tests that follow in the
    same pytest process compare against fixed references and want cuDNN's default algorithm choice"""
    import torch
    torch.backends.cudnn.benchmark = False
    yield
    torch.backends.cudnn.benchmark = False


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))
    return load


def rel_l2(a, b):
    """norm-wise relative error |a-b|_2 / |b|_2 (b = truth), float64."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def load_entry(name):
    """import <repo root>/<name>.py by path: the oracle's reference importer puts /root/reference (which has its own
    train.py / eval.py) at the head of sys.path"""
    import importlib.util
    key = 'bihome_entry_' + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(ROOT, name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod
