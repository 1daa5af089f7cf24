import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a box without a CUDA device or driver."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def host_kernels():
    """The element code of the initial-condition / copy kernels (concept_b200/csrc/pm_ic_ops.cuh, pm_copy_ops.cuh)
    compiled for the CPU (tests/ic_host_harness.cu) and loaded as a ctypes library — no GPU needed."""
    import ctypes
    import subprocess
    import tempfile
    tests = os.path.join(ROOT, 'tests')
    d = tempfile.mkdtemp(prefix='ic_harness_')
    src = os.path.join(d, 'ic_host_harness.cpp')
    with open(os.path.join(tests, 'ic_host_harness.cu')) as f, open(src, 'w') as g:
        g.write(f.read())
    lib = os.path.join(d, 'libic_harness.so')
    subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-I', '/usr/local/cuda/include',
                    '-I', os.path.join(ROOT, 'concept_b200', 'csrc'), src, '-o', lib], check=True)
    return ctypes.CDLL(lib)
