"""Multi-GPU parity (needs ≥ 2 GPUs on the box; skipped otherwise).  Launches tests/mgpu_check.py
under torchrun with 2 ranks (and 4 when available)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('nranks', [2, 4])
def test_distributed_path_matches_oracle(nranks):
    if torch.cuda.device_count() < nranks:
        pytest.skip(f'needs {nranks} GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nranks}',
           '--master-addr', '127.0.0.1', '--master-port', str(29500 + 11*nranks), os.path.join(ROOT, 'tests', 'mgpu_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert 'MGPU_CHECK PASSED' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
