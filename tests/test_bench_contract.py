"""The reference arm of bench.py runs on the CPU (C/OpenMP oracle port) and must print one JSON line with the
keys of the measurement contract; the GPU arm's keys are checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
          'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'}


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['impl'] == 'reference' and COMMON <= set(d)
    assert d['value'] > 0 and d['unit'] == 'particle-updates/s' and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and d['vs_baseline'] is None


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '3', '--warmup', '3'],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert COMMON | {'roofline', 'gpu_launches', 'clocks'} <= set(d)
    rf = d['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'kernel'} <= set(rf) and rf['bound'] == 'hbm'
    assert abs(rf['frac'] - rf['achieved']/rf['peak']) < 1e-12
    assert d['gpu_launches'] > 0 and d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['value'] < d['value']
    assert d['cpu_baseline']['value'] > 0
