"""CPU walk-through of the hand-written slab FFT (concept_b200/csrc/pm_fftcore.cuh, pm_fftops.cuh):
the device code is __host__ __device__, so a g++ build of tests/fft_host_harness.cu executes the very
same stage/tile functions sequentially and is checked here against long-double DFTs and numpy.fft."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'concept_b200', 'csrc')


@pytest.fixture(scope='module')
def harness():
    d = tempfile.mkdtemp(prefix='fft_harness_')
    src = os.path.join(d, 'fft_host_harness.cpp')
    with open(os.path.join(ROOT, 'tests', 'fft_host_harness.cu')) as f, open(src, 'w') as g:
        g.write(f.read())
    exe = os.path.join(d, 'fft_harness')
    cuda_inc = '/usr/local/cuda/include'
    subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-I', cuda_inc, '-I', CSRC, src, '-o', exe], check=True)
    return exe, d


def test_stages_against_long_double_dft(harness):
    exe, _ = harness
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'FAIL' not in r.stdout


def _sep_table(G, deconv, gauss):
    l = np.arange(G)
    k = np.where(l >= G//2, l - G, l).astype(float)
    x = k*(np.pi/G) + 2.220446049250313e-16
    return (x/np.sin(x))**deconv*np.exp(-gauss*k*k), k


@pytest.mark.parametrize('dtype,nranks,tol,G', [('f64', 1, 1e-12, 128), ('f64', 4, 1e-12, 128), ('f32', 2, 2e-5, 128), ('f64', 2, 1e-12, 256)])
def test_whole_solve_against_numpy(harness, dtype, nranks, tol, G):
    """z/y forward, x solve with the Green's function, y/z inverse — vs rfftn · factor · irfftn."""
    exe, d = harness
    rng = np.random.default_rng(5)
    rho = rng.standard_normal((G, G, G))
    sep, k = _sep_table(G, 4, 0.0)
    pre = -3.7
    fin, fsep, fout = (os.path.join(d, n) for n in ('in.bin', 'sep.bin', 'out.bin'))
    rho.tofile(fin)
    sep.tofile(fsep)
    r = subprocess.run([exe, 'solve3d', dtype, fin, fsep, fout, repr(pre), str(nranks), str(G)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(fout).reshape(G, G, G)
    # reference: Nyquist planes and origin nullified, separable factor · prefactor/k²
    kk = k[:G//2 + 1].copy()
    kk[-1] = G//2
    sepk = sep[:G//2 + 1]
    k2 = k[:, None, None]**2 + k[None, :, None]**2 + kk[None, None, :]**2
    fac = sep[:, None, None]*sep[None, :, None]*sepk[None, None, :]*pre/np.where(k2 == 0, 1, k2)
    fac[k2 == 0] = 0
    fac[G//2, :, :] = 0
    fac[:, G//2, :] = 0
    fac[:, :, G//2] = 0
    ref = np.fft.irfftn(np.fft.rfftn(rho)*fac, s=(G, G, G), axes=(0, 1, 2), norm='forward')
    err = np.max(np.abs(got - ref))/np.max(np.abs(ref))
    assert err < tol, err
