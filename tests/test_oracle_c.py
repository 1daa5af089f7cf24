"""Pins the C/OpenMP restatement (oracle/pm_oracle.c — the CPU baseline) against the reference's
golden vectors and against the numpy oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import pm_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
# the C port covers the default path: no interlacing, real-space differentiation
KICKS = [n for n in sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'kick_*.npz')))
         if 'interlace' not in n and 'fourier' not in n]


def relerr(a, b):
    return np.max(np.abs(a - b))/np.max(np.abs(b))


@pytest.mark.parametrize('name', KICKS)
def test_c_kick_matches_reference(name):
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    pos, mom = np.ascontiguousarray(d['pos']), np.ascontiguousarray(d['mom']).copy()
    C.kick_long(pos, mom, mass=float(d['mass']), boxsize=float(d['boxsize']), gridsize=int(d['gridsize']), order=int(d['order']),
                G_Newton=float(d['G_Newton']), dt_rho_over_dt1=float(d['dt_rho'])/float(d['dt_1']), dt_kick=float(d['dt_kick']),
                diff_order=int(d['diff_order']), deconvolve=bool(d['deconvolve']),
                r_scale=float(d['r_scale']) if 'r_scale' in d.files else 0.0)
    assert relerr(mom - d['mom'], d['mom_out'] - d['mom']) < 1e-11


def test_c_deposit_and_drift():
    d = np.load(os.path.join(GOLDEN, 'kick_pm_pcs_G10_d6.npz'))
    G, L = int(d['gridsize']), float(d['boxsize'])
    rho = C.deposit(np.ascontiguousarray(d['pos']), G, L, 4, 1.0)
    ref = O.deposit(d['pos'], L, G, 4, 1.0)
    assert relerr(rho, ref) < 1e-13
    d = np.load(os.path.join(GOLDEN, 'drift_G8.npz'))
    pos = np.ascontiguousarray(d['pos']).copy()
    C.drift(pos, np.ascontiguousarray(d['mom']), float(d['dt_am2'])/float(d['mass']), float(d['boxsize']))
    assert np.array_equal(pos, d['pos_out'])


def test_c_vs_numpy_oracle_pow2_threads():
    """power-of-two grid (radix-2 path) with enough particles to exercise the threaded slabs"""
    G, L, N = 32, 50.0, 40000
    rng = np.random.default_rng(3)
    pos = rng.random((N, 3))*L
    mom = rng.standard_normal((N, 3))
    kw = dict(mass=1.7, boxsize=L, gridsize=G, order=2, G_Newton=4.4985024439973154e-05, dt_rho_over_dt1=2.0, dt_kick=0.01)
    ref = O.pm_kick(pos, mom, **kw)
    got = C.kick_long(pos, mom.copy(), **kw)
    assert relerr(got - mom, ref - mom) < 1e-11
    assert abs(C.sum_mom2(got) - O.sum_mom2(got)) < 1e-12*O.sum_mom2(got)
