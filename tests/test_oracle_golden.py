"""Pins oracle/pm_oracle.py against outputs of the reference itself
(tests/golden/*.npz, made by tests/golden/gen_golden.py from /root/reference)."""
import glob
import os

import numpy as np
import pytest

from oracle import pm_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
KICKS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'kick_*.npz')))


def relerr(a, b):
    return np.max(np.abs(a - b))/np.max(np.abs(b))


def test_golden_files_present():
    assert len(KICKS) >= 13


@pytest.mark.parametrize('name', KICKS)
def test_kick_matches_reference(name):
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    taps = {}
    mom = O.pm_kick(
        d['pos'], d['mom'], mass=float(d['mass']), boxsize=float(d['boxsize']), gridsize=int(d['gridsize']),
        order=int(d['order']), G_Newton=float(d['G_Newton']), dt_rho_over_dt1=float(d['dt_rho'])/float(d['dt_1']),
        dt_kick=float(d['dt_kick']), diff_order=int(d['diff_order']), deconvolve=bool(d['deconvolve']),
        interlace=bool(d['interlace']), r_scale=float(d['r_scale']) if 'r_scale' in d.files else 0.0, taps=taps,
    )
    dmom_ref = d['mom_out'] - d['mom']
    dmom = mom - d['mom']
    # the kick itself (not mom, which is dominated by the unchanged part) to 1e-12 relative
    assert relerr(dmom, dmom_ref) < 1e-12, name
    assert relerr(taps['rho'], d['tap_rho']) < 1e-13
    if 'tap_phi' in d.files:
        assert relerr(taps['phi'], d['tap_phi']) < 1e-12
    for dim in range(3):
        assert relerr(taps[f'forcegrid{dim}'], d[f'tap_forcegrid{dim}']) < 1e-11
    if bool(d['interlace']):
        assert relerr(taps['rho_shifted'], d['tap_rho_shifted']) < 1e-13


def test_drift_matches_reference():
    d = np.load(os.path.join(GOLDEN, 'drift_G8.npz'))
    out = O.drift(d['pos'], d['mom'], float(d['dt_am2'])*float(d['a'])**0/float(d['mass']), float(d['boxsize']))
    assert np.array_equal(out, d['pos_out'])
    assert out.min() >= 0 and out.max() < float(d['boxsize'])


def test_reference_slab_layout_shape():
    G = 8
    rng = np.random.default_rng(0)
    f = O.forward_fft(rng.standard_normal((G, G, G)))
    s = O.to_reference_slab_layout(f)
    assert s.shape == (G, G, G + 2)
    assert s[3, 5, 2*2] == f[5, 3, 2].real and s[3, 5, 2*2 + 1] == f[5, 3, 2].imag
