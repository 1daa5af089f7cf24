"""Pins the short-range (P³M) part of the oracle against the reference's own output
(tests/golden/shortkick_p3m_G24.npz, made by tests/golden/gen_golden_p3m.py)."""
import os

import numpy as np

from oracle import pm_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def relerr(a, b):
    return np.max(np.abs(a - b))/np.max(np.abs(b))


def test_short_kick_matches_reference():
    d = np.load(os.path.join(GOLDEN, 'shortkick_p3m_G24.npz'))
    L, m, G_N, a = float(d['boxsize']), float(d['mass']), float(d['G_Newton']), float(d['a0'])
    nr = int(d['N_rungs'])
    table, maxr2 = O.shortrange_table(float(d['sr_scale']), float(d['sr_range']), int(d['sr_tablesize']), float(d['softening_length']))
    S = O.shortrange_sums(d['pos0'], L, float(d['sr_range']), table, maxr2)
    # real half kick: every particle on its assigned rung (main.kick_short, main.py:1173-1238)
    rung = d['rung_indices_init'].astype(np.int64)
    dmom = S*(G_N*m*m*d['dt_rungs_pair'][rung])[:, None]
    mom = d['mom_after_fake'] + dmom
    assert np.array_equal(d['mom_after_fake'], d['mom0'])          # the fake kick applies nothing
    assert relerr(mom - d['mom0'], d['mom_after_kick'] - d['mom0']) < 1e-12
    acc = dmom*(a**0/(m*(O.MACHINE_EPS + d['dt_rungs_a2'][rung])))[:, None]     # species.py:2290-2325
    assert relerr(acc, d['acc_after_kick']) < 1e-12
    # rung assignment from the fake kick's accelerations (species.py:2415-2435; fac_softening = 0.025)
    got = O.get_rung(d['acc_init'], np.zeros(len(rung), dtype=np.int8), O.rung_factor(float(d['dt']), 0.025, float(d['softening_length'])), nr)
    assert np.array_equal(got, d['rung_indices_init'])
