"""GPU tests of the host mirror (Component / interactions.gravity / drift / timeloop) — the
reference-facing operator surface — against the reference's own outputs (tests/golden)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import pm_oracle as O  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))/np.max(np.abs(b)))


def params_for(d, method='pm'):
    from concept_b200 import commons
    interp = {1: 'NGP', 2: 'CIC', 3: 'TSC', 4: 'PCS'}[int(d['order'])]
    diff = "'fourier'" if int(d['diff_order']) == 0 else int(d['diff_order'])
    il, dc = bool(d['interlace']), bool(d['deconvolve'])
    return commons.load_params(f'''
boxsize = {float(d['boxsize'])!r}*Mpc
potential_options = {{
    'gridsize': {{'gravity': {{'{method}': {int(d['gridsize'])}}}}},
    'interpolation': {{'gravity': {{'{method}': '{interp}'}}}},
    'deconvolve': {{'gravity': {{'{method}': ({dc}, {dc})}}}},
    'interlace': {{'gravity': {{'{method}': ({il}, {il})}}}},
    'differentiation': {{'default': {{'gravity': {{'pm': {diff}, 'p3m': {diff}}}}}}},
}}
select_forces = {{'matter': {{'gravity': '{method}'}}}}
H0 = 70*km/s/Mpc
Ωcdm = 0.25
Ωb = 0.05
a_begin = 0.02
''')


@pytest.mark.parametrize('name', ['kick_pm_cic_G8_d2', 'kick_pm_tsc_G12_d4', 'kick_pm_cic_G12_fourier',
                                  'kick_pm_tsc_G8_interlace', 'kick_pm_cic_G8_nodeconv', 'kick_p3m_long_cic_G24'])
def test_gravity_call_matches_reference(name):
    """The exact call main.kick_long makes: interactions.gravity(method, [c], [c], ᔑdt, 'long-range', True)"""
    from concept_b200 import commons, interactions, mesh
    from concept_b200.species import Component
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    method = str(d['method'])
    params_for(d, method)
    commons.universals.a = float(d['a'])
    N = d['pos'].shape[0]
    c = Component('matter', 'matter', N=N, mass=float(d['mass']))
    for k, s in enumerate('xyz'):
        c.populate(np.ascontiguousarray(d['pos'][:, k]), 'pos' + s)
        c.populate(np.ascontiguousarray(d['mom'][:, k]), 'mom' + s)
    ᔑdt = {'1': float(d['dt_1']), ('a**(-3*w_eff-1)', 'matter'): float(d['dt_rho']), ('a**(-3*w_eff)', 'matter'): float(d['dt_kick'])}
    interactions.gravity(method, [c], [c], ᔑdt, 'long-range', True)
    assert relerr(c.mom_mv3 - d['mom'], d['mom_out'] - d['mom']) < 1e-9
    assert np.array_equal(c.pos_mv3, d['pos'])
    mesh.free_contexts()


def test_two_components_general_path_vs_oracle():
    """Two particle components with different masses: both supply, both receive (general path of
    particle_mesh, interactions.py:2135-2330).  Oracle: the potential of the summed density."""
    from concept_b200 import commons, interactions, mesh
    from concept_b200.species import Component
    G, L = 16, 40.0
    commons.load_params(f'''
boxsize = {L}*Mpc
potential_options = {{'gridsize': {{'gravity': {{'pm': {G}}}}}, 'interpolation': {{'gravity': {{'pm': 'TSC'}}}},
                     'differentiation': {{'default': {{'gravity': {{'pm': 4, 'p3m': 4}}}}}}}}
select_forces = {{'all': {{'gravity': 'pm'}}}}
''')
    commons.universals.a = 0.5
    rng = np.random.default_rng(11)
    comps, data = [], []
    for name, N, mass in (('heavy', 3000, 5.0), ('light', 5000, 0.7)):
        pos, mom = rng.random((N, 3))*L, rng.standard_normal((N, 3))
        c = Component(name, 'matter', N=N, mass=mass)
        c.potential_gridsizes['gravity']['pm'] = (G, G)
        c.populate(pos, 'pos'); c.populate(mom, 'mom')
        comps.append(c); data.append((pos, mom, mass))
    ᔑdt = {'1': 0.02}
    for c in comps:
        ᔑdt['a**(-3*w_eff-1)', c.name] = 0.041
        ᔑdt['a**(-3*w_eff)', c.name] = 0.0199
    interactions.gravity('pm', comps, comps, ᔑdt, 'long-range', False)
    # oracle
    rho = sum(O.deposit(pos, L, G, 3, (0.041/0.02)*mass*float(G)**(-3)*(G/L)**3) for pos, _, mass in data)
    slab = O.forward_fft(rho)*O.potential_factor(G, L, commons.G_Newton, 6)
    phi = O.backward_fft(slab, G)
    for c, (pos, mom, mass) in zip(comps, data):
        ref = mom.copy()
        for dim in range(3):
            ref[:, dim] += O.gather(O.diff_grid(phi, dim, 4, L/G), pos, L, 3)*(mass*(-0.0199))
        assert relerr(c.mom_mv3 - mom, ref - mom) < 1e-9
    mesh.free_contexts()


def test_component_drift_matches_reference():
    from concept_b200 import commons, mesh
    from concept_b200.species import Component
    d = np.load(os.path.join(GOLDEN, 'drift_G8.npz'))
    params_for(np.load(os.path.join(GOLDEN, 'kick_pm_cic_G8_d2.npz')))
    commons.universals.a = float(d['a'])
    c = Component('matter', 'matter', N=d['pos'].shape[0], mass=float(d['mass']))
    c.populate(d['pos'], 'pos'); c.populate(d['mom'], 'mom')
    c.drift({'a**(-2)': float(d['dt_am2']), '1': 0.0123})
    assert np.array_equal(c.pos_mv3, d['pos_out'])
    mesh.free_contexts()


def test_timeloop_reproduces_reference_run():
    """The whole reference run (test/pure_python_pm configuration: 8³ particles, 8³ grid, a = 0.02 → 1,
    141 base steps incl. its adaptive time-step control) replayed through our timeloop.  The reference's
    own tolerances for this test: 1e-10 (compiled vs pure Python, test/pure_python_pm/analyze.py:125) and
    1e-9 (across process counts, test/nprocs_pm/analyze.py:121) on mean |Δx|/boxsize."""
    from concept_b200 import commons, main, mesh
    from concept_b200.species import Component
    d = np.load(os.path.join(GOLDEN, 'run_pm_8.npz'))
    commons.load_params('''
boxsize = 8*Mpc
potential_options = {'gridsize': {'gravity': {'pm': 8}}}
H0      = 70*km/s/Mpc
Ωcdm    = 0.25
Ωb      = 0.05
a_begin = 0.02
output_times = {'snapshot': (0.1, 0.5, 1)}
select_forces = {'matter': {'gravity': 'pm'}}
''')
    N = d['pos0'].shape[0]
    c = Component('matter', 'matter', N=N, mass=float(d['mass']))
    c.populate(d['pos0'], 'pos'); c.populate(d['mom0'], 'mom')
    snaps, steps = {}, []
    def on_dump(components, dump_time):
        snaps[f'{dump_time.a:.6f}'] = (*components[0].gather_global(), commons.universals.t)      # particles by id (the run re-orders them by cell)
    nsteps = main.timeloop([c], on_dump=on_dump, on_step=lambda *a: steps.append(a))
    assert sorted(snaps) == ['0.100000', '0.500000', '1.000000']
    assert nsteps == len(d['drift_dt']) == 142   # same number of base steps (drifts) as the reference
    assert len(steps) == 142
    L = float(d['boxsize'])
    for key, (pos, mom, t) in snaps.items():
        assert t == pytest.approx(float(d[f'snap_t_{key}']), rel=1e-11)
        dx = pos - d[f'snap_pos_{key}']
        dx -= L*np.round(dx/L)
        mean_disp = np.mean(np.sqrt((dx**2).sum(1)))/L
        assert mean_disp < 1e-9, (key, mean_disp)
        assert relerr(mom, d[f'snap_mom_{key}']) < 1e-7
    mesh.free_contexts()
