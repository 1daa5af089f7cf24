"""CPU-only tests of the host-side mirror: parameter files, units, background cosmology,
time-step integrals — against values recorded from the reference's own full run
(tests/golden/run_pm_8.npz, made by tests/golden/gen_golden_run.py)."""
import os

import numpy as np
import pytest

from concept_b200 import commons, integration

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
PM8 = '''
boxsize = 8*Mpc
potential_options = {'gridsize': {'gravity': {'pm': 8}}}
H0      = 70*km/s/Mpc
Ωcdm    = 0.25
Ωb      = 0.05
a_begin = 0.02
output_times = {'snapshot': (0.1, 0.5, 1)}
select_forces = {'matter': {'gravity': 'pm'}}
'''


def test_units_match_reference():
    d = np.load(os.path.join(GOLDEN, 'run_pm_8.npz'))
    assert commons.G_Newton == float(d['G_Newton'])          # 4.4985024439973154e-05 Mpc³ (10¹⁰ m☉)⁻¹ Gyr⁻²
    p = commons.load_params(PM8)
    assert p.H0 == pytest.approx(float(d['H0']), rel=1e-15)
    assert p.ρ_crit == pytest.approx(float(d['rho_crit']), rel=1e-14)
    assert p.boxsize == 8.0 and p.Ωm == pytest.approx(0.3)


def test_param_file_forward_references_and_h():
    p = commons.load_params('''
boxsize = 256*Mpc/h
potential_options = 2*_size
_size = 64
H0 = 67*km/(s*Mpc)
initial_conditions = {'species': 'matter', 'N': _size**3}
''')
    assert p.boxsize == pytest.approx(256/0.67, rel=1e-12)
    assert commons.gridsize_for('p3m', 64**3) == 128 and commons.gridsize_for('pm', 64**3) == 128
    assert p.interpolation_order == {'pm': 2, 'p3m': 2}
    assert p.differentiation == {'pm': 2, 'p3m': 4}
    assert p.deconvolve['pm'] == (True, True) and p.interlace['pm'] == (False, False)


def test_example_basic_defaults_to_p3m():
    """param/example_basic sets only potential_options = 128 ⇒ particles use P³M (commons.py:3664-3702)"""
    commons.load_params('''
initial_conditions = {'species': 'matter', 'N': 64**3}
boxsize = 256*Mpc/h
potential_options = 128
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = 0.02
''')
    from concept_b200.species import Component
    c = Component('matter', 'matter', N=64**3, mass=1.0)
    assert c.forces == {'gravity': 'p3m'}
    assert c.potential_gridsizes['gravity']['p3m'] == (128, 128)
    assert c.potential_differentiations['gravity']['p3m'] == 4


def test_potential_options_variants():
    p = commons.load_params('''
boxsize = 30*Mpc
potential_options = {
    'gridsize': {'gravity': {'pm': 12}},
    'interpolation': {'gravity': {'pm': 'TSC'}},
    'deconvolve': {'gravity': {'pm': (False, False)}},
    'interlace': {'gravity': {'pm': (True, True)}},
    'differentiation': {'default': {'gravity': {'pm': 'fourier', 'p3m': 6}}},
}
''')
    assert p.interpolation_order['pm'] == 3 and p.deconvolve['pm'] == (False, False)
    assert p.interlace['pm'] == (True, True) and p.differentiation == {'pm': 0, 'p3m': 6}
    assert commons.gridsize_for('pm', 1000) == 12


def test_gridsize_expressions_like_the_reference():
    """Grid sizes may be expressions in N, Ñ, gridsize and nprocs (Component.__init__'s to_float, species.py:1134-1157; the
    reference's own param/example_explanatory says 'p3m': ('2*cbrt(N)', '2*cbrt(N)')); the default is ∛Ñ (2·∛Ñ for P³M) with Ñ the
    cells of the pre-initial lattice, also for bcc (N = 2n³) and fcc (N = 4n³) loads (species.py:1192-1207)."""
    commons.load_params("""
boxsize = 100*Mpc
potential_options = {'gridsize': {'global': {'gravity': {'pm': 64, 'p3m': 128}},
                                  'matter': {'gravity': {'p3m': ('2*cbrt(N)', '2*cbrt(N)')}},
                                  'cdm': {'gravity': {'pm': 'gridsize', 'p3m': -1}}}}
""")
    assert commons.component_gridsizes('matter', 'matter', 'p3m', 64**3) == (128, 128)
    assert commons.component_gridsizes('cdm', 'cold dark matter', 'pm', 2*24**3) == (24, 24)        # ∛Ñ of a bcc load
    assert commons.component_gridsizes('cdm', 'cold dark matter', 'p3m', 16**3) == (128, 128)      # -1: the default (global)
    assert commons.gridsize_value('cbrt(Ñ)', 4*10**3) == 10 and commons.gridsize_value('2*cbrt(N) + 2*nprocs', 8**3) == 18
    assert commons.gridsize_value(47.6, 0) == 48
    commons.load_params('boxsize = 100*Mpc\n')
    assert commons.gridsize_for('pm', 2*24**3) == 24 and commons.gridsize_for('p3m', 4*10**3) == 20
    with pytest.raises(commons.ConceptAbort):
        commons.gridsize_value('2*cuberoot(N)', 8**3)


def test_softening_length_units_and_kernel_parameters():
    """select_softening_length (commons.py:3862-3873) by name / species / 'particles' / 'all', as a length or an expression;
    a unit system or softening kernel other than the reference's defaults aborts instead of being ignored."""
    from concept_b200.species import Component
    commons.load_params("boxsize = 100*Mpc\nselect_softening_length = {'matter': '0.03*boxsize/cbrt(N)', 'all': 2*kpc}\n")
    assert Component('matter', 'matter', N=1000, mass=1).softening_length == pytest.approx(0.3, rel=1e-14)
    assert Component('b', 'baryons', N=1000, mass=1).softening_length == pytest.approx(0.002, rel=1e-14)
    commons.load_params('boxsize = 100*Mpc\n')
    assert Component('matter', 'matter', N=1000, mass=1).softening_length == 0.025*100.0/1000**(1/3)
    with pytest.raises(commons.ConceptAbort):
        commons.load_params("boxsize = 100*Mpc\nsoftening_kernel = 'plummer'\n")
    commons.load_params('boxsize = 100*Mpc\n')


def test_unit_system_of_the_parameter_file():
    """unit_length / unit_time / unit_mass (commons.py:1935-1999; the reference's test/realize/param uses kpc): every number
    of the run is expressed in them — lengths scale, Newton's constant and the critical density follow, rates do not change —
    and the next parameter file starts from the defaults again."""
    from concept_b200.species import Component
    base = 'boxsize = 100*Mpc\nH0 = 70*km/(s*Mpc)\nΩb = 0.05\nΩcdm = 0.25\n'
    p0 = commons.load_params(base)
    G0, m0 = commons.G_Newton, Component('matter', 'matter', N=1000, mass=-1).softening_length
    p1 = commons.load_params("unit_length = 'kpc'\nunit_mass = 'm☉'\n" + base)
    assert commons.unit_length == 'kpc' and commons.unit_mass == 'm☉'
    assert p1.boxsize == pytest.approx(1e3*p0.boxsize, rel=1e-14) and p1.H0 == pytest.approx(p0.H0, rel=1e-14)
    assert commons.G_Newton == pytest.approx(G0*1e9/1e10, rel=1e-13)                 # length³ / mass
    assert p1.ρ_crit == pytest.approx(p0.ρ_crit*1e10/1e9, rel=1e-13)                 # mass / length³
    assert Component('matter', 'matter', N=1000, mass=-1).softening_length == pytest.approx(1e3*m0, rel=1e-14)
    with pytest.raises(commons.ConceptAbort):
        commons.load_params("unit_length = 'furlong'\n" + base)
    commons.load_params(base)
    assert commons.unit_length == 'Mpc' and commons.G_Newton == G0


def test_differentiation_order_per_component():
    """potential_options['differentiation'] keyed by component (the reference's test/concept_vs_class_pm/param asks for order 4
    for 'matter'), species.py:1217-1236"""
    from concept_b200.species import Component
    commons.load_params("""
boxsize = 100*Mpc
potential_options = {'gridsize': {'gravity': {'pm': 16}},
                     'differentiation': {'matter': {'gravity': {'pm': 4}}, 'baryons': {'gravity': {'pm': 'fourier'}}}}
""")
    assert Component('matter', 'matter', N=512, mass=1).potential_differentiations['gravity'] == {'pm': 4, 'p3m': 4}
    assert Component('b', 'baryons', N=512, mass=1).potential_differentiations['gravity'] == {'pm': 0, 'p3m': 4}
    assert Component('cdm', 'cold dark matter', N=512, mass=1).potential_differentiations['gravity'] == {'pm': 2, 'p3m': 4}


def test_shortrange_parameters_as_expressions():
    """shortrange_params (commons.py:3254-3269): scale and range as lengths or as the expressions the reference's
    example_explanatory spells out ('1.25*boxsize/gridsize', '4.5*scale'), with or without the 'gravity' level."""
    from concept_b200 import shortrange
    commons.load_params("boxsize = 100*Mpc\nshortrange_params = {'gravity': {'scale': '1.5*boxsize/gridsize', 'range': '4.0*scale', 'tablesize': 2048}}\n")
    assert shortrange.shortrange_params(50) == (3.0, 12.0, 2048)
    commons.load_params('boxsize = 100*Mpc\n')
    assert shortrange.shortrange_params(50) == (2.5, 11.25, 4096)
    commons.load_params("boxsize = 100*Mpc\nshortrange_params = {'scale': 3*Mpc}\n")
    assert shortrange.shortrange_params(50) == (3.0, 13.5, 4096)


def test_background_and_time_step_integrals_match_reference_run():
    """Every ᔑdt the reference used in its 160 kicks / 142 drifts is reproduced from our own
    background (same ODE solver settings, same natural cubic splines)."""
    d = np.load(os.path.join(GOLDEN, 'run_pm_8.npz'))
    commons.load_params(PM8)
    integration.init_time(reinitialize=True)
    assert commons.universals.a == 0.02
    for key in ('0.100000', '0.500000', '1.000000'):
        assert integration.cosmic_time(float(key)) == pytest.approx(float(d[f'snap_t_{key}']), rel=1e-11)

    class C:
        name = 'matter'
        def w_eff(self, a=-1, t=-1): return 0.0
    comps = [C()]
    # kicks: t_start = kick_t, t_end = t_start + ᔑdt['1']
    for i in (0, 1, 2, 50, 100, 159):
        t0, dt1 = float(d['kick_t'][i]), float(d['kick_dt1'][i])
        assert integration.scalefactor_integral('1', t0, t0 + dt1, comps) == pytest.approx(dt1, rel=1e-12)
        got = integration.scalefactor_integral(('a**(-3*w_eff-1)', 'matter'), t0, t0 + dt1, comps)
        assert got == pytest.approx(float(d['kick_dtrho'][i]), rel=1e-10)
        got = integration.scalefactor_integral(('a**(-3*w_eff)', 'matter'), t0, t0 + dt1, comps)
        assert got == pytest.approx(float(d['kick_dtkick'][i]), rel=1e-10)


def test_spline_matches_scipy_natural():
    x = np.linspace(1, 3, 7)
    s = integration.Spline(x, x**2, 'x²')
    assert s.eval(2.0) == pytest.approx(4.0, rel=1e-3)
    assert s.integrate(1, 3) == pytest.approx(26/3, rel=1e-3)
    assert s.integrate(3, 1) == -s.integrate(1, 3)
    with pytest.raises(SystemExit):
        s.eval(10.0)     # outside the tabulated interval: abort, like the reference


def test_static_timestepping_callable_file_replay_and_recording(tmp_path):
    """main.prepare_static_timestepping (reference main.py:499-656) and the recording in get_base_timestep_size
    (:897-912): Δt follows Δa(a) from a callable; a recorded file is replayed exactly at its recorded scale
    factors and log–log interpolated in between, stretch by stretch."""
    from concept_b200 import main
    from concept_b200.commons import universals
    from concept_b200.integration import cosmic_time, init_time, scale_factor
    commons.load_params(PM8, static_timestepping=lambda a: 0.01*a)
    init_time()
    f = main.prepare_static_timestepping()
    Δt, bottleneck = main.get_base_timestep_size([], f)
    assert bottleneck == main.bottleneck_static_timestepping
    assert Δt == pytest.approx(cosmic_time(0.02*1.01) - universals.t, rel=1e-12)
    assert f(0.995) == commons.ထ                                     # a + Δa beyond a = 1
    # recording: no file yet ⇒ the controller runs as usual and appends "a Δa" lines
    path = str(tmp_path/'sub'/'timestepping')
    commons.load_params(PM8, static_timestepping=path)
    init_time()
    assert main.prepare_static_timestepping() is None and os.path.isdir(os.path.dirname(path))

    class Matter:
        name, representation, forces, ϱ_bar = 'matter', 'particles', {}, commons.params.Ωm*commons.params.ρ_crit

        def w_eff(self, a=-1):
            return 0.0
    recorded = []
    for a in (0.02, 0.03, 0.05):
        universals.a, universals.t = a, cosmic_time(a)
        Δt, _ = main.get_base_timestep_size([Matter()])
        recorded.append((a, scale_factor(universals.t + Δt) - a))
    lines = open(path, encoding='utf-8').read().split('\n')
    assert lines[0].startswith('# Time-stepping recorded by') and lines[2].split() == ['#', 'a', 'Δa']
    table = np.loadtxt(path)
    assert np.allclose(table, np.array(recorded), rtol=1e-9)
    # replay
    commons.load_params(PM8, static_timestepping=path)
    init_time()
    f = main.prepare_static_timestepping()
    assert callable(f)
    for a, Δa in recorded:
        universals.a, universals.t = a, cosmic_time(a)
        assert f() == pytest.approx(cosmic_time(a + Δa) - universals.t, rel=1e-8)
    a_mid = 0.04
    Δa_mid = np.exp(np.interp(np.log(a_mid), np.log(table[:, 0]), np.log(table[:, 1])))
    assert f(a_mid) == pytest.approx(cosmic_time(a_mid + Δa_mid) - cosmic_time(a_mid), rel=1e-9)
    with pytest.raises(commons.ConceptAbort):
        commons.load_params(PM8, static_timestepping=str(tmp_path))
        main.prepare_static_timestepping()


def test_output_times_forms():
    """output_times as {kind: scale factors} and as {'a': {...}, 't': {...}} with None entries (param/example_explanatory)"""
    from concept_b200 import main
    from concept_b200.integration import cosmic_time, init_time
    commons.load_params(PM8, output_times={'snapshot': (0.1, 0.5, 1), 'powerspec': 0.5})
    init_time()
    times = main._dump_times()
    assert [d.a for d in times] == [0.1, 0.5, 1.0] and all(d.time_param == 'a' for d in times)
    assert main._wanted('powerspec', times[1]) and not main._wanted('powerspec', times[0]) and main._wanted('snapshot', times[2])
    t_half = cosmic_time(0.25)
    commons.load_params(PM8, output_times={'a': {'snapshot': [0.5, 1.0], 'powerspec': None, 'render2D': 1},
                                           't': {'snapshot': None, 'powerspec': t_half}})
    init_time()
    times = main._dump_times()
    assert [d.time_param for d in times] == ['t', 'a', 'a']
    assert times[0].a == pytest.approx(0.25, rel=1e-9) and main._wanted('powerspec', times[0]) and not main._wanted('snapshot', times[0])
    assert main._wanted('snapshot', times[1]) and not main._wanted('powerspec', times[1])
