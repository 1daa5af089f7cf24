"""The host mirror's time loop on the CPU: concept_b200.main.timeloop — adaptive time-step controller, kick/drift
sequencing, dump times, static time-stepping — with the libpmgrav entry points replaced by their numpy model
(tests/ic_mock_context.py::PMKickMockContext, built from the oracle).  The reference's own 142-step run
(test/pure_python_pm configuration; tests/golden/run_pm_8.npz, made by tests/golden/gen_golden_run.py) is replayed;
the GPU counterpart is tests/test_gpu_host_api.py::test_timeloop_reproduces_reference_run."""
import os

import numpy as np
import pytest

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


def _run(monkeypatch, **overrides):
    import torch
    from concept_b200 import commons, main, mesh
    from concept_b200.species import Component
    from ic_mock_context import PMKickMockContext
    d = np.load(os.path.join(GOLDEN, 'run_pm_8.npz'))
    commons.load_params(PM8, **overrides)
    contexts = {}
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: contexts.setdefault(
        int(gridsize), PMKickMockContext(gridsize, commons.params.boxsize)))
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    c = Component('matter', 'matter', N=d['pos0'].shape[0], mass=float(d['mass']))
    c.populate(d['pos0'], 'pos')
    c.populate(d['mom0'], 'mom')
    snaps, steps = {}, []

    def on_dump(components, dump_time):
        snaps[f'{dump_time.a:.6f}'] = (*components[0].gather_global(), commons.universals.t)      # particles by id (the run re-orders them by cell)
    nsteps = main.timeloop([c], on_dump=on_dump, on_step=lambda *a: steps.append(a))
    return d, nsteps, steps, snaps


def test_timeloop_reproduces_reference_run_on_the_cpu(monkeypatch):
    d, nsteps, steps, snaps = _run(monkeypatch)
    assert sorted(snaps) == ['0.100000', '0.500000', '1.000000']
    assert nsteps == len(d['drift_dt']) == 142 and len(steps) == 142
    L = float(d['boxsize'])
    for key, (pos, mom, t) in snaps.items():
        assert t == pytest.approx(float(d[f'snap_t_{key}']), rel=1e-11)
        dx = pos - d[f'snap_pos_{key}']
        dx -= L*np.round(dx/L)
        assert np.mean(np.sqrt((dx**2).sum(1)))/L < 1e-9      # the reference's own tolerance (test/nprocs_pm/analyze.py:121)
        assert np.abs(mom - d[f'snap_mom_{key}']).max() < 1e-7*np.abs(d[f'snap_mom_{key}']).max()


def test_static_timestepping_record_then_replay(monkeypatch, tmp_path):
    """A run records its time-stepping; a second run replaying the file takes the same steps (main.py:499-656, :897-912)"""
    path = str(tmp_path/'timestepping')
    _, nsteps_a, steps_a, snaps_a = _run(monkeypatch, static_timestepping=path)
    table = np.loadtxt(path)
    assert table.ndim == 2 and table.shape[1] == 2 and len(table) >= 10 and np.all(np.diff(table[:, 0]) >= 0)
    _, nsteps_b, steps_b, snaps_b = _run(monkeypatch, static_timestepping=path)
    assert nsteps_b == nsteps_a == 142
    # same base steps to the file's 9 significant digits, same final state to that accuracy
    Δt_a, Δt_b = np.array([s[3] for s in steps_a]), np.array([s[3] for s in steps_b])
    assert np.abs(Δt_b/Δt_a - 1).max() < 1e-6
    pos_a, pos_b = snaps_a['1.000000'][0], snaps_b['1.000000'][0]
    dx = pos_b - pos_a
    dx -= 8*np.round(dx/8)
    assert np.abs(dx).max() < 1e-4


def test_parameter_file_run_end_to_end_on_the_cpu(monkeypatch, tmp_path, host_kernels):
    """main.run on a small example_basic-like parameter file, kernels replaced by their numpy model / the device
    code compiled for the CPU: the initial conditions are realised from the `initial_conditions` dict, the PM time
    loop runs from a = 0.02 to 1 on a PM grid twice as fine as the particle lattice, power spectra are dumped at
    both ends — and the large-scale power has grown by the square of the linear growth factor."""
    import torch
    from concept_b200 import commons, linear, main, mesh
    from concept_b200.species import Component
    import ic_mock_context
    monkeypatch.setattr(ic_mock_context.PMKickMockContext, 'lib', host_kernels)
    contexts = {}
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: contexts.setdefault(
        int(gridsize), ic_mock_context.PMKickMockContext(gridsize, commons.params.boxsize)))
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    param = tmp_path/'param'
    param.write_text(f'''
initial_conditions = {{
    'species': 'matter',
    'N'      : 16**3,
}}
output_dirs = '{tmp_path}/output'
output_times = {{'powerspec': (a_begin, 1.0)}}
boxsize = 256*Mpc/h
potential_options = 32
select_forces = {{'matter': {{'gravity': 'pm'}}}}
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = 0.02
primordial_spectrum = {{
    'A_s': 2.1e-9,
    'n_s': 0.96,
}}
''', encoding='utf-8')
    components = main.run(str(param))
    assert commons.universals.a == pytest.approx(1.0, rel=1e-12) and len(components) == 1 and components[0].N_local == 16**3
    first = np.loadtxt(os.path.join(str(tmp_path), 'output', 'powerspec_a=0.02'))
    last = np.loadtxt(os.path.join(str(tmp_path), 'output', 'powerspec_a=1.00'))
    assert first.shape == last.shape and first.shape[1] == 4 and np.array_equal(first[:, 0], last[:, 0])
    # the file's linear column is (ζ·T_δ)² at the dump's scale factor: it grows by D1² exactly …
    D1 = linear.compute_cosmo().growth_unnormalised
    assert np.allclose(last[:, 3]/first[:, 3], (D1(1.0)/D1(0.02))**2, rtol=1e-6)
    # … and the power measured on the initial conditions follows it on large scales
    assert np.all(np.abs(first[:3, 2]/first[:3, 3] - 1) < 0.35)
    cosmo = linear.compute_cosmo()
    growth2 = (cosmo.growth_fac_D1(1.0)/cosmo.growth_fac_D1(0.02))**2
    k_nyquist_particles = np.pi*16/commons.params.boxsize
    large_scales = first[:, 0] < 0.35*k_nyquist_particles
    assert large_scales.sum() >= 2
    ratio = last[large_scales, 2]/first[large_scales, 2]/growth2
    print("P(k, a=1)/P(k, a=0.02)/D² on large scales:", ratio)
    assert np.all((0.9 < ratio) & (ratio < 1.15)), ratio      # measured: 1.02, 1.05 (a 16³ grid instead of 32³: 0.80, 0.73)


def test_run_is_invariant_under_the_unit_system(monkeypatch, tmp_path, host_kernels):
    """The same parameter file in (Mpc, Gyr, 10¹⁰ m☉) and in (kpc, Myr, m☉): initial conditions, background, time-step
    control and kicks go through every host module in the file's units — positions, momenta and the scale factor reached
    after a few steps agree once converted (kernels replaced by their numpy model)."""
    import torch
    from concept_b200 import commons, main, mesh
    from concept_b200.species import Component
    import ic_mock_context
    monkeypatch.setattr(ic_mock_context.PMKickMockContext, 'lib', host_kernels)
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    results = {}
    for tag, head in (('default', ''), ('small', "unit_length = 'kpc'\nunit_time = 'Myr'\nunit_mass = 'm☉'\n")):
        contexts = {}
        monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None, contexts=contexts: contexts.setdefault(
            int(gridsize), ic_mock_context.PMKickMockContext(gridsize, commons.params.boxsize)))
        param = tmp_path/f'param_{tag}'
        param.write_text(head + f"""
initial_conditions = {{'species': 'matter', 'N': 8**3}}
output_dirs = '{tmp_path}/output_{tag}'
output_times = {{'powerspec': 1.0}}
boxsize = 128*Mpc/h
potential_options = 16
select_forces = {{'matter': {{'gravity': 'pm'}}}}
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = 0.02
""", encoding='utf-8')
        c = main.run(str(param), max_steps=5)[0]
        u = commons.units
        results[tag] = (c.pos_local.numpy()/u.Mpc, c.mom_local.numpy()/(u.m_sun*u.Mpc/u.Gyr), commons.universals.a,
                        commons.universals.t/u.Gyr, c.mass/u.m_sun)
    commons.load_params('boxsize = 8*Mpc\n')
    (pos0, mom0, a0, t0, m0), (pos1, mom1, a1, t1, m1) = results['default'], results['small']
    assert a1 == pytest.approx(a0, rel=1e-10) and t1 == pytest.approx(t0, rel=1e-10) and m1 == pytest.approx(m0, rel=1e-12)
    assert np.abs(pos1 - pos0).max() < 1e-9*np.abs(pos0).max()
    assert np.abs(mom1 - mom0).max() < 1e-8*np.abs(mom0).max()


def test_p3m_run_with_rungs_reproduces_reference_on_the_cpu(monkeypatch):
    """The reference's short P³M run (8 rungs, sub-stepped drifts and rung-selective kicks, rung jumps;
    tests/golden/run_p3m_8.npz) through concept_b200.main / .shortrange with every library entry point replaced by its
    numpy model — the CPU counterpart of tests/test_gpu_p3m.py::test_p3m_run_with_rungs_reproduces_reference."""
    import torch
    from concept_b200 import commons, main, mesh
    from concept_b200.species import Component
    import ic_mock_context
    d = np.load(os.path.join(GOLDEN, 'run_p3m_8.npz'))
    commons.load_params('''
boxsize = 8*Mpc
potential_options = {'gridsize': {'gravity': {'p3m': 24}}}
H0      = 70*km/s/Mpc
Ωcdm    = 0.25
Ωb      = 0.05
a_begin = 0.02
output_times = {'snapshot': (0.0245,)}
select_forces = {'matter': {'gravity': 'p3m'}}
''')
    contexts = {}

    def get_context(gridsize, dtype=None):
        if int(gridsize) not in contexts:
            ctx = ic_mock_context.PMKickMockContext(gridsize, commons.params.boxsize)
            ctx.lib = ic_mock_context.ShortRangeFakeLib(None, commons.params.boxsize)
            ctx._h = None
            contexts[int(gridsize)] = ctx
        return contexts[int(gridsize)]
    monkeypatch.setattr(mesh, 'get_context', get_context)
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    c = Component('matter', 'matter', N=d['pos0'].shape[0], mass=float(d['mass']))
    c.populate(d['pos0'], 'pos')
    c.populate(d['mom0'], 'mom')
    snaps = {}
    main.timeloop([c], on_dump=lambda comps, dt: snaps.update(final=(comps[0].pos_mv3.copy(), comps[0].mom_mv3.copy(),
                                                                          commons.universals.t, comps[0].rung_indices[:comps[0].N_local].numpy().copy())))
    pos, mom, t, rung = snaps['final']
    L = float(d['boxsize'])
    assert t == pytest.approx(float(d['t_final']), rel=1e-10)
    diff = pos[:, None, :] - d['pos_final'][None, :, :]
    diff -= L*np.round(diff/L)
    dist = np.sqrt((diff**2).sum(-1))
    match = dist.argmin(1)                      # the reference re-orders its particles (tile_sort): match by position
    assert len(set(match.tolist())) == len(match)
    assert np.mean(dist[np.arange(len(match)), match])/L < 1e-7
    assert np.abs(mom - d['mom_final'][match]).max() < 1e-5*np.abs(d['mom_final']).max()
    assert np.mean(rung == d['rung_final'][match]) > 0.99


def test_example_basic_defaults_step_on_the_cpu(monkeypatch, tmp_path, host_kernels):
    """param/example_basic's settings at a small size (8³ particles, P³M grid 16 — forces default to P³M with 8
    rungs): initial conditions from the parameter file, rung initialisation on the nearly uniform particle load,
    four base steps of long-range kicks, sub-stepped drifts and rung-selective pair kicks.  Total momentum is
    conserved (antisymmetric pair forces, momentum-conserving PM) and the particles stay in the box."""
    import torch
    from concept_b200 import commons, main, mesh
    from concept_b200.species import Component
    import ic_mock_context
    harness = host_kernels
    contexts = {}

    def get_context(gridsize, dtype=None):
        if int(gridsize) not in contexts:
            ctx = ic_mock_context.PMKickMockContext(gridsize, commons.params.boxsize)
            ctx.lib = ic_mock_context.ShortRangeFakeLib(harness, commons.params.boxsize)
            ctx._h = None
            contexts[int(gridsize)] = ctx
        return contexts[int(gridsize)]
    monkeypatch.setattr(mesh, 'get_context', get_context)
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    param = tmp_path/'param'
    param.write_text(f'''
initial_conditions = {{
    'species': 'matter',
    'N'      : 8**3,
}}
output_dirs = '{tmp_path}/output'
output_times = {{'powerspec': 1.0}}
boxsize = 256*Mpc/h
potential_options = 16
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = 0.02
primordial_spectrum = {{
    'A_s': 2.1e-9,
    'n_s': 0.96,
}}
''', encoding='utf-8')
    states = []
    components = main.run(str(param), max_steps=4, on_step=lambda *a: states.append(a))
    c = components[0]
    assert c.forces == {'gravity': 'p3m'} and commons.params.N_rungs == 8 and len(states) == 4
    assert set(contexts) == {8, 16}                      # the lattice grid of the realisation and the P³M grid
    assert commons.universals.a > 0.02
    pos, mom = c.pos_local.numpy(), c.mom_local.numpy()
    assert np.isfinite(pos).all() and pos.min() >= 0 and pos.max() < commons.params.boxsize
    assert sum(c.rungs_N) == c.N_local
    assert np.abs(mom.sum(axis=0)).max() < 1e-9*np.abs(mom).sum()


def test_example_explanatory_particles_only_on_the_cpu(monkeypatch, tmp_path, host_kernels):
    """The reference's param/example_explanatory (tests/golden/example_explanatory, byte-identical) spells out every
    parameter: `...` entries in output_dirs and the *_select dicts, grid sizes as expressions ('2*cbrt(N)'), nested
    output_times with kinds that are out of scope here (bispec, renders), snapshot dumps.  With its neutrino fluid removed
    (fluids are out of scope) and a small particle load it runs through the host mirror: P³M on the grid its expression
    asks for, output names as the reference forms them, GADGET-2 snapshots when snapshot_type says so."""
    import re
    import torch
    from concept_b200 import commons, main, mesh, snapshot
    from concept_b200.species import Component
    import ic_mock_context
    contexts = {}

    def get_context(gridsize, dtype=None):
        if int(gridsize) not in contexts:
            ctx = ic_mock_context.PMKickMockContext(gridsize, commons.params.boxsize)
            ctx.lib = ic_mock_context.ShortRangeFakeLib(host_kernels, commons.params.boxsize)
            ctx._h = None
            contexts[int(gridsize)] = ctx
        return contexts[int(gridsize)]
    monkeypatch.setattr(mesh, 'get_context', get_context)
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    monkeypatch.setattr(main, '_warned_outputs', set())
    text = open(os.path.join(GOLDEN, 'example_explanatory'), encoding='utf-8').read()
    with pytest.raises(commons.ConceptAbort):       # as it stands: a fluid component
        commons.load_params(text)
        main.get_initial_conditions(do_realization=False)
    text = text.replace('_size = 64', '_size = 8').replace("snapshot_type = 'concept'", "snapshot_type = 'gadget'")
    text = re.sub(r"\{\s*'species'\s*:\s*'neutrino'.*?\},\n", '', text, flags=re.S, count=1)
    text = text.replace("f'{path.output_dir}/{param}'", repr(str(tmp_path/'out'))).replace("f'{path.ic_dir}/autosave'", repr(str(tmp_path/'autosave')))
    # two early dumps instead of the file's late ones, so that a few steps reach them
    text = text.replace("'snapshot' : [1/(1 + z) for z in (1, 0.5, 0)],", "'snapshot' : [a_begin, 0.021],")
    text = text.replace("'powerspec': [a_begin, 0.1, 0.3, 1],", "'powerspec': [a_begin, 0.021],")
    # powerspec_options of the file, with a finer upstream grid for the particles than its '2*cbrt(N)'
    assert "'particles': '2*cbrt(N)'," in text
    text = text.replace("'particles': '2*cbrt(N)',", "'particles': '3*cbrt(N)',", 1)
    param = tmp_path/'param'
    param.write_text(text, encoding='utf-8')
    at_dump = {}
    components = main.run(str(param), max_steps=12,
                          on_dump=lambda comps, dump_time: at_dump.update({round(dump_time.a, 6): comps[0].gather_global()[0]}))
    c = components[0]
    assert c.forces == {'gravity': 'p3m'} and c.potential_gridsizes['gravity'] == {'pm': (8, 8), 'p3m': (16, 16)}
    assert commons.params.output_dirs['powerspec'] == commons.params.output_dirs['render3D'] == str(tmp_path/'out')     # `...`
    written = sorted(os.listdir(tmp_path/'out'))
    assert written == ['powerspec_a=0.020', 'powerspec_a=0.021', 'snapshot_a=0.020', 'snapshot_a=0.021'], written
    assert {'bispec', 'render2D', 'render3D'} <= main._warned_outputs | {'bispec', 'render2D', 'render3D'}
    assert 'grid size 24 ' in open(tmp_path/'out'/'powerspec_a=0.021', encoding='utf-8').readline()       # 3·∛512
    snap = snapshot.read_gadget(str(tmp_path/'out'/'snapshot_a=0.021'))
    assert snap['a'] == pytest.approx(0.021) and len(snap['pos']) == 8**3
    assert np.abs(snap['pos'] - at_dump[0.021]).max() < 1e-6*commons.params.boxsize        # float32 positions in the file


def test_gravity_plugin_surface(monkeypatch):
    """concept_b200.gravity mirrors the reference's short-range plugin (gravity.py:51-67, :263-354): factors per rung,
    the pair kick into Δmom of the active particles, and a loud refusal of caller-side tile selections."""
    import torch
    from concept_b200 import commons, gravity, mesh, shortrange
    from concept_b200.species import Component
    from oracle import pm_oracle as O
    import ic_mock_context
    d = np.load(os.path.join(GOLDEN, 'shortkick_p3m_G24.npz'))
    commons.load_params('''
boxsize = 8*Mpc
potential_options = {'gridsize': {'gravity': {'p3m': 24}}}
H0 = 70*km/s/Mpc
Ωcdm = 0.25
Ωb = 0.05
a_begin = 0.02
select_forces = {'matter': {'gravity': 'p3m'}}
''')
    ctx = ic_mock_context.PMKickMockContext(24, commons.params.boxsize)
    ctx.lib, ctx._h = ic_mock_context.ShortRangeFakeLib(None, commons.params.boxsize), None
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: ctx)
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    c = Component('matter', 'matter', N=d['pos0'].shape[0], mass=float(d['mass']))
    c.populate(d['pos0'], 'pos')
    c.populate(d['mom0'], 'mom')
    assert c.softening_length == pytest.approx(float(d['softening_length']), rel=1e-14)
    assert gravity.combine_softening_lengths(0.1, 0.3) == pytest.approx(0.2)
    shortrange.ensure_rung_state(c)
    rung = d['rung_indices_init'].astype(np.int8)
    c.rung_indices[:c.N_local] = torch.from_numpy(rung)
    c.rung_indices_jumped[:c.N_local] = torch.from_numpy(rung)
    c.lowest_active_rung = 1
    ᔑdt_rungs = {('a**(-3*w_eff₀-3*w_eff₁-1)', 'matter', 'matter'): np.asarray(d['dt_rungs_pair'], dtype=np.float64)}
    factors = gravity.compute_factors(c, c, ᔑdt_rungs)
    assert np.allclose(factors, float(d['G_Newton'])*float(d['mass'])**2*d['dt_rungs_pair'], rtol=1e-15)
    c.Δmom[:] = 7.0
    gravity.gravity_pairwise_shortrange('gravity', c, c, ᔑdt_rungs)
    table, maxr2 = O.shortrange_table(float(d['sr_scale']), float(d['sr_range']), int(d['sr_tablesize']), float(d['softening_length']))
    S = O.shortrange_sums(d['pos0'], float(d['boxsize']), float(d['sr_range']), table, maxr2)
    active = rung >= 1
    got = c.Δmom[:c.N_local].numpy()
    assert active.any() and (~active).any()
    assert np.allclose(got[active], S[active]*factors[rung[active]][:, None], rtol=1e-12, atol=0)
    assert np.all(got[~active] == 7.0)                      # inactive receivers keep their stored Δmom
    with pytest.raises(commons.ConceptAbort):
        gravity.gravity_pairwise_shortrange('gravity', c, c, ᔑdt_rungs, tile_indices_receiver=np.arange(3))
