"""Initial-condition generator (reference ic.py:928-1399, :1447-1589, :2138-2283) against
golden vectors made by the unmodified reference (tests/golden/gen_golden_ic.py): primordial noise slab,
amplitude tables and the realised particles for sc / bcc / fcc lattices, 1LPT with and without
back-scaling, 2LPT and 3LPT with and without dealiasing, local non-Gaussianity, fixed amplitudes + phase shift, both
noise imprinting schemes and non-default seeds.

CPU: the oracle restatement (oracle/ic_oracle.py), the product's host side (vectorised noise, amplitudes,
linear-theory stand-in) and the orchestration of concept_b200.ic through a numpy model of the kernels.
GPU: concept_b200.ic.realize_particles through libpmgrav.so.

Stated tolerances (fp64): noise bit-exact on the host; positions 1e-11·(lattice spacing), momenta 1e-11 of
max|mom| (measured: ~1e-14 / ~1e-15).
"""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, 'golden', 'ic_*.npz')))
IDS = [os.path.basename(p)[:-4] for p in CASES]


def _golden_noise(d):
    return d['noise'][:, :, 0::2] + 1j*d['noise'][:, :, 1::2]


def _transfers(d):
    A_d, A_t, k0 = (float(x) for x in d['transfer'])
    a = float(d['a'])

    def shape(k):
        return k**2/(1 + (k/k0)**2)**1.1
    return (lambda k: -A_d*a*shape(k)), (lambda k: +A_t*a**0.5*shape(k))


def _growth(d):
    return dict(zip(d['growth_keys'].tolist(), (float(v) for v in d['growth_vals'])))


def _param_text(d):
    return (f"boxsize = {float(d['boxsize'])}*Mpc\nH0 = 70*km/s/Mpc\nΩcdm = 0.25\nΩb = 0.05\na_begin = {float(d['a'])}\n"
            f"primordial_spectrum = {{'A_s': {float(d['A_s'])}, 'n_s': {float(d['n_s'])}, 'α_s': {float(d['alpha_s'])}, "
            f"'pivot': {float(d['pivot'])}/Mpc}}\n"
            f"random_seeds = {{'primordial amplitudes': {int(d['seeds'][0])}, 'primordial phases': {int(d['seeds'][1])}}}\n"
            f"primordial_noise_imprinting = '{str(d['imprinting'])}'\n"
            f"primordial_amplitude_fixed = {bool(d['fixed'])}\nprimordial_phase_shift = {float(d['phase_shift'])!r}\n"
            f"realization_options = {{'backscale': {bool(d['backscale'])}, 'lpt': {int(d['lpt'])}, 'dealias': {bool(d['dealias'])}, 'nongaussianity': {float(d['nongaussianity']) if 'nongaussianity' in d else 0.0}}}\n")


def _expected(d):
    """[(name, species, N, lattice shifts or None, pos, mom, mass)] per realised component of a golden case"""
    n = int(d['n'])
    if 'component_names' in d.files:
        from oracle import ic_oracle as O
        total = len(d['component_names'])
        kind = {2: 2, 4: 4}[total]
        return [(str(name), str(species), n**3, [O.LATTICE_SHIFTS[kind][q]], d[f'pos_{q}'], d[f'mom_{q}'], float(d[f'mass_{q}']))
                for q, (name, species) in enumerate(zip(d['component_names'].tolist(), d['component_species'].tolist()))]
    return [('matter', 'matter', int(d['lattices'])*n**3, None, d['pos'], d['mom'], float(d['mass']))]


def _assert_particles(pos, mom, d, pos_ref=None, mom_ref=None):
    L, n = float(d['boxsize']), int(d['n'])
    pos_ref = d['pos'] if pos_ref is None else pos_ref
    mom_ref = d['mom'] if mom_ref is None else mom_ref
    dp = np.abs(pos - pos_ref)
    dp = np.minimum(dp, L - dp)          # a particle within rounding of the box edge may wrap either way
    assert dp.max() < 1e-11*L/n
    assert np.abs(mom - mom_ref).max() < 1e-11*np.abs(mom_ref).max()


# --------------------------------------------------------------------------------------------- oracle (CPU)
@pytest.mark.parametrize('path', CASES, ids=IDS)
def test_oracle_matches_reference(path):
    from oracle import ic_oracle as O
    d = np.load(path)
    n = int(d['n'])
    noise = O.primordial_noise(n, int(d['seeds'][0]), int(d['seeds'][1]), bool(d['fixed']), float(d['phase_shift']),
                               str(d['imprinting']))
    assert np.array_equal(noise, _golden_noise(d))
    Td, Tt = _transfers(d)
    prim = dict(A_s=float(d['A_s']), n_s=float(d['n_s']), alpha_s=float(d['alpha_s']), pivot=float(d['pivot']))
    for variable, T in ((0, Td), (1, Tt)):
        if f'amplitudes{variable}' in d:
            assert np.allclose(O.get_amplitudes(n, float(d['boxsize']), T, prim), d[f'amplitudes{variable}'], rtol=1e-14, atol=0)
    for name, species, N, shifts, pos_ref, mom_ref, mass in _expected(d):
        pos, mom = O.realize_particles(n, int(d['lattices']), float(d['boxsize']), float(d['a']), float(d['H']), mass,
                                       float(d['w_eff']), noise, Td, Tt, prim, bool(d['backscale']), int(d['lpt']),
                                       bool(d['dealias']), _growth(d), float(d['nongaussianity']) if 'nongaussianity' in d else 0.0,
                                       shifts)
        _assert_particles(pos, mom, d, pos_ref, mom_ref)


def test_cases_cover_the_options():
    seen = {(int(d['lattices']), bool(d['backscale']), int(d['lpt']), bool(d['dealias']), bool(d['fixed']), str(d['imprinting']))
            for d in map(np.load, CASES)}
    assert {s[0] for s in seen} == {1, 2, 4} and {s[2] for s in seen} == {1, 2, 3}
    assert any(s[1] for s in seen) and any(s[3] for s in seen) and any(s[4] for s in seen)
    assert any('nongaussianity' in d and float(d['nongaussianity']) for d in map(np.load, CASES))
    assert {s[5] for s in seen} == {'simple', 'distributed'}


# --------------------------------------------------------------------------------- product, host side (CPU)
@pytest.mark.parametrize('path', CASES, ids=IDS)
def test_host_noise_is_bit_identical_to_reference(path):
    from concept_b200 import commons, ic
    d = np.load(path)
    p = commons.load_params(_param_text(d))
    noise = ic.generate_primordial_noise(int(d['n']), p.primordial_amplitude_fixed, p.primordial_phase_shift)
    ref = _golden_noise(d)
    if bool(d['fixed']) or float(d['phase_shift']):
        assert np.abs(noise - ref).max() < 4e-16      # vectorised cos/sin may differ from the scalar ones by an ulp
    else:
        assert np.abs(noise - ref).max() < 4e-16
        assert np.array_equal(np.abs(noise) > 0, np.abs(ref) > 0)


def test_prng_mirrors_reference_streams():
    """cached scalar draws == whole-array draws == the oracle's restatement (ic.py:67-232)"""
    from concept_b200 import commons, ic
    from oracle import ic_oracle as O
    commons.load_params('boxsize = 8*Mpc\n')
    a, b, c = ic.PseudoRandomNumberGenerator(1000), ic.PseudoRandomNumberGenerator(1000), O.PRNG(1000)
    x = np.array([a.rayleigh(0.5) for _ in range(5000)])
    assert np.array_equal(x, b.rayleigh_array(5000, 0.5))
    assert np.array_equal(x, np.array([c.rayleigh(0.5) for _ in range(5000)]))
    child_a, child_c = a.spawn(2**32 - 3), c.spawn(2**32 - 3)
    u = np.array([child_a.uniform(-np.pi, np.pi) for _ in range(100)])
    assert np.array_equal(u, np.array([child_c.uniform(-np.pi, np.pi) for _ in range(100)]))
    assert np.array_equal(u, ic.PseudoRandomNumberGenerator(1000).spawn(2**32 - 3).uniform_array(100, -np.pi, np.pi))


def test_noise_is_nested_in_grid_size():
    """Enlarging the grid only adds the outer shell (ic.py:944-950)"""
    from concept_b200 import commons, ic
    commons.load_params('boxsize = 8*Mpc\n')
    small, large = ic.generate_primordial_noise(8), ic.generate_primordial_noise(12)
    k = np.arange(-3, 4)
    assert np.array_equal(small[np.ix_(k % 8, k % 8, np.arange(4))], large[np.ix_(k % 12, k % 12, np.arange(4))])
    # Hermitian symmetry of the kk = 0 plane
    for kj in range(-5, 6):
        for ki in range(-5, 6):
            assert large[kj % 12, ki % 12, 0] == np.conj(large[-kj % 12, -ki % 12, 0])


def test_noise_slices_on_threads_equal_the_sequential_noise(monkeypatch):
    """From 64³ on the kj slices are filled by a few host threads (their streams and slab planes are independent): the
    result is the sequential one bit for bit, and the inner 8³ modes still equal the reference-pinned 8³ noise."""
    import os
    from concept_b200 import commons, ic
    commons.load_params('boxsize = 8*Mpc\n')
    monkeypatch.setattr(os, 'cpu_count', lambda: 8)
    threaded = ic.generate_primordial_noise(64)
    monkeypatch.setattr(os, 'cpu_count', lambda: 1)
    sequential = ic.generate_primordial_noise(64)
    assert np.array_equal(threaded, sequential)
    small = ic.generate_primordial_noise(8)
    k = np.arange(-3, 4)
    assert np.array_equal(small[np.ix_(k % 8, k % 8, np.arange(4))], threaded[np.ix_(k % 64, k % 64, np.arange(4))])


def test_fourier_curve_is_a_bijection():
    from concept_b200 import ic
    from oracle import ic_oracle as O
    G = 10
    n = G*G*(G//2 + 1)
    ki, kj, kk = ic.get_fourier_curve_coords(np.arange(n))
    assert len({(a, b, c) for a, b, c in zip(ki.tolist(), kj.tolist(), kk.tolist())}) == n
    assert ki.min() == -G//2 and ki.max() == G//2 - 1 and kk.min() == 0 and kk.max() == G//2
    assert all((ki[q], kj[q], kk[q]) == O.fourier_curve_coords(q) for q in range(n))
    order = ic.fourier_curve_slice_order(G)
    assert list(zip(order[0].tolist(), order[1].tolist())) == O.fourier_curve_slice(G)


class _Spline:
    def __init__(self, f):
        self.f = f

    def eval(self, k):
        return self.f(k)


def _install_golden_linear_theory(monkeypatch, d):
    from concept_b200 import ic
    Td, Tt = _transfers(d)
    growth = _growth(d)

    class Cosmo:
        def __getattr__(self, key):
            if key.startswith('growth_fac_'):
                return lambda a, v=growth[key[len('growth_fac_'):]]: v
            raise AttributeError(key)
    monkeypatch.setattr(ic, 'compute_transfer', lambda component, variable, *args, **kwargs: (_Spline(Td if variable == 0 else Tt), None))
    monkeypatch.setattr(ic, 'compute_cosmo', lambda *args, **kwargs: Cosmo())
    monkeypatch.setitem(ic.n_particles_realized, 'components_tally', 0)
    monkeypatch.setitem(ic.n_particles_realized, 'particles_tally', 0)


@pytest.mark.parametrize('path', CASES, ids=IDS)
def test_amplitudes_match_reference(path, monkeypatch):
    from concept_b200 import commons, ic
    from concept_b200.species import Component
    d = np.load(path)
    commons.load_params(_param_text(d))
    _install_golden_linear_theory(monkeypatch, d)
    c = Component('matter', 'matter', N=int(d['lattices'])*int(d['n'])**3)
    c.realization_options = dict(commons.params.realization_options)
    for variable in (0, 1):
        if f'amplitudes{variable}' in d:
            table = ic.get_amplitudes(int(d['n']), c, float(d['a']), variable=variable)
            assert np.allclose(table, d[f'amplitudes{variable}'], rtol=1e-14, atol=0)


@pytest.mark.parametrize('kernels', ['numpy-model', 'device-code-on-cpu'])
@pytest.mark.parametrize('path', CASES, ids=IDS)
def test_orchestration_through_kernel_model(path, kernels, monkeypatch, host_kernels):
    """concept_b200.ic.realize_particles with the kernels replaced by their numpy model, and by the device code
    itself compiled for the CPU (tests/ic_mock_context.py)"""
    import torch
    from concept_b200 import commons, ic, integration, mesh
    from concept_b200.species import Component
    import ic_mock_context
    MockContext = ic_mock_context.MockContext
    if kernels == 'device-code-on-cpu':
        MockContext = ic_mock_context.HostKernelContext
        monkeypatch.setattr(MockContext, 'lib', host_kernels)
    d = np.load(path)
    commons.load_params(_param_text(d))
    integration.init_time()
    assert integration.hubble(float(d['a'])) == pytest.approx(float(d['H']), rel=1e-13)
    _install_golden_linear_theory(monkeypatch, d)
    contexts = {}
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: contexts.setdefault(
        int(gridsize), MockContext(gridsize, commons.params.boxsize)))
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    expected = _expected(d)
    comps = [Component(name, species, N=N) for name, species, N, *_ in expected]
    id_bgn = 0
    for c, (_, _, N, _, pos_ref, mom_ref, mass) in zip(comps, expected):
        ic.realize_particles(c, float(d['a']), components_all=comps)
        assert c.mass == pytest.approx(mass, rel=1e-13)
        assert c.N_local == c.N == len(pos_ref)
        assert np.array_equal(c.ids[:c.N].numpy(), id_bgn + np.arange(c.N))
        _assert_particles(c.pos[:c.N].numpy(), c.mom[:c.N].numpy(), d, pos_ref, mom_ref)
        id_bgn += N


def test_linear_theory_stand_in():
    """Growth factors have the right matter-era limits, T_EH → 1 on large scales and the normalisation gives
    σ₈ ≈ 0.8 for a Planck-like cosmology (A_s = 2.1e-9) — i.e. units and the ζ → δ relation are right."""
    from concept_b200 import commons, ic, linear
    p = commons.load_params('boxsize = 512*Mpc\nH0 = 67*km/(s*Mpc)\nΩb = 0.049\nΩcdm = 0.27\n')
    cosmo = linear.compute_cosmo()
    a = 1e-3
    assert cosmo.growth_unnormalised(a) == pytest.approx(a, rel=1e-3)
    assert cosmo.growth_fac_D1(1.0) == pytest.approx(1, rel=1e-12)        # normalised like integration.py:1140-1148
    assert cosmo.growth_fac_D1(a) == pytest.approx(a/cosmo.growth_unnormalised(1.0), rel=1e-3)
    assert cosmo.growth_fac_f1(a) == pytest.approx(1, rel=1e-3)
    assert cosmo.growth_fac_f2(a) == pytest.approx(2, rel=1e-3)
    for key, ratio, power in (('D2', 3/7, 2), ('D3a', 1/3, 3), ('D3b', 10/21, 3), ('D3c', 1/7, 3)):
        assert getattr(cosmo, f'growth_fac_{key}')(a)/cosmo.growth_fac_D1(a)**power == pytest.approx(ratio, rel=1e-3)
        assert getattr(cosmo, f'growth_fac_f{key[1:]}')(a) == pytest.approx(power, rel=1e-3)
    Ωm1 = p.Ωm
    assert cosmo.growth_fac_f1(1.0) == pytest.approx(Ωm1**0.55, rel=2e-2)
    assert cosmo.growth_fac_f2(1.0) == pytest.approx(2*Ωm1**(6/11), rel=2e-2)
    assert 0.75 < cosmo.growth_unnormalised(1.0) < 0.82          # Λ suppression of growth for Ωm = 0.319
    assert linear.eisenstein_hu_nowiggle(1e-6) == pytest.approx(1, abs=1e-4)
    k = np.logspace(-4, 1.5, 4000)                       # 1/Mpc
    T, _ = linear.compute_transfer(None, 0, 0, a=1.0)
    power = (T.eval_array(k)*ic.get_primordial_curvature_perturbation(k))**2      # P(k) = |T·ζ|² (ic.py:545-551)
    R = 8/0.67
    x = k*R
    W = 3*(np.sin(x) - x*np.cos(x))/x**3
    f = k**2*power*W**2/(2*np.pi**2)
    σ8 = np.sqrt(np.sum(0.5*(f[1:] + f[:-1])*np.diff(k)))
    assert 0.7 < σ8 < 0.9
    Tθ, _ = linear.compute_transfer(None, 1, 0, a=0.5)
    Tδ, _ = linear.compute_transfer(None, 0, 0, a=0.5)
    from concept_b200.integration import hubble
    assert Tθ.eval(0.1)/Tδ.eval(0.1) == pytest.approx(-0.5*hubble(0.5)*cosmo.growth_fac_f1(0.5), rel=1e-12)
    assert Tδ.eval(0.1) < 0
    # tabulated input takes precedence
    linear.install_transfer(k, -k**2, 3*k)
    try:
        assert linear.compute_transfer(None, 0, 0, a=1.0)[0].eval(0.02) == pytest.approx(-4e-4, rel=1e-6)
        assert linear.compute_transfer(None, 1, 0, a=1.0)[0].eval(0.02) == pytest.approx(0.06, rel=1e-6)
    finally:
        linear.install_transfer(None, None)


def test_unsupported_options_abort():
    from concept_b200 import commons, ic
    from concept_b200.species import Component
    with pytest.raises(commons.ConceptAbort):
        commons.load_params('boxsize = 8*Mpc\nrealization_options = {"lpt": 4}\n')
    commons.load_params('boxsize = 8*Mpc\n')
    with pytest.raises(commons.ConceptAbort):
        ic.realize_particles(Component('matter', 'matter', N=8**3 + 1), 0.02)      # not on a lattice
    with pytest.raises(commons.ConceptAbort):
        ic.realize_particles(Component('matter', 'matter', N=7**3), 0.02)          # odd FFT grid


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize('path', CASES, ids=IDS)
def test_gpu_realize_particles_matches_reference(path, monkeypatch):
    pytest.importorskip('torch')
    from concept_b200 import commons, ic, mesh
    from concept_b200.species import Component
    d = np.load(path)
    commons.load_params(_param_text(d))
    _install_golden_linear_theory(monkeypatch, d)
    expected = _expected(d)
    comps = [Component(name, species, N=N) for name, species, N, *_ in expected]
    results = []
    for c in comps:
        ic.realize_particles(c, float(d['a']), components_all=comps)
        results.append((c.pos_local.cpu().numpy(), c.mom_local.cpu().numpy(), c.ids[:c.N_local].cpu().numpy()))
    mesh.free_contexts()
    id_bgn = 0
    for c, (pos, mom, ids), (_, _, N, _, pos_ref, mom_ref, mass) in zip(comps, results, expected):
        assert np.array_equal(ids, id_bgn + np.arange(c.N))
        assert c.mass == pytest.approx(mass, rel=1e-13)
        _assert_particles(pos, mom, d, pos_ref, mom_ref)
        id_bgn += N


@pytest.mark.gpu
@pytest.mark.parametrize('lpt,dealias', [(1, False), (2, False), (2, True), (3, True)])
def test_gpu_realize_64cubed_against_oracle(lpt, dealias):
    """example_basic's particle load (64³ on a 256 Mpc/h box) with the analytic linear-theory stand-in,
    GPU vs the oracle fed with the same transfer functions and growth factors."""
    pytest.importorskip('torch')
    from concept_b200 import commons, ic, integration, linear, mesh
    from concept_b200.species import Component
    from oracle import ic_oracle as O
    n, a = 64, 0.02
    p = commons.load_params(f'''
boxsize = 256*Mpc/h
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = {a}
realization_options = {{'lpt': {lpt}, 'dealias': {dealias}, 'backscale': {lpt >= 2}}}
''')
    integration.init_time()
    ic.n_particles_realized.update(components_tally=0, particles_tally=0)
    c = Component('matter', 'matter', N=n**3)
    ic.realize_particles(c, a)
    pos, mom = c.pos_local.cpu().numpy(), c.mom_local.cpu().numpy()
    mesh.free_contexts()
    cosmo = linear.compute_cosmo()
    growth = {key: getattr(cosmo, f'growth_fac_{key}')(a) for key in ('D1', 'f1', 'D2', 'f2', 'D3a', 'f3a', 'D3b', 'f3b', 'D3c', 'f3c')}
    Td, Tt = linear.compute_transfer(c, 0, n, a=a)[0], linear.compute_transfer(c, 1, n, a=a)[0]
    ps = p.primordial_spectrum
    prim = dict(A_s=ps['A_s'], n_s=ps['n_s'], alpha_s=ps['α_s'], pivot=ps['pivot'])
    noise = ic.generate_primordial_noise(n)          # pinned to the reference by the CPU tests above
    pos_o, mom_o = O.realize_particles(n, 1, p.boxsize, a, integration.hubble(a), c.mass, 0.0, noise, Td.eval_array,
                                       Tt.eval_array, prim, lpt >= 2, lpt, dealias, growth)
    cell = p.boxsize/n
    dp = np.abs(pos - pos_o)
    dp = np.minimum(dp, p.boxsize - dp)
    assert dp.max() < 1e-10*cell
    assert np.abs(mom - mom_o).max() < 1e-10*np.abs(mom_o).max()
    # the realisation is a small perturbation of the lattice with the linear-theory amplitude: rms displacement
    # of a few per cent of the spacing at a = 0.02
    lattice = (np.stack(np.meshgrid(*[np.arange(n)]*3, indexing='ij'), axis=-1).reshape(-1, 3) + 0.5)*cell
    disp = pos - lattice
    disp -= p.boxsize*np.rint(disp/p.boxsize)
    rms = np.sqrt((disp**2).sum(axis=1).mean())/cell
    assert 0.005 < rms < 0.2


EXAMPLE_BASIC = '''
initial_conditions = {
    'species': 'matter',
    'N'      : 64**3,
}
output_dirs = '%s'
output_times = {'powerspec': 1.0}
boxsize = 256*Mpc/h
potential_options = 128
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = 0.02
primordial_spectrum = {
    'A_s': 2.1e-9,
    'n_s': 0.96,
}
'''   # the settings of param/example_basic


def test_get_initial_conditions_realises_the_component_dict(monkeypatch, tmp_path):
    """main.get_initial_conditions (snapshot.py:3425-3474) for param/example_basic's `initial_conditions` dict,
    kernels replaced by their numpy model: a 64³ lattice displaced by a few per cent of the spacing."""
    import torch
    from concept_b200 import commons, ic, integration, main, mesh
    from concept_b200.species import Component
    from ic_mock_context import MockContext
    p = commons.load_params(EXAMPLE_BASIC % tmp_path)
    integration.init_time()
    contexts = {}
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: contexts.setdefault(
        int(gridsize), MockContext(gridsize, commons.params.boxsize)))
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    components = main.get_initial_conditions()
    assert [c.name for c in components] == ['matter'] and list(contexts) == [64]
    c = components[0]
    assert c.N == c.N_local == 64**3 and c.forces == {'gravity': 'p3m'}
    assert c.mass == pytest.approx(c.ϱ_bar*p.boxsize**3/c.N, rel=1e-14)
    pos, mom = c.pos[:c.N].numpy(), c.mom[:c.N].numpy()
    assert pos.min() >= 0 and pos.max() < p.boxsize
    cell = p.boxsize/64
    lattice = (np.stack(np.meshgrid(*[np.arange(64)]*3, indexing='ij'), axis=-1).reshape(-1, 3) + 0.5)*cell
    disp = pos - lattice
    disp -= p.boxsize*np.rint(disp/p.boxsize)
    assert 0.03 < np.sqrt((disp**2).sum(axis=1).mean())/cell < 0.1
    # Zel'dovich: mom = a·m·u with u = a·H·f·ψ up to the difference between the θ and δ transfer functions (none
    # in the analytic stand-in): the two fields are proportional
    a = commons.universals.a
    cosmo = ic.compute_cosmo()
    expected = a*c.mass*a*integration.hubble(a)*cosmo.growth_fac_f1(a)*disp
    assert np.abs(mom - expected).max() < 1e-9*np.abs(expected).max()


@pytest.mark.gpu
def test_gpu_example_basic_initial_conditions_and_powerspec(tmp_path):
    """param/example_basic up to its first steps on the GPU: realise the 64³ component, measure P(k) of the
    initial conditions with the estimator (PCS, interlaced, deconvolved, grid 128) — it must reproduce the
    linear-theory input below the particle Nyquist frequency — then take two P³M base steps."""
    pytest.importorskip('torch')
    from concept_b200 import commons, ic, integration, linear, main, mesh
    p = commons.load_params(EXAMPLE_BASIC % tmp_path)
    integration.init_time()
    components = main.get_initial_conditions()
    c = components[0]
    assert c.N_local == 64**3
    dump_time = main.DumpTime(commons.universals.a)
    k, power, n_modes = main.dump_powerspec(components, dump_time)
    table = np.loadtxt(os.path.join(str(tmp_path), f'powerspec_a={dump_time.a:.2f}'))
    assert table.shape == (len(k), 4) and np.allclose(table[:, 2], power, rtol=1e-7)
    T, _ = linear.compute_transfer(c, 0, 64, a=commons.universals.a)
    linear_power = (T.eval_array(k)*ic.get_primordial_curvature_perturbation(k))**2
    assert np.allclose(table[:, 3], linear_power, rtol=1e-7)       # the 'linear' column of powerspec_select's default
    k_nyquist = np.pi*64/p.boxsize
    good = (n_modes >= 100) & (k < 0.9*k_nyquist)
    assert good.sum() >= 8
    assert np.abs(power[good]/linear_power[good] - 1).max() < 0.1
    steps = main.timeloop(components, max_steps=2)
    assert steps == 2 and commons.universals.a > p.a_begin
    pos = c.pos_local.cpu().numpy()
    assert np.isfinite(pos).all() and pos.min() >= 0 and pos.max() < p.boxsize
    mesh.free_contexts()


EXAMPLE_BASIC_SHA256 = '3131b62e31673345ad282766d55d53dbb9857ef73c8f8f265775387166be3b34'


def test_vendored_example_basic_is_the_reference_file():
    """tests/golden/example_basic is param/example_basic of the reference, byte for byte (40 lines of parameters, no code)"""
    import hashlib
    with open(os.path.join(HERE, 'golden', 'example_basic'), 'rb') as f:
        assert hashlib.sha256(f.read()).hexdigest() == EXAMPLE_BASIC_SHA256


@pytest.mark.gpu
def test_gpu_example_basic_runs_unmodified_to_a1(tmp_path):
    """BASELINE.json configs[0] / north star: `concept -p param/example_basic` — here `python -m concept_b200 -p` on the
    unmodified file: 64³ particles realised on the fly, P³M on a 128³ grid with 8 rungs, a = 0.02 → 1, power spectrum
    dumped at a = 1.  Physics check in the spirit of test/concept_vs_class_pm/analyze.py:56 (10 %): the large-scale
    power follows linear growth; the small scales have gone non-linear (the short-range force is at work)."""
    import subprocess
    import sys
    import time
    root = os.path.dirname(HERE)
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get('PYTHONPATH', ''))
    t0 = time.time()
    r = subprocess.run([sys.executable, '-m', 'concept_b200', '-p', os.path.join(HERE, 'golden', 'example_basic')], cwd=str(tmp_path),
                       env=env, capture_output=True, text=True, timeout=1500)
    wall = time.time() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'a = 1' in r.stdout.splitlines()[-1]
    table = np.loadtxt(os.path.join(str(tmp_path), 'output', 'example_basic', 'powerspec_a=1.00'))
    k, modes, power, linear_power = table.T
    large = (modes >= 100) & (k < 0.1)
    assert large.sum() >= 2 and np.abs(power[large]/linear_power[large] - 1).max() < 0.1
    small = k > 0.9
    assert (power[small]/linear_power[small]).min() > 4
    print(f'example_basic to a = 1: {wall:.1f} s wall')


@pytest.mark.parametrize('rank', [0, 1])
def test_replicated_realisation_keeps_the_rank_slab(rank, monkeypatch):
    """Two ranks (host logic only, kernels replaced by their numpy model): every rank realises the whole set and
    keeps the particles of its own x-slab — together exactly the single-rank realisation, ids global."""
    import torch
    from concept_b200 import commons, communication, ic, integration, mesh
    from concept_b200.species import Component
    from ic_mock_context import MockContext
    d = np.load(os.path.join(HERE, 'golden', 'ic_1lpt_sc_G8.npz'))
    commons.load_params(_param_text(d) + "potential_options = {'gridsize': {'gravity': {'pm': 8}}}\nselect_forces = {'matter': {'gravity': 'pm'}}\n")
    integration.init_time()
    _install_golden_linear_theory(monkeypatch, d)
    monkeypatch.setattr(communication, 'rank', rank)
    monkeypatch.setattr(communication, 'nprocs', 2)
    monkeypatch.setattr(communication, 'master', rank == 0)
    made = []
    monkeypatch.setattr(ic, '_get_context', lambda gridsize: made.append(gridsize) or MockContext(gridsize, commons.params.boxsize))
    monkeypatch.setattr(mesh, 'get_context', lambda *a, **k: pytest.fail('the shared multi-rank context must not be used'))
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    c = Component('matter', 'matter', N=8**3)
    ic.realize_particles(c, float(d['a']))
    assert made == [8]
    L = float(d['boxsize'])
    mine = np.clip((d['pos'][:, 0]*(8/L)).astype(np.int64), 0, 7)//4 == rank
    assert c.N == 512 and c.N_local == mine.sum() and 0 < c.N_local < 512
    ids = c.ids[:c.N_local].numpy()
    assert np.array_equal(ids, np.nonzero(mine)[0])
    assert np.abs(c.pos[:c.N_local].numpy() - d['pos'][mine]).max() < 1e-11
    assert np.abs(c.mom[:c.N_local].numpy() - d['mom'][mine]).max() < 1e-11*np.abs(d['mom']).max()
    assert c.N_allocated >= c.N_local


def test_growth_factors_match_the_reference_background():
    """linear.CosmoResults against the unmodified reference's own matter + Λ growth factors (integration.py:1104-1290
    through its temporal splines; tests/golden/gen_golden_growth.py).  The reference tabulates on ~4600 points and
    evaluates log–log splines; agreement is at the 1e-4 level of that tabulation."""
    from concept_b200 import commons, linear
    d = np.load(os.path.join(HERE, 'golden', 'growth_factors.npz'))
    commons.load_params('boxsize = 64*Mpc\nH0 = 67*km/(s*Mpc)\nΩb = 0.049\nΩcdm = 0.27\na_begin = 0.02\n')
    assert commons.params.H0 == pytest.approx(float(d['H0']), rel=1e-14) and commons.params.Ωm == pytest.approx(float(d['Om']), rel=1e-14)
    cosmo = linear.compute_cosmo()
    for key in ('D', 'f', 'D2', 'f2', 'D3a', 'f3a', 'D3b', 'f3b', 'D3c', 'f3c'):
        ours = np.array([getattr(cosmo, 'growth_fac_' + {'D': 'D1', 'f': 'f1'}.get(key, key))(a) for a in d['a']])
        assert np.abs(ours/d[key] - 1).max() < 3e-4, key
