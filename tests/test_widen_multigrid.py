"""PM / P³M long-range kicks with component-specific upstream and downstream grid sizes (SURVEY §8 a5/a10/a16;
reference particle_mesh interactions.py:1985-2335, interpolate_upstream mesh.py:492-616,
add_upstream_to_global_slabs :618-710, copy_modes :980-1322) against golden vectors made by the unmodified reference
(tests/golden/gen_golden_multigrid.py): two particle components, upstream/downstream grids smaller and larger than
the global one, CIC/TSC/PCS, finite-difference and Fourier differentiation, interlacing, the Gaussian-split P³M
potential, with and without deconvolution.

CPU: the oracle restatement (oracle/pm_oracle.py::pm_kick_multigrid) and the host orchestration of
concept_b200.interactions with the kernels replaced by a numpy model (copy_modes: the device code on the CPU).
GPU: interactions.gravity through libpmgrav.so.

Tolerance: Δmom to 1e-9 of max|Δmom| (the kicks are ~1e-4 of the momenta, so Δmom itself carries ~1e-12)."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, 'golden', 'multigrid_*.npz')))
IDS = [os.path.basename(p)[:-4] for p in CASES]
ORDER_NAME = {1: 'NGP', 2: 'CIC', 3: 'TSC', 4: 'PCS'}
SPECIES = {'cdm': 'cold dark matter', 'baryons': 'baryons'}


def _param_text(d):
    m = str(d['method'])
    diff = 'fourier' if int(d['diff_order']) == 0 else int(d['diff_order'])
    grids = ''.join(f"        '{name}': {{'gravity': {{'{m}': {tuple(int(x) for x in d[f'grids_{name}'])}}}}},\n"
                    for name in d['names'].tolist())
    return f'''
boxsize = {float(d['boxsize'])}*Mpc
potential_options = {{
    'gridsize': {{
        'global': {{'gravity': {{'{m}': {int(d['gridsize_global'])}}}}},
{grids}    }},
    'interpolation': {{'gravity': {{'{m}': '{ORDER_NAME[int(d['order'])]}'}}}},
    'deconvolve': {{'gravity': {{'{m}': ({bool(d['deconvolve'])}, {bool(d['deconvolve'])})}}}},
    'interlace': {{'gravity': {{'{m}': ({bool(d['interlace'])}, {bool(d['interlace'])})}}}},
    'differentiation': {{'default': {{'gravity': {{'pm': {diff!r}, 'p3m': {diff!r}}}}}}},
}}
select_forces = {{'all': {{'gravity': '{m}'}}}}
H0 = 70*km/s/Mpc
Ωcdm = 0.25
Ωb = 0.05
a_begin = 0.02
'''


def _assert_kicks(d, moms):
    for name, mom in zip(d['names'].tolist(), moms):
        kick_ref = d[f'mom_out_{name}'] - d[f'mom_{name}']
        assert np.abs((mom - d[f'mom_{name}']) - kick_ref).max() < 1e-9*np.abs(kick_ref).max(), name


@pytest.mark.parametrize('path', CASES, ids=IDS)
def test_oracle_matches_reference(path):
    from oracle import pm_oracle as O
    d = np.load(path)
    comps = []
    for name in d['names'].tolist():
        up, down = (int(x) for x in d[f'grids_{name}'])
        comps.append(dict(pos=d[f'pos_{name}'], mom=d[f'mom_{name}'], mass=float(d[f'mass_{name}']), upstream=up, downstream=down,
                          dt_rho=float(d[f'dt_rho_{name}']), dt_kick=float(d[f'dt_kick_{name}']), diff_order=int(d['diff_order'])))
    moms = O.pm_kick_multigrid(comps, boxsize=float(d['boxsize']), gridsize_global=int(d['gridsize_global']), order=int(d['order']),
                               G_Newton=float(d['G_Newton']), dt1=float(d['dt1']), deconvolve=bool(d['deconvolve']),
                               interlace=bool(d['interlace']), r_scale=float(d['r_scale']))
    _assert_kicks(d, moms)


def test_oracle_copy_modes_properties():
    """Between equal grids copy_modes is fourier_operate; up- then down-scaling returns the modes inside the small
    grid's Nyquist cube unchanged (the two half-cell phases cancel)."""
    from oracle import pm_oracle as O
    rng = np.random.default_rng(3)
    slab = O.forward_fft(rng.standard_normal((8, 8, 8)))
    same = O.copy_modes(slab, 8, 2, (-.5, -.5, -.5), 2)
    expect = slab*O.deconv_factor(8, 2)*0.5*O.interlace_phase(8, (-.5, -.5, -.5))
    expect[~O.mode_mask(8)] = 0
    assert np.allclose(same, expect, rtol=1e-14, atol=0)
    back = O.copy_modes(O.copy_modes(slab, 12), 8)
    inside = slab.copy()
    inside[~O.mode_mask(8)] = 0
    assert np.allclose(back, inside, rtol=1e-13, atol=1e-13)


def _components(d):
    from concept_b200 import commons
    from concept_b200.species import Component
    comps, ᔑdt = [], {'1': float(d['dt1'])}
    for name in d['names'].tolist():
        c = Component(name, SPECIES[name], N=len(d[f'pos_{name}']), mass=float(d[f'mass_{name}']))
        c.set_particles(d[f'pos_{name}'], d[f'mom_{name}'])
        assert c.potential_gridsizes['gravity'][str(d['method'])] == tuple(int(x) for x in d[f'grids_{name}'])
        ᔑdt['a**(-3*w_eff-1)', name] = float(d[f'dt_rho_{name}'])
        ᔑdt['a**(-3*w_eff)', name] = float(d[f'dt_kick_{name}'])
        comps.append(c)
    commons.universals.a = float(d['a'])
    return comps, ᔑdt


@pytest.mark.parametrize('path', CASES, ids=IDS)
def test_orchestration_through_kernel_model(path, monkeypatch, host_kernels):
    import torch
    from concept_b200 import commons, interactions, mesh
    from concept_b200.species import Component
    import ic_mock_context
    monkeypatch.setattr(ic_mock_context.MeshMockContext, 'lib', host_kernels)
    d = np.load(path)
    commons.load_params(_param_text(d))
    assert commons.shortrange_scale(int(d['gridsize_global'])) == pytest.approx(float(d['r_scale'])) or str(d['method']) == 'pm'
    contexts = {}
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: contexts.setdefault(
        int(gridsize), ic_mock_context.MeshMockContext(gridsize, commons.params.boxsize)))
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    comps, ᔑdt = _components(d)
    interactions.gravity(str(d['method']), comps, comps, ᔑdt, 'long-range', False)
    used = {int(x) for name in d['names'].tolist() for x in d[f'grids_{name}']} | {int(d['gridsize_global'])}
    assert set(contexts) == used
    _assert_kicks(d, [c.mom_local.numpy() for c in comps])


def test_mixed_gridsizes_need_slabs_of_eight_planes(monkeypatch):
    """On several ranks every grid is cut into x-slabs (tests/mgpu_check.py runs the exchange on GPUs): a grid that cannot
    be cut aborts like the reference's slab decomposition (fft.c:105-212 / mesh.py:3779-3783)."""
    from concept_b200 import commons, communication, interactions
    d = np.load(CASES[0])
    commons.load_params(_param_text(d))
    monkeypatch.setattr(communication, 'nprocs', 2)
    with pytest.raises(commons.ConceptAbort):
        interactions._particle_mesh_mixed_gridsizes([], [], [8], [8], 12, 'a²ρ', 'gravity', 'pm', 'gravity', 2, True, True,
                                                    False, {}, ('a**(-3*w_eff)', 'component'))


@pytest.mark.gpu
@pytest.mark.parametrize('path', CASES, ids=IDS)
def test_gpu_mixed_gridsizes_match_reference(path):
    pytest.importorskip('torch')
    from concept_b200 import commons, interactions, mesh
    d = np.load(path)
    commons.load_params(_param_text(d))
    comps, ᔑdt = _components(d)
    interactions.gravity(str(d['method']), comps, comps, ᔑdt, 'long-range', False)
    moms = [c.mom_local.cpu().numpy() for c in comps]
    mesh.free_contexts()
    _assert_kicks(d, moms)


# ---------------------------------------------------------------------- group power spectra, mixed upstream grids
PK_CASES = sorted(glob.glob(os.path.join(HERE, 'golden', 'pkgroup_*.npz')))
PK_IDS = [os.path.basename(p)[:-4] for p in PK_CASES]


def _pk_bins(d):
    return {str(k): float(v) for k, v in zip(d['bins_per_decade_keys'], d['bins_per_decade_vals'])}


@pytest.mark.parametrize('path', PK_CASES, ids=PK_IDS)
def test_group_powerspec_oracle_matches_reference(path):
    """analysis.compute_powerspec of a two-component group whose upstream grids differ from the global one
    (tests/golden/gen_golden_powerspec_group.py)"""
    from concept_b200 import analysis
    from oracle import pm_oracle as O
    d = np.load(path)
    G, L, a = int(d['gridsize']), float(d['boxsize']), float(d['a'])
    comps = [dict(pos=d[f'pos_{n}'], mass=float(d[f'mass_{n}']), w_eff=float(d[f'w_eff_{n}']), upstream=int(d[f'upstream_{n}']))
             for n in d['names'].tolist()]
    slab = O.density_fourier_group(comps, a, L, G, int(d['order']), bool(d['deconvolve']), str(d['interlace']) == 'bcc')
    k2_max, _ = analysis.get_powerspec_bins(G, str(d['k_max']), _pk_bins(d), boxsize=L)
    assert k2_max == int(d['k2_max'])
    power_k2, count_k2 = O.power_by_k2(slab, k2_max)
    _, idx, centers, n_modes = analysis.get_powerspec_bins(G, str(d['k_max']), _pk_bins(d), count_k2, boxsize=L)
    assert np.array_equal(n_modes, d['n_modes']) and np.allclose(centers, d['k_bin_centers'], rtol=1e-13, atol=0)
    power = np.zeros(len(centers))
    np.add.at(power, idx, power_k2)
    rho_bar = sum(a**(-3*(1 + float(d[f'w_eff_{n}'])))*float(d[f'varrho_bar_{n}']) for n in d['names'].tolist())
    power *= rho_bar**(-2)*L**3/n_modes
    assert np.abs(power/d['power'] - 1).max() < 1e-11


def _pk_components(d):
    from concept_b200 import commons
    from concept_b200.species import Component
    commons.load_params(f"boxsize = {float(d['boxsize'])}*Mpc\nH0 = 70*km/s/Mpc\nΩcdm = 0.25\nΩb = 0.05\n")
    commons.universals.a = float(d['a'])
    comps = []
    for name in d['names'].tolist():
        c = Component(name, SPECIES[name], N=len(d[f'pos_{name}']), mass=float(d[f'mass_{name}']))
        c.set_particles(d[f'pos_{name}'], np.zeros_like(d[f'pos_{name}']))
        assert abs(c.ϱ_bar/float(d[f'varrho_bar_{name}']) - 1) < 1e-12
        comps.append(c)
    return comps


def _pk_product(d, comps):
    from concept_b200 import analysis
    return analysis.powerspec(comps, int(d['gridsize']), ORDER_NAME[int(d['order'])], bool(d['deconvolve']),
                              str(d['interlace']) == 'bcc', str(d['k_max']), _pk_bins(d),
                              gridsizes_upstream=[int(d[f'upstream_{n}']) for n in d['names'].tolist()])


@pytest.mark.parametrize('path', PK_CASES, ids=PK_IDS)
def test_group_powerspec_orchestration_through_kernel_model(path, monkeypatch, host_kernels):
    import torch
    from concept_b200 import commons, mesh
    from concept_b200.species import Component
    import ic_mock_context
    monkeypatch.setattr(ic_mock_context.MeshMockContext, 'lib', host_kernels)
    d = np.load(path)
    contexts = {}
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    comps = _pk_components(d)
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: contexts.setdefault(
        int(gridsize), ic_mock_context.MeshMockContext(gridsize, commons.params.boxsize)))
    centers, power, n_modes = _pk_product(d, comps)
    assert set(contexts) == {int(d['gridsize'])} | {int(d[f'upstream_{n}']) for n in d['names'].tolist()}
    assert np.array_equal(n_modes, d['n_modes'])
    assert np.abs(power/d['power'] - 1).max() < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize('path', PK_CASES, ids=PK_IDS)
def test_gpu_group_powerspec_matches_reference(path):
    pytest.importorskip('torch')
    from concept_b200 import mesh
    d = np.load(path)
    comps = _pk_components(d)
    centers, power, n_modes = _pk_product(d, comps)
    mesh.free_contexts()
    assert np.array_equal(n_modes, d['n_modes'])
    assert np.allclose(centers, d['k_bin_centers'], rtol=1e-13, atol=0)
    assert np.abs(power/d['power'] - 1).max() < 1e-9      # the north star asks for 1e-4; measured on the CPU model: 1e-15
