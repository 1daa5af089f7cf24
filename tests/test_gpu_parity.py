"""GPU parity tests: libpmgrav (through the C ABI) vs the golden vectors produced by the reference
itself (tests/golden) and vs the numpy oracle on seeded inputs.

Tolerances (fp64 grid).  Bit parity is impossible by construction: the reference is built with
-ffast-math and lets FFTW pick plans by timing; its own cross-build / cross-nprocs tolerance is 1e-9
(test/optimizations/analyze.py:18-22, test/nprocs_pm/analyze.py:121).  Here:
  grids (density, potential):  1e-12 relative to the grid maximum (summation order of atomics / cuFFT);
  kick Δmom:                   KICK_RTOL = 1e-9 relative to max|Δmom| — the force is a difference of
                               neighbouring potential values, so FFT rounding (1e-16·|φ|) is amplified
                               by |φ|/|Δφ| ~ 1e4-1e5; measured errors are 1e-13 … 3e-11;
  drift:                       bit exact.
"""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import pm_oracle as O  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
KICKS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'kick_*.npz')))
KICK_RTOL = 1e-9


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))/np.max(np.abs(b)))


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')


def golden_params(d):
    from concept_b200.pmsolver import make_kick_params
    return make_kick_params(
        mass=float(d['mass']), boxsize=float(d['boxsize']), gridsize=int(d['gridsize']), order=int(d['order']),
        G_Newton=float(d['G_Newton']), dt_rho_over_dt1=float(d['dt_rho'])/float(d['dt_1']), dt_kick=float(d['dt_kick']),
        diff_order=int(d['diff_order']), deconvolve=bool(d['deconvolve']), interlace=bool(d['interlace']),
        r_scale=float(d['r_scale']) if 'r_scale' in d.files else 0.0)


@pytest.mark.parametrize('name', KICKS)
def test_kick_long_matches_reference_golden(name):
    from concept_b200.pmsolver import PMContext
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    ctx = PMContext(int(d['gridsize']), float(d['boxsize']))
    pos, mom = dev(d['pos']), dev(d['mom'])
    s = torch.zeros(1, dtype=torch.float64, device='cuda')
    ctx.kick_long(pos, mom, golden_params(d), sum_mom2=s)
    torch.cuda.synchronize()
    dmom = mom.cpu().numpy() - d['mom']
    dmom_ref = d['mom_out'] - d['mom']
    assert relerr(dmom, dmom_ref) < KICK_RTOL, name
    assert np.array_equal(pos.cpu().numpy(), d['pos'])
    assert abs(s.item() - O.sum_mom2(d['mom_out'])) < 1e-12*O.sum_mom2(d['mom_out'])
    ctx.close()


@pytest.mark.parametrize('name', ['kick_pm_cic_G8_d2', 'kick_pm_tsc_G12_d4', 'kick_pm_pcs_G10_d6', 'kick_pm_ngp_G8_d2',
                                  'kick_pm_cic_G8_edges', 'kick_p3m_long_tsc_G24'])
def test_stages_match_reference_taps(name):
    """deposit → density, solve → potential, diff → force grids, each against the reference's taps."""
    from concept_b200.pmsolver import PMContext, PM_TAP_FORCE, PM_TAP_REAL
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    p = golden_params(d)
    ctx = PMContext(int(d['gridsize']), float(d['boxsize']))
    pos = dev(d['pos'])
    ctx.grid_zero()
    ctx.deposit(pos, p.order, p.contribution)
    rho = ctx.get_grid(PM_TAP_REAL)
    assert relerr(rho, d['tap_rho']) < 1e-13
    ctx.fft_forward()
    ctx.kspace_potential(p.prefactor, p.deconv_order, p.gauss, 1.0)
    ctx.fft_backward()
    phi = ctx.get_grid(PM_TAP_REAL)
    assert relerr(phi, d['tap_phi']) < 1e-12
    for dim in range(3):
        ctx.diff(dim, p.diff_order)
        f = ctx.get_grid(PM_TAP_FORCE)
        assert relerr(f, d[f'tap_forcegrid{dim}']) < 1e-11
        # un-fused gather from the explicit force grid == reference's per-dimension application
        mom = dev(d['mom'])
        ctx.gather(PM_TAP_FORCE, pos, mom, p.order, dim, p.kick_factor)
        got = mom.cpu().numpy()[:, dim] - d['mom'][:, dim]
        ref = d['mom_out'][:, dim] - d['mom'][:, dim]
        assert relerr(got, ref) < KICK_RTOL
    ctx.close()


def test_fourier_slab_matches_oracle():
    from concept_b200.pmsolver import PMContext, PM_TAP_FOURIER
    G, L = 16, 10.0
    rng = np.random.default_rng(5)
    rho = rng.standard_normal((G, G, G))
    ctx = PMContext(G, L)
    ctx.set_grid(rho)
    ctx.fft_forward()
    f = ctx.get_grid(PM_TAP_FOURIER)
    ref = O.forward_fft(rho)
    assert relerr(f, ref) < 1e-13
    ctx.kspace_potential(-L**2*4.5e-5/np.pi, 4, 0.0, 1.0)
    f2 = ctx.get_grid(PM_TAP_FOURIER)
    ref2 = ref*O.potential_factor(G, L, 4.5e-5, 4)
    assert relerr(f2, ref2) < 1e-13
    assert np.all(f2[G//2] == 0) and np.all(f2[:, G//2] == 0) and np.all(f2[:, :, G//2] == 0) and f2[0, 0, 0] == 0
    ctx.close()


def test_drift_bit_exact_vs_reference_golden():
    from concept_b200.pmsolver import PMContext
    d = np.load(os.path.join(GOLDEN, 'drift_G8.npz'))
    ctx = PMContext(8, float(d['boxsize']))
    pos, mom = dev(d['pos']), dev(d['mom'])
    ctx.drift(pos, mom, float(d['dt_am2'])/float(d['mass']))
    out = pos.cpu().numpy()
    assert np.array_equal(out, d['pos_out'])      # integer-like bar: bit exact
    ctx.close()


@pytest.mark.parametrize('order,diff_order,interlace', [(2, 2, False), (3, 4, False), (4, 2, False), (2, 0, False), (3, 2, True)])
def test_kick_vs_oracle_seeded_32(order, diff_order, interlace):
    """Seeded random + clustered particles at a size the numpy oracle finishes in seconds."""
    from concept_b200.pmsolver import PMContext, make_kick_params
    G, L, N = 32, 100.0, 20000
    rng = np.random.default_rng(1234 + order)
    pos = rng.random((N, 3))*L
    pos[:5000] = (0.5 + 0.01*rng.standard_normal((5000, 3)))*L % L     # a clump: heavy atomic contention
    mom = rng.standard_normal((N, 3))
    kw = dict(mass=2.5, boxsize=L, gridsize=G, order=order, G_Newton=4.4985024439973154e-05,
              dt_rho_over_dt1=1.9, dt_kick=0.011, diff_order=diff_order, interlace=interlace)
    ref = O.pm_kick(pos, mom, **kw)
    ctx = PMContext(G, L)
    dpos, dmom = dev(pos), dev(mom)
    ctx.kick_long(dpos, dmom, make_kick_params(**kw))
    got = dmom.cpu().numpy()
    assert relerr(got - mom, ref - mom) < KICK_RTOL
    ctx.close()


def test_kick_long_host_entry_point():
    from concept_b200.pmsolver import PMContext, make_kick_params
    d = np.load(os.path.join(GOLDEN, 'kick_pm_cic_G8_d2.npz'))
    ctx = PMContext(int(d['gridsize']), float(d['boxsize']))
    pos, mom = d['pos'].copy(), d['mom'].copy()
    s = ctx.kick_long_host(pos, mom, golden_params(d), dt_over_mass=0.0, want_sum=True)
    assert relerr(mom - d['mom'], d['mom_out'] - d['mom']) < KICK_RTOL
    assert abs(s - O.sum_mom2(d['mom_out'])) < 1e-12*s
    # with a drift folded in
    pos2, mom2 = d['pos'].copy(), d['mom'].copy()
    ctx.kick_long_host(pos2, mom2, golden_params(d), dt_over_mass=0.37)
    assert np.array_equal(pos2, O.drift(d['pos'], mom2, 0.37, float(d['boxsize'])))
    ctx.close()


def test_f32_grid_within_stated_tolerance():
    """Config 3 (fp32 grid/FFT, fp64 particles) has no reference counterpart;
    stated tolerance: rms force error ≤ 1e-5 of the rms force per kick."""
    from concept_b200.pmsolver import PMContext, make_kick_params
    G, L, N = 32, 100.0, 30000
    rng = np.random.default_rng(7)
    pos = rng.random((N, 3))*L
    mom = np.zeros((N, 3))
    kw = dict(mass=2.5, boxsize=L, gridsize=G, order=3, G_Newton=4.4985024439973154e-05,
              dt_rho_over_dt1=1.9, dt_kick=0.011, diff_order=2)
    ref = O.pm_kick(pos, mom, **kw)
    ctx = PMContext(G, L, dtype='f32')
    dpos, dmom = dev(pos), dev(mom)
    ctx.kick_long(dpos, dmom, make_kick_params(**kw))
    got = dmom.cpu().numpy()
    rms = np.sqrt(np.mean((got - ref)**2))/np.sqrt(np.mean(ref**2))
    assert rms < 1e-5
    ctx.close()


def test_errors_are_loud():
    from concept_b200 import _lib
    from concept_b200.pmsolver import PMContext
    with pytest.raises(_lib.PMError):
        PMContext(7, 1.0)          # odd grid size
    ctx = PMContext(8, 1.0)
    pos = torch.zeros((4, 3), dtype=torch.float64, device='cuda')
    with pytest.raises(_lib.PMError):
        ctx.deposit(pos, 5, 1.0)   # order ∉ {1,2,3,4}: the reference abort()s (mesh.py:1533)
    with pytest.raises(_lib.PMError):
        ctx.fft_backward()         # real-space data in the slab
    ctx.close()


def test_empty_component_is_a_noop():
    from concept_b200.pmsolver import PMContext, make_kick_params
    ctx = PMContext(8, 8.0)
    pos = torch.zeros((0, 3), dtype=torch.float64, device='cuda')
    mom = torch.zeros((0, 3), dtype=torch.float64, device='cuda')
    p = make_kick_params(mass=1.0, boxsize=8.0, gridsize=8, order=2, G_Newton=1.0, dt_rho_over_dt1=1.0, dt_kick=1.0)
    ctx.kick_long(pos, mom, p)
    ctx.drift(pos, mom, 1.0)
    assert np.all(ctx.get_grid() == 0)
    ctx.close()


@pytest.mark.parametrize('dtype,tol', [('f64', 1e-12), ('f32', 2e-5)])
def test_fused_solve_matches_three_call_path_and_oracle_G64(dtype, tol):
    """pm_solve_fused (2-D cuFFT + fused x pass: FFT · Green's function · inverse FFT) ==
    pm_fft_forward + pm_kspace_potential + pm_fft_backward == oracle, on a random density."""
    from concept_b200.pmsolver import PMContext
    G, L = 64, 100.0
    rng = np.random.default_rng(42)
    rho = rng.standard_normal((G, G, G))
    pref, D, gauss = -L**2*4.4985e-5/np.pi, 4, (2*np.pi/L*1.25*L/G)**2
    ctx = PMContext(G, L, dtype=dtype)
    assert ctx.fused_solve_available
    for g in (0.0, gauss):
        ctx.set_grid(rho)
        ctx.fft_forward(); ctx.kspace_potential(pref, D, g, 1.0); ctx.fft_backward()
        phi_a = ctx.get_grid()
        ctx.set_grid(rho)
        ctx.solve_fused(pref, D, g)
        phi_b = ctx.get_grid()
        ref = O.backward_fft(O.forward_fft(rho)*O.potential_factor(G, L, 4.4985e-5, D, r_scale=(1.25*L/G if g else 0.0)), G)
        assert relerr(phi_b, ref) < tol
        assert relerr(phi_b, phi_a) < tol
    ctx.close()


def test_fused_solve_G512_kick_matches_three_call_path():
    """At the benchmark grid (512³) the fused path must give the same kick as the cuFFT-3D + k-space path."""
    from concept_b200.pmsolver import PMContext, make_kick_params
    G, L, N = 512, 512.0, 200000
    rng = np.random.default_rng(9)
    pos = rng.random((N, 3))*L
    mom = np.zeros((N, 3))
    kw = dict(mass=1.0, boxsize=L, gridsize=G, order=2, G_Newton=4.4985024439973154e-05, dt_rho_over_dt1=2.0, dt_kick=1e-3)
    ctx = PMContext(G, L)
    assert ctx.fused_solve_available
    out = []
    for fused in (True, False):
        ctx.set_fused_solve(fused)
        dpos, dmom = dev(pos), dev(mom)
        ctx.kick_long(dpos, dmom, make_kick_params(**kw))
        out.append(dmom.cpu().numpy())
    assert relerr(out[0], out[1]) < KICK_RTOL
    assert np.max(np.abs(out[0])) > 0
    ctx.close()


def test_sort_particles_by_cell():
    """pm_sort_particles: stable reorder by (x plane, y row, z); ids travel with the particles and the kick
    is the same particle by particle."""
    import torch
    from concept_b200.pmsolver import PMContext, make_kick_params
    G, L, N = 64, 100.0, 200_003
    rng = np.random.default_rng(9)
    pos_h, mom_h = rng.random((N, 3))*L, rng.standard_normal((N, 3))
    ctx = PMContext(G, L)
    pos, mom = torch.as_tensor(pos_h, device='cuda'), torch.as_tensor(mom_h, device='cuda')
    ids = torch.arange(N, dtype=torch.int64, device='cuda')
    ctx.sort_particles(pos, mom, ids)
    p, m, i = pos.cpu().numpy(), mom.cpu().numpy(), ids.cpu().numpy()
    assert np.array_equal(np.sort(i), np.arange(N))
    assert np.array_equal(p, pos_h[i]) and np.array_equal(m, mom_h[i])
    cell = np.minimum((p*(G/L)).astype(np.int64), G - 1)
    key = (cell[:, 0]*G + cell[:, 1])*G + cell[:, 2]
    assert np.all(np.diff(key) >= 0)
    same = np.diff(key) == 0
    assert np.all(np.diff(i)[same] > 0)            # stable within a cell
    # the kick does not care about the order
    kw = dict(mass=1.0, boxsize=L, gridsize=G, order=2, G_Newton=4.4985024439973154e-05, dt_rho_over_dt1=2.0, dt_kick=0.01)
    ctx.kick_long(pos, mom, make_kick_params(**kw))
    pos2, mom2 = torch.as_tensor(pos_h, device='cuda'), torch.as_tensor(mom_h, device='cuda')
    ctx.kick_long(pos2, mom2, make_kick_params(**kw))
    d = mom2.cpu().numpy() - mom_h
    # (+ a few ulp of |mom| ~ 1: the kick itself is only ~1e-6 here)
    assert np.max(np.abs((mom.cpu().numpy() - m) - d[i])) < 1e-11*np.max(np.abs(d)) + 2e-15
    ctx.close()
