"""GPU parity of the hand-written slab transform (pm_fft.cu) behind pm_solve_fused / pm_kick_long:
against the numpy oracle (G = 128), against the cuFFT-based paths of the same library (G = 128, 256,
512), run-to-run bit stability of the dependency-ordered (L2-resident) schedule, and the whole kick."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

G_NEWTON = 4.4985024439973154e-05


def _ctx(G, L, dtype='f64'):
    from concept_b200.pmsolver import PMContext
    return PMContext(G, L, dtype=dtype)


def _solve(ctx, rho, mode, pre, deconv, gauss=0.0):
    ctx.set_fused_solve(mode)
    ctx.set_grid(rho)
    if mode == 'unfused':
        ctx.fft_forward()
        ctx.kspace_potential(pre, deconv, gauss, 1.0)
        ctx.fft_backward()
    else:
        ctx.solve_fused(pre, deconv, gauss)
        ctx.check_async_error()
    return ctx.get_grid()


def _numpy_solve(rho, pre, deconv, gauss):
    G = rho.shape[0]
    l = np.arange(G)
    k = np.where(l >= G//2, l - G, l).astype(float)
    x = k*(np.pi/G) + 2.220446049250313e-16
    d = x/np.sin(x)
    kz = k[:G//2 + 1].copy(); kz[-1] = G//2
    dz = d[:G//2 + 1]
    k2 = k[:, None, None]**2 + k[None, :, None]**2 + kz[None, None, :]**2
    fac = ((d[:, None, None]*d[None, :, None])*dz[None, None, :])**deconv*pre/np.where(k2 == 0, 1, k2)*np.exp(-gauss*k2)
    fac[k2 == 0] = 0
    fac[G//2, :, :] = 0; fac[:, G//2, :] = 0; fac[:, :, G//2] = 0
    return np.fft.irfftn(np.fft.rfftn(rho)*fac, s=rho.shape, axes=(0, 1, 2), norm='forward')


@pytest.mark.parametrize('mode', ['fft2_l2', 'fft2_split'])
@pytest.mark.parametrize('gauss', [0.0, 3e-3])
def test_solve_G128_against_numpy(mode, gauss):
    G, L = 128, 100.0
    rho = np.random.default_rng(1).standard_normal((G, G, G))
    ctx = _ctx(G, L)
    got = _solve(ctx, rho, mode, -2.5, 4, gauss)
    ref = _numpy_solve(rho, -2.5, 4, gauss)
    ctx.close()
    assert np.max(np.abs(got - ref))/np.max(np.abs(ref)) < 1e-12


@pytest.mark.parametrize('G', [128, 256, 512])
def test_solve_matches_cufft_paths(G):
    L = 300.0
    rng = np.random.default_rng(G)
    rho = rng.standard_normal((G, G, G))
    ctx = _ctx(G, L)
    ref = _solve(ctx, rho, 'unfused', -1.7, 6)
    scale = np.max(np.abs(ref))
    for mode in ('fft2_l2', 'fft2_split'):
        got = _solve(ctx, rho, mode, -1.7, 6)
        assert np.max(np.abs(got - ref))/scale < 1e-12, mode
    if G == 512:
        got = _solve(ctx, rho, 'cufft2d', -1.7, 6)
        assert np.max(np.abs(got - ref))/scale < 1e-12
    ctx.close()


def test_solve_G1024_matches_cufft_path():
    """BASELINE.json configs[3]: the 1024³ grid (radix-16 first stage, 64 KB tiles) — the potential of 2·10⁶ random
    particles through the hand-written transforms against the cuFFT 3-D path of the same library."""
    from concept_b200.pmsolver import make_kick_params
    G, L, N = 1024, 1024.0, 2_000_000
    pos = torch.rand((N, 3), dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(5))*L
    p = make_kick_params(mass=1.0, boxsize=L, gridsize=G, order=2, G_Newton=G_NEWTON, dt_rho_over_dt1=2.0, dt_kick=1e-3, diff_order=2)
    ctx = _ctx(G, L)
    assert ctx.hand_fft_available
    ctx.grid_zero(); ctx.deposit(pos, 2, p.contribution)
    ctx.solve_fused(p.prefactor, p.deconv_order, p.gauss)
    ctx.check_async_error()
    got = torch.as_tensor(ctx.get_grid())
    ctx.grid_zero(); ctx.deposit(pos, 2, p.contribution)
    ctx.fft_forward(); ctx.kspace_potential(p.prefactor, p.deconv_order, p.gauss, 1.0); ctx.fft_backward()
    ref = torch.as_tensor(ctx.get_grid())
    ctx.close()
    assert float((got - ref).abs().max()/ref.abs().max()) < 1e-12


def test_l2_schedule_is_bit_stable():
    """The dependency-ordered schedule must not depend on timing: 12 runs, identical bits."""
    G, L = 512, 512.0
    rho = np.random.default_rng(3).standard_normal((G, G, G))
    ctx = _ctx(G, L)
    first = _solve(ctx, rho, 'fft2_l2', -1.0, 4)
    split = _solve(ctx, rho, 'fft2_split', -1.0, 4)
    assert np.array_equal(first, split)
    for _ in range(11):
        assert np.array_equal(_solve(ctx, rho, 'fft2_l2', -1.0, 4), first)
    ctx.close()


@pytest.mark.parametrize('G', [128, 512])
def test_solve_fp32_grid(G):
    L = 200.0
    rho = np.random.default_rng(7).standard_normal((G, G, G))
    c64 = _ctx(G, L)
    ref = _solve(c64, rho, 'unfused', -1.3, 6)
    c64.close()
    c32 = _ctx(G, L, 'f32')
    got = _solve(c32, rho, 'auto', -1.3, 6)
    got_cufft = _solve(c32, rho, 'unfused', -1.3, 6)
    c32.close()
    scale = np.max(np.abs(ref))
    e_new, e_cufft = np.max(np.abs(got - ref))/scale, np.max(np.abs(got_cufft - ref))/scale
    assert e_new < 2e-5, e_new
    assert e_new < 4*e_cufft + 1e-6, (e_new, e_cufft)    # as accurate as the library transform in fp32


@pytest.mark.parametrize('G,order,diff', [(128, 2, 2), (256, 3, 4)])
def test_kick_against_oracle(G, order, diff):
    """Whole long-range kick through the hand-written transforms vs the numpy oracle."""
    from concept_b200.pmsolver import make_kick_params
    from oracle import pm_oracle as O
    L, N = 256.0, 200_000
    rng = np.random.default_rng(11)
    pos_h, mom_h = rng.random((N, 3))*L, rng.standard_normal((N, 3))
    kw = dict(mass=2.1, boxsize=L, gridsize=G, order=order, G_Newton=G_NEWTON, dt_rho_over_dt1=2.0, dt_kick=0.01, diff_order=diff)
    ref = O.pm_kick(pos_h, mom_h, **kw) - mom_h
    ctx = _ctx(G, L)
    assert ctx.fused_solve_available
    ctx.set_fused_solve('fft2_l2')
    pos, mom = torch.as_tensor(pos_h, device='cuda'), torch.as_tensor(mom_h, device='cuda')
    ctx.kick_long(pos, mom, make_kick_params(**kw))
    ctx.check_async_error()
    got = mom.cpu().numpy() - mom_h
    ctx.close()
    # tolerance: tests/test_gpu_parity.py (Δmom 1e-9 of max|Δmom|)
    assert np.max(np.abs(got - ref))/np.max(np.abs(ref)) < 1e-9


@pytest.mark.parametrize('order,diff,dtype', [(2, 2, 'f64'), (3, 4, 'f64'), (4, 8, 'f64'), (1, 1, 'f64'), (3, 2, 'f32')])
def test_kick_drift_equals_kick_then_drift(order, diff, dtype):
    """pm_kick_drift (drift fused into the gather/kick kernel) == pm_kick_long followed by pm_drift,
    and the deduplicated gather matches the numpy oracle."""
    from concept_b200.pmsolver import make_kick_params
    from oracle import pm_oracle as O
    G, L, N = 64, 128.0, 50_000
    rng = np.random.default_rng(order*10 + diff)
    pos_h, mom_h = rng.random((N, 3))*L, rng.standard_normal((N, 3))
    pos_h[:64] = np.floor(pos_h[:64]/(L/G))*(L/G)          # particles exactly on cell edges
    pos_h[64] = np.nextafter(L, 0)
    kw = dict(mass=0.7, boxsize=L, gridsize=G, order=order, G_Newton=G_NEWTON, dt_rho_over_dt1=1.5, dt_kick=0.02, diff_order=diff)
    params = make_kick_params(**kw)
    dtm = 0.37
    ctx = _ctx(G, L, dtype)
    p1, m1 = torch.as_tensor(pos_h, device='cuda'), torch.as_tensor(mom_h, device='cuda')
    s1 = torch.zeros(1, dtype=torch.float64, device='cuda')
    ctx.kick_long(p1, m1, params, sum_mom2=s1)
    ctx.drift(p1, m1, dtm)
    p2, m2 = torch.as_tensor(pos_h, device='cuda'), torch.as_tensor(mom_h, device='cuda')
    s2 = torch.zeros(1, dtype=torch.float64, device='cuda')
    ctx.kick_drift(p2, m2, params, dtm, sum_mom2=s2)
    ctx.close()
    dm = (m1 - torch.as_tensor(mom_h, device='cuda')).abs().max().item()
    # the deposit's atomics make the two potentials differ in the last bits
    # (+ a few ulp of |mom| ~ 1: the kick itself is only ~1e-6 here)
    assert (m1 - m2).abs().max().item() < (1e-11 if dtype == 'f64' else 1e-5)*dm + 2e-15
    assert torch.equal(p2, torch.as_tensor(O.drift(pos_h, m2.cpu().numpy(), dtm, L), device='cuda'))
    assert abs(s1.item() - s2.item()) < 1e-9*abs(s1.item())
    if dtype == 'f64':
        ref = O.pm_kick(pos_h, mom_h, **kw) - mom_h
        got = m2.cpu().numpy() - mom_h
        assert np.max(np.abs(got - ref))/np.max(np.abs(ref)) < 1e-9


def test_kick_long_host_pipelined_matches_device_path():
    """pm_kick_long_host with enough particles runs the chunked H2D / deposit / gather / D2H pipeline;
    it must give what the device-resident pm_kick_drift gives."""
    from concept_b200.pmsolver import make_kick_params
    from oracle import pm_oracle as O
    G, L, N = 128, 200.0, 700_001
    rng = np.random.default_rng(21)
    pos_h, mom_h = rng.random((N, 3))*L, rng.standard_normal((N, 3))
    kw = dict(mass=1.1, boxsize=L, gridsize=G, order=2, G_Newton=G_NEWTON, dt_rho_over_dt1=1.5, dt_kick=0.02, diff_order=2)
    params = make_kick_params(**kw)
    dtm = 0.21
    ctx = _ctx(G, L)
    p1, m1 = torch.as_tensor(pos_h, device='cuda'), torch.as_tensor(mom_h, device='cuda')
    s1 = torch.zeros(1, dtype=torch.float64, device='cuda')
    ctx.kick_drift(p1, m1, params, dtm, sum_mom2=s1)
    ph, mh = pos_h.copy(), mom_h.copy()
    for _ in range(2):      # twice: staging buffers and events are reused
        ph[:], mh[:] = pos_h, mom_h
        s2 = ctx.kick_long_host(ph, mh, params, dt_over_mass=dtm, want_sum=True)
    ctx.check_async_error()
    ctx.close()
    dm = np.max(np.abs(m1.cpu().numpy() - mom_h))
    assert np.max(np.abs(mh - m1.cpu().numpy())) < 1e-11*dm + 2e-15
    assert np.array_equal(ph, O.drift(pos_h, mh, dtm, L))
    assert abs(s2 - s1.item()) < 1e-9*abs(s2)


def test_empty_and_single_particle_on_the_hand_written_path():
    """Edge cases through pm_kick_drift at G = 128: no particles (a no-op that still leaves a clean grid),
    and a single particle (its self-force through the mesh, against the oracle)."""
    from concept_b200.pmsolver import make_kick_params
    from oracle import pm_oracle as O
    G, L = 128, 64.0
    kw = dict(mass=1.0, boxsize=L, gridsize=G, order=2, G_Newton=G_NEWTON, dt_rho_over_dt1=1.0, dt_kick=0.5, diff_order=2)
    params = make_kick_params(**kw)
    ctx = _ctx(G, L)
    pos = torch.zeros((0, 3), dtype=torch.float64, device='cuda')
    mom = torch.zeros((0, 3), dtype=torch.float64, device='cuda')
    ctx.kick_drift(pos, mom, params, 0.1)
    ctx.sort_particles(pos, mom)
    ctx.check_async_error()
    assert np.all(ctx.get_grid() == 0)
    pos_h = np.array([[L - 1e-9, 0.3*L, 17.123]])
    mom_h = np.array([[0.1, -0.2, 0.3]])
    pos, mom = torch.as_tensor(pos_h, device='cuda'), torch.as_tensor(mom_h, device='cuda')
    ctx.kick_drift(pos, mom, params, 0.1)
    ctx.check_async_error()
    ref = O.pm_kick(pos_h, mom_h, **kw)
    scale = max(np.max(np.abs(ref - mom_h)), 1e-30)
    assert np.max(np.abs(mom.cpu().numpy() - ref)) < 1e-9*scale + 1e-15
    assert np.array_equal(pos.cpu().numpy(), O.drift(pos_h, mom.cpu().numpy(), 0.1, L))
    ctx.close()
