"""Size-independent properties of the hot path at the benchmark's full size (BASELINE.json configs[1]: 256³
particles on a 512³ grid, CIC, fp64, deconvolution order 4, finite-difference order 2), where the oracle is too
slow to be the checker:

  * momentum conservation — deposit and gather use the same window, the Green's function is even and the centred
    difference odd, so pair forces are antisymmetric and Σ Δmom vanishes to rounding;
  * translation covariance — moving every particle by a whole number of cells (periodically) moves the kicks with it;
  * the fused kick+drift equals kick followed by drift, and drifting forth and back restores the positions.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

G, N_SIDE, L = 512, 256, 512.0
G_NEWTON = 4.4985024439973154e-05


def _kick_params(**kw):
    from concept_b200.pmsolver import make_kick_params
    return make_kick_params(mass=1.3, boxsize=L, gridsize=G, order=2, G_Newton=G_NEWTON, dt_rho_over_dt1=2.0, dt_kick=1e-3,
                            diff_order=2, **kw)


def test_full_size_momentum_conservation_and_translation():
    from concept_b200.pmsolver import PMContext
    from concept_b200.synthetic import zeldovich_particles
    pos, _ = zeldovich_particles(N_SIDE, L, sigma_spacing=0.3, seed=0, device='cuda')
    N = pos.shape[0]
    assert N == N_SIDE**3
    ctx = PMContext(G, L)
    assert ctx.hand_fft_available
    mom = torch.zeros_like(pos)
    ctx.kick_long(pos, mom, _kick_params())
    total = mom.sum(dim=0).cpu().numpy()
    scale = float(mom.abs().sum().item())/3
    assert scale > 0
    assert np.abs(total).max() < 1e-9*scale, (total, scale)
    # translate by (3, −5, 7) cells: cell size is exactly 1.0, so only the last bits of small coordinates change
    shift = torch.tensor([3.0, -5.0, 7.0], dtype=torch.float64, device='cuda')
    pos_shifted = torch.remainder(pos + shift, L)
    pos_shifted = torch.where(pos_shifted >= L, torch.zeros_like(pos_shifted), pos_shifted).contiguous()
    mom_shifted = torch.zeros_like(pos)
    ctx.kick_long(pos_shifted, mom_shifted, _kick_params())
    err = float((mom_shifted - mom).abs().max().item())/float(mom.abs().max().item())
    assert err < 1e-8, err
    ctx.close()


def test_full_size_kick_drift_fusion_and_drift_round_trip():
    from concept_b200.pmsolver import PMContext
    from concept_b200.synthetic import zeldovich_particles
    pos0, mom0 = zeldovich_particles(N_SIDE, L, sigma_spacing=0.3, seed=1, device='cuda', mass=1.3, vel_factor=0.05)
    ctx = PMContext(G, L)
    dt_over_mass = 0.37
    # fused
    pos_a, mom_a = pos0.clone(), mom0.clone()
    ctx.kick_drift(pos_a, mom_a, _kick_params(), dt_over_mass)
    # kick, then drift
    pos_b, mom_b = pos0.clone(), mom0.clone()
    ctx.kick_long(pos_b, mom_b, _kick_params())
    ctx.drift(pos_b, mom_b, dt_over_mass)
    assert float((mom_a - mom_b).abs().max().item()) <= 1e-9*float((mom_b - mom0).abs().max().item())
    d = (pos_a - pos_b).abs()
    d = torch.minimum(d, L - d)
    assert float(d.max().item()) < 1e-9
    # forth and back
    pos_c = pos0.clone()
    ctx.drift(pos_c, mom0, dt_over_mass)
    ctx.drift(pos_c, mom0, -dt_over_mass)
    d = (pos_c - pos0).abs()
    d = torch.minimum(d, L - d)
    assert float(d.max().item()) < 1e-12*L
    assert float(pos_c.min().item()) >= 0 and float(pos_c.max().item()) < L
    ctx.close()
