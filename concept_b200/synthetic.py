"""Synthetic particle loads for benchmarks and large parity tests:
a simple-cubic lattice at cell centres plus a Zel'dovich displacement drawn from a Gaussian
random field with P(k) ∝ k^n_s·T²_BBKS(k), scaled to a chosen rms displacement (in units of the
inter-particle spacing), and momenta proportional to the displacement.

Generated with torch (CPU or CUDA) from a fixed seed — the same arrays feed the GPU path, the
oracle and the CPU baseline.  This is input generation, not part of the timed hot path.
"""
import math

import torch


def _bbks(k, gamma=0.2):
    q = k/gamma
    q = torch.clamp(q, min=1e-12)
    return torch.log(1 + 2.34*q)/(2.34*q)*(1 + 3.89*q + (16.1*q)**2 + (5.46*q)**3 + (6.71*q)**4)**(-0.25)


def zeldovich_particles(n_side, boxsize, sigma_spacing=0.3, seed=0, n_s=0.96, device='cpu', mass=1.0, vel_factor=1.0):
    """Returns pos (N,3), mom (N,3) float64 on `device`, N = n_side³."""
    n = int(n_side)
    gen = torch.Generator(device='cpu')
    gen.manual_seed(int(seed))
    dev = torch.device(device)
    # white noise on CPU for reproducibility across devices, then move
    noise = torch.randn((n, n, n), generator=gen, dtype=torch.float64).to(dev)
    dk = torch.fft.rfftn(noise)
    k1 = torch.fft.fftfreq(n, d=1.0/n, device=dev, dtype=torch.float64)*(2*math.pi/boxsize)
    kz = torch.fft.rfftfreq(n, d=1.0/n, device=dev, dtype=torch.float64)*(2*math.pi/boxsize)
    kx, ky, kzz = k1[:, None, None], k1[None, :, None], kz[None, None, :]
    k2 = kx**2 + ky**2 + kzz**2
    k2[0, 0, 0] = 1.0
    k = torch.sqrt(k2)
    amp = torch.sqrt(k**n_s*_bbks(k)**2)
    amp[0, 0, 0] = 0.0
    dk = dk*amp
    spacing = boxsize/n
    psi = []
    for kd in (kx, ky, kzz):
        comp = torch.fft.irfftn(1j*kd/k2*dk, s=(n, n, n))
        psi.append(comp)
    psi = torch.stack(psi, dim=-1).reshape(-1, 3)
    rms = torch.sqrt((psi**2).sum(dim=1).mean()/3)
    psi = psi*(sigma_spacing*spacing/rms)
    idx = (torch.arange(n, device=dev, dtype=torch.float64) + 0.5)*spacing
    q = torch.stack(torch.meshgrid(idx, idx, idx, indexing='ij'), dim=-1).reshape(-1, 3)
    pos = torch.remainder(q + psi, boxsize)
    pos = torch.where(pos >= boxsize, torch.zeros_like(pos), pos)
    mom = psi*(mass*vel_factor)
    return pos.contiguous(), mom.contiguous()
