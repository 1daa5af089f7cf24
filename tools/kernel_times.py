"""Per-kernel CUDA-event timings of one PM cycle (development aid; bench.py is the judged number).

    [PM_LIB_PATH=…] python tools/kernel_times.py [--grid 512] [--n 256] [--order 2] [--diff 2] [--dtype f64] [--reps 7] [--tag x]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200 import _lib  # noqa: E402
if os.environ.get('PM_LIB_PATH'):
    _lib.LIB_PATH = os.path.abspath(os.environ['PM_LIB_PATH'])
from concept_b200.pmsolver import PMContext, make_kick_params  # noqa: E402
from concept_b200.synthetic import zeldovich_particles  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=256)
    ap.add_argument('--grid', type=int, default=512)
    ap.add_argument('--order', type=int, default=2)
    ap.add_argument('--diff', type=int, default=2)
    ap.add_argument('--dtype', default='f64')
    ap.add_argument('--sigma', type=float, default=0.3)
    ap.add_argument('--reps', type=int, default=7)
    ap.add_argument('--tag', default='')
    ap.add_argument('--solve-mode', default='auto')
    ap.add_argument('--sort', action='store_true', help='re-order the particles by grid cell first (pm_sort_particles)')
    ap.add_argument('--check', action='store_true', help='compare the potential with the cuFFT path of the same library')
    a = ap.parse_args()
    L = 512.0*a.grid/512
    pos, mom = zeldovich_particles(a.n, L, a.sigma, seed=0, device='cuda')
    N = pos.shape[0]
    ctx = PMContext(a.grid, L, dtype=a.dtype)
    ctx.set_fused_solve(a.solve_mode)
    if a.sort:
        ctx.sort_particles(pos, mom)
    p = make_kick_params(mass=1.0, boxsize=L, gridsize=a.grid, order=a.order, G_Newton=4.4985024439973154e-05,
                         dt_rho_over_dt1=2.0, dt_kick=1e-3, diff_order=a.diff)
    s = torch.zeros(1, dtype=torch.float64, device='cuda')
    staged = ctx.hand_fft_available
    stages = [('grid_zero', lambda: ctx.grid_zero()), ('deposit', lambda: ctx.deposit(pos, p.order, p.contribution))]
    if staged:
        stages += [('fft2d_fwd', lambda: ctx.solve_fused_stage(p.prefactor, p.deconv_order, p.gauss, 1)),
                   ('xsolve', lambda: ctx.solve_fused_stage(p.prefactor, p.deconv_order, p.gauss, 2)),
                   ('fft2d_inv', lambda: ctx.solve_fused_stage(p.prefactor, p.deconv_order, p.gauss, 3))]
    else:
        stages += [('solve', lambda: ctx.solve_fused(p.prefactor, p.deconv_order, p.gauss))]
    stages += [('gather_kick_drift', lambda: ctx.gather_kick_drift(pos, mom, p.order, p.diff_order, p.kick_factor, 1e-4, None, s))]
    times = {k: [] for k, _ in stages}
    total = []
    for rep in range(a.reps + 2):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
        evs[0].record()
        for i, (_, fn) in enumerate(stages):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        if rep >= 2:
            for i, (k, _) in enumerate(stages):
                times[k].append(evs[i].elapsed_time(evs[i + 1]))
            total.append(evs[0].elapsed_time(evs[-1]))
    ctx.check_async_error()
    med = lambda v: sorted(v)[len(v)//2]
    out = {'tag': a.tag, 'lib': os.path.basename(_lib.LIB_PATH), 'grid': a.grid, 'N': N, 'order': a.order, 'diff': a.diff, 'dtype': a.dtype,
           'ms': {k: round(med(v), 4) for k, v in times.items()}, 'cycle_ms': round(med(total), 4)}
    if a.check and staged:
        # potential of the hand-written transforms vs the cuFFT 3-D path of the same library on the same density
        ctx.grid_zero(); ctx.deposit(pos, p.order, p.contribution)
        ctx.solve_fused(p.prefactor, p.deconv_order, p.gauss)
        phi = torch.as_tensor(ctx.get_grid())
        ctx.grid_zero(); ctx.deposit(pos, p.order, p.contribution)
        ctx.fft_forward(); ctx.kspace_potential(p.prefactor, p.deconv_order, p.gauss, 1.0); ctx.fft_backward()
        ref = torch.as_tensor(ctx.get_grid())
        out['phi_relerr_vs_cufft'] = float((phi - ref).abs().max()/ref.abs().max())
    print(json.dumps(out))


if __name__ == '__main__':
    main()
