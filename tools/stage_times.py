"""Per-stage CUDA-event timings of one PM cycle (development aid; bench.py is the judged number).

    python tools/stage_times.py [--n 256] [--grid 512] [--order 2] [--dtype f64] [--sigma 0.3] [--reps 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200 import _lib  # noqa: E402
if os.environ.get('PM_LIB_PATH'):      # experiment builds of the library (development only)
    _lib.LIB_PATH = os.environ['PM_LIB_PATH']
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
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--nofuse', action='store_true', help='three-call solve (cuFFT 3-D + k-space kernel) instead of the fused x-solve')
    ap.add_argument('--solve-mode', default='auto', help='auto | fft2_l2 | fft2_split | cufft2d (see PMContext.SOLVE_MODES)')
    ap.add_argument('--nodriftfuse', action='store_true', help='separate gather_kick and drift kernels')
    ap.add_argument('--sort', action='store_true', help='pm_sort_particles before timing (after --shuffle: restores locality); its time is reported')
    ap.add_argument('--shuffle', action='store_true', help='random particle order (worst-case locality)')
    a = ap.parse_args()
    L = 512.0
    pos, mom = zeldovich_particles(a.n, L, a.sigma, seed=0, device='cuda')
    if a.shuffle:
        perm = torch.randperm(pos.shape[0], device='cuda')
        pos, mom = pos[perm].contiguous(), mom[perm].contiguous()
    N = pos.shape[0]
    ctx = PMContext(a.grid, L, dtype=a.dtype)
    ctx.set_fused_solve(a.solve_mode)
    sort_ms = None
    if a.sort:
        ctx.sort_particles(pos, mom)     # warm-up: scratch allocation
        if a.shuffle:
            perm = torch.randperm(pos.shape[0], device='cuda')
            pos, mom = pos[perm].contiguous(), mom[perm].contiguous()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.sort_particles(pos, mom); e1.record(); torch.cuda.synchronize()
        sort_ms = e0.elapsed_time(e1)
    p = make_kick_params(mass=1.0, boxsize=L, gridsize=a.grid, order=a.order, G_Newton=4.4985024439973154e-05,
                         dt_rho_over_dt1=2.0, dt_kick=1e-3, diff_order=a.diff)
    s = torch.zeros(1, dtype=torch.float64, device='cuda')
    stages = [
        ('grid_zero', lambda: ctx.grid_zero()),
        ('deposit', lambda: ctx.deposit(pos, p.order, p.contribution)),
        *([('solve_fused', lambda: ctx.solve_fused(p.prefactor, p.deconv_order, p.gauss))]
          if (ctx.fused_solve_available and not a.nofuse) else [
            ('fft_forward', lambda: ctx.fft_forward()),
            ('kspace', lambda: ctx.kspace_potential(p.prefactor, p.deconv_order, p.gauss, 1.0)),
            ('fft_backward', lambda: ctx.fft_backward())]),
        *([('gather_kick_drift', lambda: ctx.gather_kick_drift(pos, mom, p.order, p.diff_order, p.kick_factor, 1e-4, None, s))]
          if not a.nodriftfuse else [
            ('gather_kick', lambda: ctx.gather_kick(pos, mom, p.order, p.diff_order, p.kick_factor, None, s)),
            ('drift', lambda: ctx.drift(pos, mom, 1e-4))]),
    ]
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
    G3 = a.grid**3
    es = 8 if a.dtype == 'f64' else 4
    alg = {'grid_zero': es*G3, 'deposit': 24*N + es*G3, 'fft_forward': 2*es*G3, 'kspace': 2*es*G3, 'fft_backward': 2*es*G3, 'solve_fused': 4*es*G3,
           'gather_kick': 72*N + es*G3, 'drift': 72*N, 'gather_kick_drift': 96*N + es*G3}
    ctx.check_async_error()
    out = {'N': N, 'grid': a.grid, 'sort_ms': sort_ms, 'solve_mode': a.solve_mode, 'order': a.order, 'dtype': a.dtype, 'sigma': a.sigma, 'shuffle': a.shuffle,
           'device_bytes': ctx.device_bytes, 'stages_ms': {}, 'stages_GBps': {}}
    for k in times:
        t = sorted(times[k])[len(times[k])//2]
        out['stages_ms'][k] = round(t, 4)
        out['stages_GBps'][k] = round(alg[k]/t/1e6, 1)
    t = sorted(total)[len(total)//2]
    out['cycle_ms'] = round(t, 4)
    out['particle_updates_per_s'] = N/(t*1e-3)
    out['B_alg_GB'] = (120*N + 6*es*G3)/1e9
    out['roofline_frac_of_6553.6'] = (120*N + 6*es*G3)/(t*1e-3)/6553.6e9
    print(json.dumps(out))


if __name__ == '__main__':
    main()
