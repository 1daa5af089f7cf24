"""Runs the fused Poisson solve a few times at the benchmark size (for ncu captures).
env: PM_G, PM_DTYPE, PM_SOLVE_MODE (PMContext.SOLVE_MODES key), PM_REPS"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200.pmsolver import PMContext  # noqa: E402
from concept_b200.synthetic import zeldovich_particles  # noqa: E402

L, G = 512.0, int(os.environ.get('PM_G', 512))
ctx = PMContext(G, L, dtype=os.environ.get('PM_DTYPE', 'f64'))
ctx.set_fused_solve(os.environ.get('PM_SOLVE_MODE', 'auto'))
pos, _ = zeldovich_particles(G//2, L, 0.3, seed=0, device='cuda')
for _ in range(int(os.environ.get('PM_REPS', 2))):
    ctx.grid_zero()
    ctx.deposit(pos, 2, 1.0)
    ctx.solve_fused(-1.0, 4, 0.0)
torch.cuda.synchronize()
ctx.check_async_error()
print('ok')
