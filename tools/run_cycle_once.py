"""Runs a few PM cycles at the benchmark size (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200.pmsolver import PMContext, make_kick_params  # noqa: E402
from concept_b200.synthetic import zeldovich_particles  # noqa: E402

L, G, n = 512.0, int(os.environ.get('PM_G', 512)), int(os.environ.get('PM_N', 256))
pos, mom = zeldovich_particles(n, L, 0.3, seed=0, device='cuda')
ctx = PMContext(G, L, dtype=os.environ.get('PM_DTYPE', 'f64'))
p = make_kick_params(mass=1.0, boxsize=L, gridsize=G, order=int(os.environ.get('PM_ORDER', 2)), G_Newton=4.4985e-5,
                     dt_rho_over_dt1=2.0, dt_kick=1e-3)
s = torch.zeros(1, dtype=torch.float64, device='cuda')
for _ in range(int(os.environ.get('PM_CYCLES', 3))):
    ctx.kick_long(pos, mom, p, sum_mom2=s)
    ctx.drift(pos, mom, 1e-4)
torch.cuda.synchronize()
print('ok', s.item())
