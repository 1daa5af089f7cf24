"""Short-range pair kick of 256^3 particles on a 512^3-grid scale (development aid): run under
ncu --metrics gpu__time_duration.sum to list the kernels of one kick."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200.pmsolver import PMContext  # noqa: E402
from concept_b200.shortrange import PairKick  # noqa: E402
from concept_b200.synthetic import zeldovich_particles  # noqa: E402

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
G, L = 2*n_side, 2.0*n_side
pos, mom = zeldovich_particles(n_side, L, 0.3, seed=0, device='cuda')
ctx = PMContext(G, L)
job = PairKick(ctx, L, G, pos.shape[0], pos.shape[0], 4.4985024439973154e-05)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    job.kick(pos, mom, pos.shape[0])
    torch.cuda.synchronize(); print(f'kick {1e3*(time.perf_counter() - t0):.3f} ms')
print(job.pair_stats(pos, pos.shape[0]))
