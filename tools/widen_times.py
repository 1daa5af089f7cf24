"""Wall times of the steps either side of the stepping path at the benchmark size (development aid): realisation of
256^3 particles from a parameter file (primordial noise on the host, 1LPT or 2LPT on the GPU), the P(k) estimator on a
512^3 grid (PCS, two interlaced lattices), a GADGET-2 snapshot written and read back, and pm_sort_particles.

    python tools/widen_times.py [--size 256] [--lpt 1|2]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200 import analysis, commons, main, mesh, snapshot  # noqa: E402


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--lpt', type=int, default=1)
    a = ap.parse_args()
    n, G = a.size, 2*a.size
    commons.load_params(f'''
initial_conditions = {{'species': 'matter', 'N': {n}**3}}
boxsize = {2*n}*Mpc/h
potential_options = {G}
realization_options = {{'LPT': {a.lpt}}}
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = 0.02
primordial_spectrum = {{'A_s': 2.1e-9, 'n_s': 0.96}}
''')
    main.init_time()
    out = {'particles': n**3, 'grid': G, 'lpt': a.lpt}
    out['realize_s'], comps = wall(main.get_initial_conditions)
    c = comps[0]
    out['powerspec_first_s'], _ = wall(lambda: analysis.powerspec([c], G))
    out['powerspec_s'], pk = wall(lambda: analysis.powerspec([c], G))
    d = tempfile.mkdtemp(prefix='snap_')
    path = os.path.join(d, 'snap')
    out['snapshot_save_s'], _ = wall(lambda: snapshot.save(c, path))
    files = [os.path.join(d, f) for f in os.listdir(d)]
    out['snapshot_bytes'] = sum(os.path.getsize(f) for f in files)
    a_now = commons.universals.a
    out['snapshot_load_s'], _ = wall(lambda: snapshot.load(files[0] if len(files) == 1 else path))
    commons.universals.a = a_now
    ctx = mesh.get_context(G)
    out['cell_sort_s'], _ = wall(lambda: ctx.sort_particles(c.pos_local, c.mom_local, c.ids[:c.N_local]))
    print(json.dumps(out))


if __name__ == '__main__':
    run()
