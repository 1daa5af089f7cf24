"""Wall time per base step of the host time loop at the benchmark size (development aid): 256^3 particles realised from
a parameter file, PM on a 512^3 grid, `main.run` for a few dozen steps; the step times come from the on_step callback
(device synchronised), so they contain everything the host mirror does between two kicks.

    python tools/timeloop_times.py [--size 256] [--steps 40] [--method pm|p3m]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200 import communication, main, mesh  # noqa: E402


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--method', default='pm')
    a = ap.parse_args()
    communication.init()
    n, G = a.size, 2*a.size
    param = f'''
initial_conditions = {{'species': 'matter', 'N': {n}**3}}
output_times = {{'powerspec': 1.0}}
output_dirs = '/tmp/timeloop_times_output'
boxsize = {2*n}*Mpc/h
potential_options = {{'gridsize': {{'gravity': {{'{a.method}': {G}}}}}}}
select_forces = {{'matter': {{'gravity': '{a.method}'}}}}
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = 0.02
primordial_spectrum = {{'A_s': 2.1e-9, 'n_s': 0.96}}
'''
    stamps = []

    def on_step(time_step, t, scale_factor, Δt):
        torch.cuda.synchronize()
        stamps.append((time.perf_counter(), scale_factor))
    t0 = time.perf_counter()
    main.run(param, on_step=on_step, max_steps=a.steps)
    torch.cuda.synchronize()
    dts = [1e3*(b[0] - a_[0]) for a_, b in zip(stamps, stamps[1:])]
    tail = sorted(dts[len(dts)//4:])
    if communication.master:
        print(json.dumps({'ranks': communication.nprocs, 'particles': n**3, 'grid': G, 'method': a.method, 'steps': len(stamps),
                          'setup_and_first_step_s': round(stamps[0][0] - t0, 3),
                          'ms_per_base_step_median': round(tail[len(tail)//2], 3), 'ms_per_base_step_min': round(tail[0], 3),
                          'ms_per_base_step_max': round(tail[-1], 3), 'a_reached': stamps[-1][1]}))
    mesh.free_contexts()
    communication.finalize()


if __name__ == '__main__':
    run()
