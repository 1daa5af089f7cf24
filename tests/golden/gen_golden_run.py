"""Golden FULL RUN: the unmodified reference's main.timeloop (main.py:102) on the
test/pure_python_pm configuration (8³ particles, 8³ PM grid, boxsize 8 Mpc, a = 0.02 → 1,
snapshots at a = 0.1, 0.5, 1), in pure-Python mode under oracle/ref_sandbox.py.

Recorded in tests/golden/run_pm_8.npz:
  ICs (pos0, mom0, mass), every long-range kick (t_start, t_end implied by its ᔑdt scalars) and every
  drift (ᔑdt['a**(-2)']) in call order, the Δt / a / t history, and pos/mom at each snapshot time.
ICs are injected (the reference's IC generator needs CLASS, absent here): a cell-centred lattice
with seeded random displacements and small random momenta.

    python tests/golden/gen_golden_run.py            (≈10-15 min of pure-Python reference time)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SANDBOX = '/tmp/concept_ref_sandbox_run'

PARAM = '''
initial_conditions = 'injected'
output_dirs        = {'snapshot': param.dir + '/output'}
output_bases       = {'snapshot': 'snapshot'}
output_times       = {'snapshot': (0.1, 0.5, 1)}
snapshot_type      = 'concept'
boxsize = 8*Mpc
potential_options = {'gridsize': {'gravity': {'pm': 8}}}
H0      = 70*km/s/Mpc
Ωcdm    = 0.25
Ωb      = 0.05
a_begin = 0.02
select_forces = {'matter': {'gravity': 'pm'}}
enable_class_background = False
'''


def main():
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    ref_sandbox.build_sandbox(SANDBOX)
    ref_sandbox.enter_reference(SANDBOX, PARAM, jobid=7)
    import commons
    from commons import universals, boxsize, G_Newton, H0, Ωb, Ωcdm, ρ_crit
    import species, snapshot, interactions
    N = 8**3
    L = float(boxsize)
    rng = np.random.Generator(np.random.PCG64DXSM(2024))
    idx = (np.arange(8) + 0.5)*(L/8)
    q = np.stack(np.meshgrid(idx, idx, idx, indexing='ij'), -1).reshape(-1, 3)
    pos0 = np.mod(q + 0.05*(L/8)*rng.standard_normal((N, 3)), L)
    mass = float((Ωb + Ωcdm)*ρ_crit*L**3/N)
    mom0 = 1e-3*mass*rng.standard_normal((N, 3))*(L/8)   # small peculiar momenta
    log = dict(kick_dt1=[], kick_dtrho=[], kick_dtkick=[], kick_t=[], kick_a=[], drift_dt=[], drift_t=[],
               order=[])
    snaps = {}

    def get_ic(*args, **kw):
        comp = species.Component('matter', 'matter', N=N, mass=mass)
        for d, s in enumerate('xyz'):
            comp.populate(np.ascontiguousarray(pos0[:, d]), 'pos' + s)
            comp.populate(np.ascontiguousarray(mom0[:, d]), 'mom' + s)
        return [comp]

    def save(components, filename, *args, **kw):
        c = components[0]
        key = f'{universals.a:.6f}'
        snaps[key] = (np.array(c.pos_mv3[:N]).copy(), np.array(c.mom_mv3[:N]).copy(), float(universals.t))
        print('SNAPSHOT at a =', universals.a, flush=True)
        return filename
    snapshot.get_initial_conditions = get_ic
    snapshot.save = save
    orig_gravity = interactions.gravity

    def gravity_tap(method, receivers, suppliers, ᔑdt, interaction_type, printout):
        log['kick_dt1'].append(ᔑdt['1'])
        log['kick_dtrho'].append(ᔑdt['a**(-3*w_eff-1)', 'matter'])
        log['kick_dtkick'].append(ᔑdt['a**(-3*w_eff)', 'matter'])
        log['kick_t'].append(universals.t)
        log['kick_a'].append(universals.a)
        log['order'].append(0)
        return orig_gravity(method, receivers, suppliers, ᔑdt, interaction_type, printout)
    interactions.gravity = gravity_tap
    orig_drift = species.Component.drift

    def drift_tap(self, ᔑdt, a_next=-1):
        log['drift_dt'].append(ᔑdt['a**(-2)'])
        log['drift_t'].append(universals.t)
        log['order'].append(1)
        return orig_drift(self, ᔑdt, a_next)
    species.Component.drift = drift_tap
    try:
        import main  # noqa: F401  — importing main with jobid != -1 runs timeloop() (main.py:2437-2473)
    except SystemExit as e:
        print('reference exited with', e.code)
    out = dict(pos0=pos0, mom0=mom0, mass=mass, boxsize=L, G_Newton=float(G_Newton), H0=float(H0), Omega_m=float(Ωb + Ωcdm),
               gridsize=8, a_begin=0.02, rho_crit=float(ρ_crit))
    for k, v in log.items():
        out[k] = np.asarray(v)
    for key, (p, m, t) in snaps.items():
        out[f'snap_pos_{key}'] = p
        out[f'snap_mom_{key}'] = m
        out[f'snap_t_{key}'] = t
    np.savez_compressed(os.path.join(HERE, 'run_pm_8.npz'), **out)
    print('kicks', len(log['kick_dt1']), 'drifts', len(log['drift_dt']), 'snapshots', sorted(snaps))


if __name__ == '__main__':
    main()
