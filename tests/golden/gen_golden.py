"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED
reference (CO*N*CEPT, /root/reference) in its own pure-Python mode under the
sandbox of oracle/ref_sandbox.py.

Run in the build container only:   python tests/golden/gen_golden.py [case ...]
The GPU box never runs this (no /root/reference there); it only reads the .npz.

Each case is one fresh Python process because the reference freezes parameters
into module globals at import (commons.py:2040-2042).  What is recorded per case:
inputs (pos, mom, scalars, ᔑdt), and the reference's outputs — `mom` after
interactions.gravity(...) (interactions.py:2854), plus taps of the deposited
density grid, the real-space potential and the three force grids (without
ghosts), captured by wrapping mesh functions in the interactions namespace.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SANDBOX = '/tmp/concept_ref_sandbox'

# name -> dict(param=..., kind=..., **kw)
_COSMO = '''
H0 = 70*km/s/Mpc
Ωcdm = 0.25
Ωb = 0.05
a_begin = 0.02
enable_class_background = False
'''


def _pm(boxsize, G, method, interp, diff, interlace=False, deconv=True):
    return dict(boxsize=boxsize, G=G, method=method, interp=interp, diff=diff, interlace=interlace, deconv=deconv)


def _pm_param(boxsize, G, method, interp, diff, interlace=False, deconv=True):
    return f'''
boxsize = {boxsize}*Mpc
potential_options = {{
    'gridsize': {{'gravity': {{'{method}': {G}}}}},
    'interpolation': {{'gravity': {{'{method}': '{interp}'}}}},
    'deconvolve': {{'gravity': {{'{method}': ({deconv}, {deconv})}}}},
    'interlace': {{'gravity': {{'{method}': ({interlace}, {interlace})}}}},
    'differentiation': {{'default': {{'gravity': {{'pm': {diff!r}, 'p3m': {diff!r}}}}}}},
}}
select_forces = {{'matter': {{'gravity': '{method}'}}}}
''' + _COSMO


CASES = {
    # PM long-range kicks: interpolation orders, differentiation orders
    'kick_pm_cic_G8_d2':   dict(kind='kick', method='pm', pm=_pm(8, 8, 'pm', 'CIC', 2), N=512, seed=1, boxsize=8.0),
    'kick_pm_ngp_G8_d2':   dict(kind='kick', method='pm', pm=_pm(8, 8, 'pm', 'NGP', 2), N=300, seed=2, boxsize=8.0),
    'kick_pm_tsc_G12_d4':  dict(kind='kick', method='pm', pm=_pm(30, 12, 'pm', 'TSC', 4), N=512, seed=3, boxsize=30.0),
    'kick_pm_pcs_G10_d6':  dict(kind='kick', method='pm', pm=_pm(20, 10, 'pm', 'PCS', 6), N=400, seed=4, boxsize=20.0),
    'kick_pm_cic_G16_d8':  dict(kind='kick', method='pm', pm=_pm(64, 16, 'pm', 'CIC', 8), N=700, seed=5, boxsize=64.0),
    'kick_pm_cic_G8_d1':   dict(kind='kick', method='pm', pm=_pm(8, 8, 'pm', 'CIC', 1), N=256, seed=6, boxsize=8.0),
    'kick_pm_cic_G12_fourier': dict(kind='kick', method='pm', pm=_pm(24, 12, 'pm', 'CIC', 'fourier'), N=512, seed=7, boxsize=24.0),
    'kick_pm_tsc_G8_interlace': dict(kind='kick', method='pm', pm=_pm(8, 8, 'pm', 'TSC', 2, interlace=True), N=256, seed=8, boxsize=8.0),
    'kick_pm_cic_G8_nodeconv': dict(kind='kick', method='pm', pm=_pm(8, 8, 'pm', 'CIC', 2, deconv=False), N=256, seed=9, boxsize=8.0),
    # Edge cases: particles exactly on cell edges / box boundary, clustered in one cell
    'kick_pm_cic_G8_edges': dict(kind='kick', method='pm', pm=_pm(8, 8, 'pm', 'CIC', 2), N=343, seed=10, boxsize=8.0, edges=True),
    'kick_pm_tsc_G8_cluster': dict(kind='kick', method='pm', pm=_pm(8, 8, 'pm', 'TSC', 2), N=300, seed=11, boxsize=8.0, cluster=True),
    # P3M long-range part (Gaussian-split Green's function), default diff order 4
    'kick_p3m_long_cic_G24': dict(kind='kick', method='p3m', pm=_pm(48, 24, 'p3m', 'CIC', 4), N=512, seed=12, boxsize=48.0),
    'kick_p3m_long_tsc_G24': dict(kind='kick', method='p3m', pm=_pm(48, 24, 'p3m', 'TSC', 4), N=512, seed=13, boxsize=48.0),
    # Drift
    'drift_G8': dict(kind='drift', method='pm', pm=_pm(8, 8, 'pm', 'CIC', 2), N=512, seed=14, boxsize=8.0),
}


def worker(name):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    case = CASES[name]
    ref_sandbox.enter_reference(SANDBOX, _pm_param(**case['pm']), jobid=abs(hash(name)) % 100000 + 1)
    import commons
    from commons import universals, boxsize, G_Newton, nghosts
    import species, mesh, interactions
    rng = np.random.Generator(np.random.PCG64DXSM(case['seed']))
    N = case['N']
    L = float(boxsize)
    assert abs(L - case['boxsize']) < 1e-12
    pos = rng.random((N, 3))*L
    if case.get('edges'):
        # lattice exactly on cell edges (pos = i*cellsize) incl. 0; some at L-tiny
        G = 8
        n = round(N**(1/3))
        ijk = np.stack(np.meshgrid(*(np.arange(n),)*3, indexing='ij'), -1).reshape(-1, 3)
        pos = (ijk % G)*(L/G)*1.0
        pos[::7] += 0.5*(L/G)          # some exactly on cell centres
        pos[3::11] = np.nextafter(L, 0)  # largest representable value below L
        pos = np.mod(pos, L)
        pos[pos == L] = 0
    if case.get('cluster'):
        pos = (0.37 + 0.05*rng.random((N, 3)))*L  # all within ~half a cell
        pos[:20] = rng.random((20, 3))*L
    mass = 3.7 + 0.1*case['seed']
    mom = rng.standard_normal((N, 3))*mass*0.5
    a = 0.5
    universals.a = a
    universals.t = 1.0
    comp = species.Component('matter', 'matter', N=N, mass=mass)
    for d, s in enumerate('xyz'):
        comp.populate(np.ascontiguousarray(pos[:, d]), 'pos' + s)
        comp.populate(np.ascontiguousarray(mom[:, d]), 'mom' + s)
    pm = case['pm']
    out = dict(pos=pos.copy(), mom=mom.copy(), mass=mass, a=a, boxsize=L, G_Newton=float(G_Newton),
               nghosts=int(nghosts), gridsize=pm['G'], method=pm['method'],
               order={'NGP': 1, 'CIC': 2, 'TSC': 3, 'PCS': 4}[pm['interp']],
               diff_order=0 if pm['diff'] == 'fourier' else pm['diff'],
               interlace=bool(pm['interlace']), deconvolve=bool(pm['deconv']))
    Δt = 0.0123
    if case['kind'] == 'drift':
        mom_big = rng.standard_normal((N, 3))*mass*L*40  # several box crossings
        for d, s in enumerate('xyz'):
            comp.populate(np.ascontiguousarray(mom_big[:, d]), 'mom' + s)
        ᔑdt = {'a**(-2)': Δt/a**2, '1': Δt}
        comp.drift(ᔑdt)
        out.update(mom=mom_big, dt_am2=ᔑdt['a**(-2)'], pos_out=np.array(comp.pos_mv3[:N]).copy())
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        return
    ᔑdt = {
        '1': Δt,
        ('a**(-3*w_eff-1)', 'matter'): Δt/a*1.01,
        ('a**(-3*w_eff)', 'matter'): Δt*0.99,
    }
    taps = {}
    ng = int(nghosts)

    def noghosts(g):
        g = np.asarray(g)
        return g[ng:-ng, ng:-ng, ng:-ng].copy()
    orig_diff = interactions.diff_domaingrid
    def diff_tap(grid, dim, order, *args, **kw):
        if 'phi' not in taps:
            taps['phi'] = noghosts(grid)
        res = orig_diff(grid, dim, order, *args, **kw)
        taps[f'force{dim}'] = noghosts(res)
        return res
    interactions.diff_domaingrid = diff_tap
    orig_add = mesh.add_upstream_to_global_slabs
    def add_tap(grid_upstream, *args, **kw):
        # grid_upstream has been through communicate_ghosts('+=') (mesh.py:609)
        key = 'rho' if 'rho' not in taps else 'rho_shifted'
        taps[key] = noghosts(grid_upstream)
        return orig_add(grid_upstream, *args, **kw)
    mesh.add_upstream_to_global_slabs = add_tap
    orig_apply = interactions.apply_particle_mesh_force
    def apply_tap(grid, dim, *args, **kw):
        key = f'forcegrid{dim}' if f'forcegrid{dim}' not in taps else f'forcegrid{dim}_shifted'
        taps[key] = noghosts(grid)
        return orig_apply(grid, dim, *args, **kw)
    interactions.apply_particle_mesh_force = apply_tap
    interactions.gravity(case['method'], [comp], [comp], ᔑdt, 'long-range', True)
    out.update(
        dt_1=ᔑdt['1'], dt_rho=ᔑdt['a**(-3*w_eff-1)', 'matter'], dt_kick=ᔑdt['a**(-3*w_eff)', 'matter'],
        mom_out=np.array(comp.mom_mv3[:N]).copy(),
        pos_after=np.array(comp.pos_mv3[:N]).copy(),
    )
    if case['method'] == 'p3m':
        out['r_scale'] = float(commons.shortrange_params['gravity']['scale'])
    out.update({'tap_' + k: v for k, v in taps.items()})
    assert np.array_equal(out['pos_after'], pos)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'ok; taps:', sorted(taps))


def main():
    names = sys.argv[1:] or list(CASES)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    if not os.path.isdir(SANDBOX + '/src'):
        ref_sandbox.build_sandbox(SANDBOX)
    procs = []
    for name in names:
        procs.append((name, subprocess.Popen([sys.executable, __file__, '--worker', name],
                                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        if len(procs) >= 8:
            n, p = procs.pop(0)
            o, _ = p.communicate()
            print(f'[{n}] rc={p.returncode}\n' + '\n'.join(o.strip().split('\n')[-4:]))
    for n, p in procs:
        o, _ = p.communicate()
        print(f'[{n}] rc={p.returncode}\n' + '\n'.join(o.strip().split('\n')[-4:]))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--worker':
        worker(sys.argv[2])
    else:
        main()
