"""Golden vectors for PM kicks with component-specific upstream/downstream grid sizes:
runs the UNMODIFIED reference (interactions.gravity → particle_mesh :1985-2335 → mesh.interpolate_upstream
:492-616, add_upstream_to_global_slabs :618-710, copy_modes :980-1322) in its pure-Python mode under
oracle/ref_sandbox.py with two particle components whose upstream/downstream grids differ from the global one.

Run in the build container only:   python tests/golden/gen_golden_multigrid.py [case ...]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SANDBOX = '/tmp/concept_ref_sandbox'

CASES = {
    # (upstream, downstream) per component and the global grid size
    'multigrid_pm_cic_8_12_g12': dict(boxsize=24.0, gglobal=12, grids={'cdm': (8, 8), 'baryons': (12, 12)}, interp='CIC',
                                      diff=2, interlace=False, deconv=True, N={'cdm': 300, 'baryons': 200}, seed=41),
    'multigrid_pm_tsc_16_8_g12_up_down': dict(boxsize=30.0, gglobal=12, grids={'cdm': (16, 10), 'baryons': (8, 16)}, interp='TSC',
                                              diff=4, interlace=False, deconv=True, N={'cdm': 250, 'baryons': 150}, seed=42),
    'multigrid_pm_cic_interlace_fourier': dict(boxsize=20.0, gglobal=10, grids={'cdm': (8, 12), 'baryons': (10, 10)}, interp='CIC',
                                               diff='fourier', interlace=True, deconv=True, N={'cdm': 200, 'baryons': 120}, seed=43),
    'multigrid_p3m_long_pcs_nodeconv': dict(boxsize=48.0, gglobal=24, grids={'cdm': (24, 24), 'baryons': (28, 20)}, interp='PCS',
                                            diff=4, interlace=False, deconv=False, N={'cdm': 300, 'baryons': 100}, seed=44,
                                            method='p3m'),
}


def param_text(c):
    m = c.get('method', 'pm')
    grids = ''.join(f"        '{name}': {{'gravity': {{'{m}': {tuple(g)}}}}},\n" for name, g in c['grids'].items())
    return f'''
boxsize = {c['boxsize']}*Mpc
potential_options = {{
    'gridsize': {{
        'global': {{'gravity': {{'{m}': {c['gglobal']}}}}},
{grids}    }},
    'interpolation': {{'gravity': {{'{m}': '{c['interp']}'}}}},
    'deconvolve': {{'gravity': {{'{m}': ({c['deconv']}, {c['deconv']})}}}},
    'interlace': {{'gravity': {{'{m}': ({c['interlace']}, {c['interlace']})}}}},
    'differentiation': {{'default': {{'gravity': {{'pm': {c['diff']!r}, 'p3m': {c['diff']!r}}}}}}},
}}
select_forces = {{'all': {{'gravity': '{m}'}}}}
H0 = 70*km/s/Mpc
Ωcdm = 0.25
Ωb = 0.05
a_begin = 0.02
enable_class_background = False
'''


def worker(name):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    c = CASES[name]
    ref_sandbox.enter_reference(SANDBOX, param_text(c), jobid=abs(hash(name)) % 100000 + 1)
    import commons
    from commons import universals, boxsize, G_Newton, shortrange_params
    import species, interactions
    rng = np.random.Generator(np.random.PCG64DXSM(c['seed']))
    L = float(boxsize)
    a = 0.5
    universals.a, universals.t = a, 1.0
    Δt = 0.0123
    method = c.get('method', 'pm')
    comps, out = [], dict(boxsize=L, a=a, G_Newton=float(G_Newton), gridsize_global=c['gglobal'], method=method,
                          order={'NGP': 1, 'CIC': 2, 'TSC': 3, 'PCS': 4}[c['interp']],
                          diff_order=0 if c['diff'] == 'fourier' else c['diff'], interlace=bool(c['interlace']),
                          deconvolve=bool(c['deconv']), names=np.array(list(c['grids'])),
                          r_scale=float(shortrange_params['gravity']['scale']) if method == 'p3m' else 0.0)
    ᔑdt = {'1': Δt}
    for q, (cname, N) in enumerate(c['N'].items()):
        pos = rng.random((N, 3))*L
        pos[:N//3] = (0.3 + 0.1*q + 0.08*rng.standard_normal((N//3, 3)))*L % L
        mass = 2.9 + 1.3*q
        mom = rng.standard_normal((N, 3))*mass*0.5
        comp = species.Component(cname, {'cdm': 'cold dark matter', 'baryons': 'baryons'}[cname], N=N, mass=mass)
        for d, s in enumerate('xyz'):
            comp.populate(np.ascontiguousarray(pos[:, d]), 'pos' + s)
            comp.populate(np.ascontiguousarray(mom[:, d]), 'mom' + s)
        g = comp.potential_gridsizes['gravity'][method]
        assert (int(g.upstream), int(g.downstream)) == tuple(c['grids'][cname]), (g, c['grids'][cname])
        comps.append(comp)
        ᔑdt['a**(-3*w_eff-1)', cname] = Δt/a*(1.01 + 0.02*q)
        ᔑdt['a**(-3*w_eff)', cname] = Δt*(0.99 - 0.03*q)
        out.update({f'pos_{cname}': pos.copy(), f'mom_{cname}': mom.copy(), f'mass_{cname}': mass,
                    f'grids_{cname}': np.array(c['grids'][cname]), f'dt_rho_{cname}': ᔑdt['a**(-3*w_eff-1)', cname],
                    f'dt_kick_{cname}': ᔑdt['a**(-3*w_eff)', cname]})
    out['dt1'] = Δt
    interactions.gravity(method, comps, comps, ᔑdt, 'long-range', True)
    for comp in comps:
        N = comp.N
        out[f'mom_out_{comp.name}'] = np.array(comp.mom_mv3[:N]).copy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'ok', {comp.name: float(np.abs(out[f'mom_out_{comp.name}'] - out[f'mom_{comp.name}']).max()) for comp in comps})


def main():
    names = sys.argv[1:] or list(CASES)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    if not os.path.isdir(SANDBOX + '/src'):
        ref_sandbox.build_sandbox(SANDBOX)
    procs = [(n, subprocess.Popen([sys.executable, __file__, '--worker', n], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True)) for n in names]
    for n, p in procs:
        o, _ = p.communicate()
        print(f'[{n}] rc={p.returncode}\n' + '\n'.join(o.strip().split('\n')[-3:]))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--worker':
        worker(sys.argv[2])
    else:
        main()
