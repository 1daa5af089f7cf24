"""Golden vectors for the power spectrum of a GROUP of components with component-specific upstream grid sizes
(analysis.compute_powerspec, analysis.py:500-579 → mesh.interpolate_upstream :492-616 → add_upstream_to_global_slabs
:618-710 → copy_modes :980-1322): the UNMODIFIED reference in its pure-Python mode under oracle/ref_sandbox.py.

Run in the build container only:   python tests/golden/gen_golden_powerspec_group.py [case ...]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SANDBOX = '/tmp/concept_ref_sandbox'

CASES = {
    'pkgroup_pcs_12_16_g16': dict(boxsize=64.0, up={'cdm': 12, 'baryons': 16}, gglobal=16, interp='PCS', interlace=True,
                                          deconv=True, N={'cdm': 900, 'baryons': 700}, seed=51),
    'pkgroup_cic_20_12_g16': dict(boxsize=100.0, up={'cdm': 20, 'baryons': 12}, gglobal=16, interp='CIC', interlace=False,
                                          deconv=True, N={'cdm': 800, 'baryons': 500}, seed=52),
}


def param_text(c):
    up = ', '.join(f"'{k}': {v}" for k, v in c['up'].items())
    return f'''
boxsize = {c['boxsize']}*Mpc
H0 = 70*km/s/Mpc
Ωcdm = 0.25
Ωb = 0.05
a_begin = 0.02
enable_class_background = False
powerspec_options = {{
    'upstream gridsize': {{{up}}},
    'global gridsize': {{('cdm', 'baryons'): {c['gglobal']}}},
    'interpolation': '{c['interp']}',
    'interlace': {c['interlace']},
    'deconvolve': {c['deconv']},
}}
powerspec_select = {{('cdm', 'baryons'): {{'data': True, 'linear': False, 'corrected': False, 'plot': False}}}}
'''


def worker(name):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    c = CASES[name]
    ref_sandbox.enter_reference(SANDBOX, param_text(c), jobid=abs(hash(name)) % 100000 + 1)
    import commons
    from commons import universals, boxsize
    import species, analysis
    rng = np.random.Generator(np.random.PCG64DXSM(c['seed']))
    L = float(boxsize)
    a = 0.7
    universals.a, universals.t = a, 1.0
    comps, out = [], dict(boxsize=L, a=a, names=np.array(list(c['N'])))
    centres = rng.random((5, 3))*L
    for q, (cname, N) in enumerate(c['N'].items()):
        pos = rng.random((N, 3))*L
        k = N//2
        pos[:k] = (centres[rng.integers(0, 5, k)] + rng.standard_normal((k, 3))*0.04*L) % L
        mass = 2.5 + 0.8*q
        comp = species.Component(cname, {'cdm': 'cold dark matter', 'baryons': 'baryons'}[cname], N=N, mass=mass)
        for d, s in enumerate('xyz'):
            comp.populate(np.ascontiguousarray(pos[:, d]), 'pos' + s)
            comp.populate(np.zeros(N), 'mom' + s)
        comps.append(comp)
        out.update({f'pos_{cname}': pos, f'mass_{cname}': mass, f'varrho_bar_{cname}': float(comp.ϱ_bar),
                    f'w_eff_{cname}': float(comp.w_eff(a=a)), f'upstream_{cname}': c['up'][cname]})
    decls = analysis.get_powerspec_declarations(comps)
    decl = [d for d in decls if len(d.components) == 2][0]
    analysis.compute_powerspec(decl)
    assert [int(comp.powerspec_upstream_gridsize) for comp in comps] == [c['up'][n] for n in c['N']]
    out.update(gridsize=int(decl.gridsize), order=int(decl.interpolation), deconvolve=bool(decl.deconvolve),
               interlace=str(decl.interlace), k2_max=int(decl.k2_max), k_bin_centers=np.asarray(decl.k_bin_centers).copy(),
               n_modes=np.asarray(decl.n_modes).copy(), power=np.asarray(decl.power).copy(),
               bins_per_decade_keys=np.array([str(k) for k in decl.bins_per_decade.keys()]),
               bins_per_decade_vals=np.array([float(v) for v in decl.bins_per_decade.values()]), k_max=str(decl.k_max))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'ok: bins', len(out['power']), 'gridsize', out['gridsize'], 'P[0..3]', out['power'][:3])


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
        print(f'[{n}] rc={p.returncode}\n' + '\n'.join(o.strip().split('\n')[-6:]))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--worker':
        worker(sys.argv[2])
    else:
        main()
