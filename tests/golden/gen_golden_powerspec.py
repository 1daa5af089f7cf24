"""Golden vectors for the power-spectrum estimator: runs the UNMODIFIED reference
(analysis.get_powerspec_declarations + analysis.compute_powerspec, analysis.py:235-579) in its
pure-Python mode under oracle/ref_sandbox.py and records bins and P(k).

Run in the build container only:   python tests/golden/gen_golden_powerspec.py [case ...]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SANDBOX = '/tmp/concept_ref_sandbox'

CASES = {
    # defaults of powerspec_options (commons.py:3354-3385): PCS, deconvolution, bcc interlacing, k_max = Nyquist
    'powerspec_pcs_G16': dict(G=16, boxsize=64.0, N=1500, seed=31, interp='PCS', interlace=True, deconv=True),
    'powerspec_pcs_G32': dict(G=32, boxsize=200.0, N=4000, seed=34, interp='PCS', interlace=True, deconv=True),
    'powerspec_cic_G12_plain': dict(G=12, boxsize=30.0, N=900, seed=32, interp='CIC', interlace=False, deconv=True),
    'powerspec_tsc_G20_nodeconv': dict(G=20, boxsize=100.0, N=1200, seed=33, interp='TSC', interlace=True, deconv=False),
}


def param_text(c):
    return f'''
boxsize = {c['boxsize']}*Mpc
H0 = 70*km/s/Mpc
Ωcdm = 0.25
Ωb = 0.05
a_begin = 0.02
enable_class_background = False
powerspec_options = {{
    'gridsize': {c['G']},
    'interpolation': '{c['interp']}',
    'interlace': {c['interlace']},
    'deconvolve': {c['deconv']},
}}
powerspec_select = {{'all': {{'data': True, 'linear': False, 'corrected': False, 'plot': False}}}}
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
    N, L = c['N'], float(boxsize)
    # clustered particle load: uniform + a few Gaussian clumps, so that P(k) is not pure shot noise
    pos = rng.random((N, 3))*L
    centres = rng.random((5, 3))*L
    k = N//2
    pos[:k] = (centres[rng.integers(0, 5, k)] + rng.standard_normal((k, 3))*0.04*L) % L
    mass = 2.5
    a = 0.7
    universals.a = a
    universals.t = 1.0
    comp = species.Component('matter', 'matter', N=N, mass=mass)
    for d, s in enumerate('xyz'):
        comp.populate(np.ascontiguousarray(pos[:, d]), 'pos' + s)
        comp.populate(np.zeros(N), 'mom' + s)
    decl = analysis.get_powerspec_declarations([comp])[0]
    analysis.compute_powerspec(decl)
    out = dict(pos=pos, mass=mass, a=a, boxsize=L, gridsize=int(decl.gridsize), order=int(decl.interpolation),
               deconvolve=bool(decl.deconvolve), interlace=str(decl.interlace), varrho_bar=float(comp.ϱ_bar),
               w_eff=float(comp.w_eff(a=a)), k2_max=int(decl.k2_max), k_bin_indices=np.asarray(decl.k_bin_indices).copy(),
               k_bin_centers=np.asarray(decl.k_bin_centers).copy(), n_modes=np.asarray(decl.n_modes).copy(),
               power=np.asarray(decl.power).copy(),
               bins_per_decade_keys=np.array([str(k) for k in decl.bins_per_decade.keys()]),
               bins_per_decade_vals=np.array([float(v) for v in decl.bins_per_decade.values()]),
               k_max=str(decl.k_max))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'ok: bins', len(out['power']), 'k2_max', out['k2_max'], 'P[0..3]', out['power'][:3])


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
