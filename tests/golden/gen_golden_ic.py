"""Golden vectors for the initial-condition generator: runs the UNMODIFIED reference
(ic.realize_particles, ic.py:1199-1399 → generate_primordial_noise :928-1163, realize_grid :670-782,
carryout_1lpt :1447-1509, carryout_2lpt :1539-1589, displace_particles :2249-2283) in its pure-Python
mode under oracle/ref_sandbox.py and records the primordial noise slab, the amplitude tables and the
realised particles.

CLASS is not available, so `ic.compute_transfer` / `ic.compute_cosmo` (the two CLASS-backed look-ups
the realisation makes) are replaced by analytic stand-ins whose parameters are stored in the golden
file; everything downstream of them is the reference's own code.

Run in the build container only:   python tests/golden/gen_golden_ic.py [case ...]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SANDBOX = '/tmp/concept_ref_sandbox'

CASES = {
    'ic_1lpt_sc_G8': dict(n=8, lattices=1, boxsize=64.0),
    'ic_1lpt_sc_G12_backscale': dict(n=12, lattices=1, boxsize=100.0, backscale=True),
    'ic_1lpt_bcc_G6': dict(n=6, lattices=2, boxsize=48.0),
    'ic_1lpt_fcc_G4_backscale': dict(n=4, lattices=4, boxsize=40.0, backscale=True),
    'ic_1lpt_sc_G10_fixed': dict(n=10, lattices=1, boxsize=80.0, fixed=True, phase_shift='π'),
    'ic_1lpt_sc_G8_simple': dict(n=8, lattices=1, boxsize=64.0, imprinting='simple'),
    'ic_2lpt_sc_G12': dict(n=12, lattices=1, boxsize=100.0, lpt=2),
    'ic_2lpt_sc_G8_dealias': dict(n=8, lattices=1, boxsize=64.0, lpt=2, dealias=True),
    'ic_1lpt_sc_G16_seeds': dict(n=16, lattices=1, boxsize=128.0, seeds=(11, 22)),
    'ic_3lpt_sc_G8': dict(n=8, lattices=1, boxsize=64.0, lpt=3),
    'ic_3lpt_sc_G8_dealias_backscale': dict(n=8, lattices=1, boxsize=64.0, lpt=3, dealias=True, backscale=True),
    'ic_1lpt_two_components_G6': dict(n=6, lattices=1, boxsize=48.0, components=2),
    'ic_1lpt_sc_G8_nongauss': dict(n=8, lattices=1, boxsize=64.0, nongauss=0.6),
    'ic_2lpt_sc_G10_nongauss_backscale': dict(n=10, lattices=1, boxsize=80.0, nongauss=-0.4, backscale=True, lpt=2),
}

# analytic stand-ins for the CLASS transfer functions: T_δ(k, a) = −A_δ·a·k²/(1 + (k/k0)²)^1.1,
# T_θ(k, a) = +A_θ·a^½·k²/(1 + (k/k0)²)^1.1 — the shape only has to be smooth and k-dependent
TRANSFER = dict(A_delta=1.5e8, A_theta=0.7e8, k0=0.07)
GROWTH = dict(D1=0.0251, f1=0.993, D2=-2.7e-4, f2=1.98, D3a=5.3e-6, f3a=2.97, D3b=3.2e-6, f3b=2.96, D3c=1.1e-6, f3c=2.95)


def param_text(c):
    seeds = c.get('seeds')
    return f'''
boxsize = {c['boxsize']}*Mpc
H0 = 70*km/s/Mpc
Ωcdm = 0.25
Ωb = 0.05
a_begin = 0.02
enable_class_background = False
primordial_spectrum = {{'A_s': 2.1e-9, 'n_s': 0.96, 'α_s': 0.01, 'pivot': 0.05/Mpc}}
realization_options = {{
    'backscale': {bool(c.get('backscale', False))},
    'lpt': {c.get('lpt', 1)},
    'dealias': {bool(c.get('dealias', False))},
    'nongaussianity': {c.get('nongauss', 0)},
}}
primordial_amplitude_fixed = {bool(c.get('fixed', False))}
primordial_phase_shift = {c.get('phase_shift', 0)}
primordial_noise_imprinting = '{c.get('imprinting', 'distributed')}'
''' + (f"random_seeds = {{'primordial amplitudes': {seeds[0]}, 'primordial phases': {seeds[1]}}}\n" if seeds else '')


def worker(name):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    c = CASES[name]
    ref_sandbox.enter_reference(SANDBOX, param_text(c), jobid=abs(hash(name)) % 100000 + 1)
    import commons
    from commons import universals, boxsize, a_begin
    import species, ic, mesh
    from integration import hubble

    class Spline:
        def __init__(self, f):
            self.f = f

        def eval(self, k):
            return self.f(k)

    calls = []

    def compute_transfer(component, variable, gridsize, specific_multi_index=None, a=-1, a_next=-1,
                         gauge='N-body', get='spline', weight=None, backscale=False):
        calls.append((int(variable), int(gridsize), float(a), bool(backscale)))
        shape = lambda k: k**2/(1 + (k/TRANSFER['k0'])**2)**1.1
        if variable == 0:
            return Spline(lambda k: -TRANSFER['A_delta']*a*shape(k)), None
        return Spline(lambda k: +TRANSFER['A_theta']*a**0.5*shape(k)), None

    class Cosmo:
        def __getattr__(self, key):
            if key.startswith('growth_fac_'):
                return lambda a, v=GROWTH[key[len('growth_fac_'):]]: v
            raise AttributeError(key)

    ic.compute_transfer = compute_transfer
    ic.compute_cosmo = lambda *args, **kwargs: Cosmo()
    amplitudes = {}
    get_amplitudes = ic.get_amplitudes

    def get_amplitudes_tap(gridsize, component, a, a_next=-1, variable=-1, multi_index=None, factor=1):
        out = get_amplitudes(gridsize, component, a, a_next, variable, multi_index, factor)
        amplitudes[int(variable)] = np.asarray(out).copy()
        return out
    ic.get_amplitudes = get_amplitudes_tap

    n, nl = c['n'], c['lattices']
    a = float(a_begin)
    universals.a = a
    extra = {}
    if c.get('components', 1) == 2:
        # two simple-cubic components are interleaved into one bcc lattice (ic.py:1248-1251); ids continue
        comps = [species.Component('cdm', 'cold dark matter', N=n**3), species.Component('baryons', 'baryons', N=n**3)]
        for other in comps:
            ic.realize_particles(other, a)
        comp = comps[0]
        for q, other in enumerate(comps):
            extra[f'pos_{q}'] = np.asarray(other.pos_mv).copy().reshape(-1, 3)
            extra[f'mom_{q}'] = np.asarray(other.mom_mv).copy().reshape(-1, 3)
            extra[f'mass_{q}'] = float(other.mass)
        extra['component_names'] = np.array(['cdm', 'baryons'])
        extra['component_species'] = np.array(['cold dark matter', 'baryons'])
    else:
        comp = species.Component('matter', 'matter', N=nl*n**3)
        assert comp.preic_lattice == {1: 'sc', 2: 'bcc', 4: 'fcc'}[nl]
        ic.realize_particles(comp, a)
    pos = np.asarray(comp.pos_mv).copy().reshape(-1, 3)
    mom = np.asarray(comp.mom_mv).copy().reshape(-1, 3)
    # the primordial noise on its own, in the reference's transposed Fourier layout [j][i][2·kk(+1)]
    noise = np.zeros((n, n, n + 2))
    ic.generate_primordial_noise(noise, comp.realization_options['fixedamplitude'], comp.realization_options['phaseshift'])
    out = dict(n=n, lattices=nl, boxsize=float(boxsize), a=a, H=float(hubble(a)), mass=float(comp.mass),
               varrho_bar=float(comp.ϱ_bar), w_eff=float(comp.w_eff(a=a)),
               backscale=bool(c.get('backscale', False)), lpt=int(c.get('lpt', 1)), dealias=bool(c.get('dealias', False)),
               nongaussianity=float(c.get('nongauss', 0)), fixed=bool(c.get('fixed', False)), phase_shift=float(comp.realization_options['phaseshift']),
               imprinting=str(c.get('imprinting', 'distributed')),
               seeds=np.array([commons.random_seeds['primordial amplitudes'], commons.random_seeds['primordial phases']]),
               noise=noise, pos=pos, mom=mom,
               transfer=np.array([TRANSFER['A_delta'], TRANSFER['A_theta'], TRANSFER['k0']]),
               growth_keys=np.array(list(GROWTH)), growth_vals=np.array(list(GROWTH.values())),
               A_s=2.1e-9, n_s=0.96, alpha_s=0.01, pivot=0.05,
               transfer_calls=np.array(calls))
    for variable, table in amplitudes.items():
        out[f'amplitudes{variable}'] = table
    out.update(extra)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'ok: N', len(pos), 'pos[0]', pos[0], 'mom[0]', mom[0], 'calls', calls)


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
        print(f'[{n}] rc={p.returncode}\n' + '\n'.join(o.strip().split('\n')[-4:]))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--worker':
        worker(sys.argv[2])
    else:
        main()
