"""Golden GADGET-2 snapshot: the UNMODIFIED reference (snapshot.save with snapshot_type = 'gadget',
snapshot.py:640-2640) run in its pure-Python mode under oracle/ref_sandbox.py writes a file; the input
particle arrays and the bytes of the file are stored in tests/golden/snapshot_gadget_*.npz.

Run in the build container only:   python tests/golden/gen_golden_snapshot.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SANDBOX = '/tmp/concept_ref_sandbox'

CASES = {
    'snapshot_gadget_a05': dict(boxsize=64.0, N=300, seed=41, a=0.5, mass=2.5, H0=70.0, Ωcdm=0.25, Ωb=0.05),
    'snapshot_gadget_a1': dict(boxsize=256.0, N=777, seed=42, a=1.0, mass=0.37, H0=67.0, Ωcdm=0.27, Ωb=0.049),
}


def worker(name):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    c = CASES[name]
    param = f'''
boxsize = {c['boxsize']}*Mpc
H0 = {c['H0']}*km/s/Mpc
Ωcdm = {c['Ωcdm']}
Ωb = {c['Ωb']}
a_begin = 0.02
enable_class_background = False
snapshot_type = 'gadget'
'''
    ref_sandbox.enter_reference(SANDBOX, param, jobid=abs(hash(name)) % 100000 + 1)
    import commons
    from commons import universals, boxsize
    import species, snapshot
    rng = np.random.Generator(np.random.PCG64DXSM(c['seed']))
    N, L = c['N'], float(boxsize)
    pos = rng.random((N, 3))*L
    pos[:5] = 0.0                       # particles at the origin
    mom = rng.standard_normal((N, 3))*c['mass']*30
    universals.a = c['a']
    universals.t = 3.0
    comp = species.Component('matter', 'matter', N=N, mass=c['mass'])
    for d, s in enumerate('xyz'):
        comp.populate(np.ascontiguousarray(pos[:, d]), 'pos' + s)
        comp.populate(np.ascontiguousarray(mom[:, d]), 'mom' + s)
    fn = snapshot.save(comp, f'/tmp/{name}')
    data = np.frombuffer(open(fn, 'rb').read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), pos=pos, mom=mom, mass=c['mass'], a=c['a'], boxsize=L,
                        H0=float(commons.H0), Omega_m=float(commons.Ωm), file_bytes=data)
    print(name, 'ok', len(data), 'bytes')


def main():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    if not os.path.isdir(SANDBOX + '/src'):
        ref_sandbox.build_sandbox(SANDBOX)
    for n in sys.argv[1:] or list(CASES):
        p = subprocess.run([sys.executable, __file__, '--worker', n], capture_output=True, text=True)
        print(f'[{n}] rc={p.returncode}\n' + '\n'.join((p.stdout + p.stderr).strip().split('\n')[-3:]))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--worker':
        worker(sys.argv[2])
    else:
        main()
