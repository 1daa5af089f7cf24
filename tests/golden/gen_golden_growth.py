"""Golden values of the growth factors D1, f1, D2, f2, D3a … f3c: the UNMODIFIED reference's own matter + Λ
background (integration.init_time → solve_matterΛ_background, integration.py:1043-1188; growth ODEs :1104-1290)
in its pure-Python mode under oracle/ref_sandbox.py, evaluated through its temporal splines.

Run in the build container only:   python tests/golden/gen_golden_growth.py
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'oracle'))
import ref_sandbox
if not os.path.isdir('/tmp/concept_ref_sandbox/src'):
    ref_sandbox.build_sandbox('/tmp/concept_ref_sandbox')
ref_sandbox.enter_reference('/tmp/concept_ref_sandbox', '''
boxsize = 64*Mpc
H0 = 67*km/(s*Mpc)
Ωb = 0.049
Ωcdm = 0.27
a_begin = 0.02
enable_class_background = False
''', jobid=88123)
import commons
from commons import *
import integration
from integration import init_time, temporal_splines
init_time()
print([k for k in dir(temporal_splines) if k.startswith('a_')])
avals = np.array([0.02, 0.05, 0.1, 0.3, 0.5, 0.8, 1.0])
out = {'a': avals, 'H0': float(H0), 'Om': float(Ωm)}
for key in ('D', 'f', 'D2', 'f2', 'D3a', 'f3a', 'D3b', 'f3b', 'D3c', 'f3c'):
    sp = getattr(temporal_splines, 'a_' + key)
    out[key] = np.array([sp.eval(a) for a in avals])
    print(key, out[key][:3])
np.savez(os.path.join(HERE, 'growth_factors.npz'), **out)
