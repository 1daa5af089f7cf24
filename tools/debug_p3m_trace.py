import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200 import commons, main, mesh, interactions, shortrange
from concept_b200.species import Component
d = np.load('tests/golden/run_p3m_8.npz')
commons.load_params('''
boxsize = 8*Mpc
potential_options = {'gridsize': {'gravity': {'p3m': 24}}}
H0      = 70*km/s/Mpc
Ωcdm    = 0.25
Ωb      = 0.05
a_begin = 0.02
output_times = {'snapshot': (0.0245,)}
select_forces = {'matter': {'gravity': 'p3m'}}
''')
c = Component('matter', 'matter', N=512, mass=float(d['mass']))
c.populate(d['pos0'], 'pos'); c.populate(d['mom0'], 'mom')
trace = []
cnt = {'short': 0}
orig = interactions.gravity
def tap(method, receivers, suppliers, ᔑdt, interaction_type, printout=True):
    res = orig(method, receivers, suppliers, ᔑdt, interaction_type, printout)
    if 'long' in interaction_type:
        k = len(trace)
        pos, mom = c.pos_mv3, c.mom_mv3
        L = 8.0
        if k < len(d['trace_t']):
            dx = pos - d['trace_pos'][k]; dx -= L*np.round(dx/L)
            print(f'kick {k}: t={commons.universals.t:.8f} ref t={d["trace_t"][k]:.8f} n_short={cnt["short"]} ref={d["trace_n_short"][k]} '
                  f'max|dx|={np.abs(dx).max():.3e} mom relerr={np.abs(mom-d["trace_mom"][k]).max()/np.abs(d["trace_mom"][k]).max():.3e} rungs={c.rungs_N}')
        trace.append(1)
    else:
        k = cnt['short']
        cnt['short'] += 1
        if k < len(d['strace_t']):
            N = 512
            pos = c.pos_mv3
            dm = c.Δmom[:N].cpu().numpy()
            rung = c.rung_indices[:N].cpu().numpy(); jumped = c.rung_indices_jumped[:N].cpu().numpy()
            dx = pos - d['strace_pos'][k]; dx -= 8.0*np.round(dx/8.0)
            ref_dm = d['strace_dmom'][k]
            bad = np.argmax(np.abs(dm - ref_dm).sum(1))
            print(f'short {k}: lowest_active={c.lowest_active_rung} ref={d["strace_lowest_active"][k]} max|dx|={np.abs(dx).max():.2e} '
                  f'dmom err={np.abs(dm-ref_dm).max()/max(np.abs(ref_dm).max(),1e-300):.2e} rung mismatches={(rung!=d["strace_rung"][k]).sum()} '
                  f'jumped mismatches={(jumped!=d["strace_jumped"][k]).sum()} worst i={bad} rung={rung[bad]} jumped={jumped[bad]} ref jumped={d["strace_jumped"][k][bad]} dm={dm[bad]} ref={ref_dm[bad]}')
    return res
interactions.gravity = tap
main.timeloop([c])
