import os, sys
import numpy as np
from scipy.optimize import linear_sum_assignment
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200 import commons, main, interactions
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
L = 8.0
def compare(tag, pos, ref_pos, mom=None, ref_mom=None):
    diff = pos[:, None, :] - ref_pos[None, :, :]
    diff -= L*np.round(diff/L)
    dist = np.sqrt((diff**2).sum(-1))
    r, cidx = linear_sum_assignment(dist)
    dd = dist[r, cidx]
    msg = f'{tag}: matched dist median {np.median(dd):.2e} p99 {np.percentile(dd, 99):.2e} max {dd.max():.2e} n>1e-6: {(dd > 1e-6).sum()}'
    if mom is not None:
        dm = np.abs(mom - ref_mom[cidx]).max(1)/np.abs(ref_mom).max()
        msg += f' | mom err median {np.median(dm):.2e} max {dm.max():.2e} n>1e-6: {(dm>1e-6).sum()}'
    print(msg)
k = [0]
ks = [0]
orig = interactions.gravity
def tap(method, receivers, suppliers, ᔑdt, interaction_type, printout=True):
    res = orig(method, receivers, suppliers, ᔑdt, interaction_type, printout)
    if 'short' in interaction_type:
        kk = ks[0]; ks[0] += 1
        if kk < len(d['strace_t']):
            pos = c.pos_mv3
            diff = pos[:, None, :] - d['strace_pos'][kk][None, :, :]
            diff -= L*np.round(diff/L)
            dist = np.sqrt((diff**2).sum(-1))
            r, cidx = linear_sum_assignment(dist)
            dm = c.Δmom[:512].cpu().numpy(); ref_dm = d['strace_dmom'][kk][cidx]
            rung = c.rung_indices[:512].cpu().numpy(); jumped = c.rung_indices_jumped[:512].cpu().numpy()
            err = np.abs(dm - ref_dm).max(1)/max(np.abs(ref_dm).max(), 1e-300)
            bad = np.argmax(err)
            print(f'short {kk}: active>={c.lowest_active_rung} (ref {d["strace_lowest_active"][kk]}) pos max {dist[r, cidx].max():.1e} dmom err max {err.max():.1e} n>1e-9 {(err>1e-9).sum()} '
                  f'rung mism {(rung != d["strace_rung"][kk][cidx]).sum()} jumped mism {(jumped != d["strace_jumped"][kk][cidx]).sum()} '
                  f'| worst: rung {rung[bad]} jumped {jumped[bad]} ref rung {d["strace_rung"][kk][cidx][bad]} ref jumped {d["strace_jumped"][kk][cidx][bad]} dm {dm[bad]} ref {ref_dm[bad]}')
    if 'long' in interaction_type and k[0] < 4:
        compare(f'long kick {k[0]}', c.pos_mv3, d['trace_pos'][k[0]], c.mom_mv3, d['trace_mom'][k[0]])
        k[0] += 1
    return res
interactions.gravity = tap
snaps = {}
main.timeloop([c], on_dump=lambda comps, dt: snaps.update(final=(comps[0].pos_mv3.copy(), comps[0].mom_mv3.copy())))
compare('final', snaps['final'][0], d['pos_final'], snaps['final'][1], d['mom_final'])
