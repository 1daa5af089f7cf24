"""Golden vectors for the P³M short-range path, from the unmodified reference in pure-Python mode.

  shortkick_p3m_G24.npz  one fake short kick (rung assignment, main.py:1639-1675 → kick_short fake=True,
                         main.py:1173-1238) followed by a real half short-range kick (main.kick_short),
                         with Δmom→acc conversion (species.py:2290-2325): rung_indices, acc, mom.
  run_p3m_8.npz          a short full P³M run (8³ particles, 24³ grid — tiling needs ≥ 4 tiles across the
                         box ⇒ G ≥ 23, species.py:3971) with the whole rung machinery
                         (driftkick_short, main.py:1347-1624): ICs, final pos/mom, Δt history.

    python tests/golden/gen_golden_p3m.py [shortkick|run]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

PARAM = '''
initial_conditions = 'injected'
output_dirs        = {{'snapshot': param.dir + '/output'}}
output_times       = {{'snapshot': ({a_end},)}}
boxsize = 8*Mpc
potential_options = {{'gridsize': {{'gravity': {{'p3m': 24}}}}}}
H0      = 70*km/s/Mpc
Ωcdm    = 0.25
Ωb      = 0.05
a_begin = 0.02
select_forces = {{'matter': {{'gravity': 'p3m'}}}}
enable_class_background = False
'''


def make_ics(np, L, N, rho_crit, Om):
    rng = np.random.Generator(np.random.PCG64DXSM(77))
    n = round(N**(1/3))
    idx = (np.arange(n) + 0.5)*(L/n)
    q = np.stack(np.meshgrid(idx, idx, idx, indexing='ij'), -1).reshape(-1, 3)
    pos0 = np.mod(q + 0.15*(L/n)*rng.standard_normal((N, 3)), L)
    # a few tight pairs/clumps so that several rungs get populated
    pos0[:40] = np.mod(pos0[40:80] + 0.01*(L/n)*rng.standard_normal((40, 3)), L)
    pos0[80:100] = np.mod(pos0[100] + 0.003*(L/n)*rng.standard_normal((20, 3)), L)
    mass = float(Om*rho_crit*L**3/N)
    mom0 = 1e-3*mass*rng.standard_normal((N, 3))*(L/n)
    return pos0, mom0, mass


def worker(kind):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_sandbox
    sandbox = f'/tmp/concept_ref_sandbox_p3m_{kind}'
    ref_sandbox.build_sandbox(sandbox)
    ref_sandbox.enter_reference(sandbox, PARAM.format(a_end=0.0245 if kind == 'run' else 1.0), jobid=9)
    import commons
    from commons import universals, boxsize, G_Newton, Ωb, Ωcdm, ρ_crit, N_rungs, shortrange_params, softening_kernel
    import species, snapshot, interactions
    N, L = 8**3, float(boxsize)
    pos0, mom0, mass = make_ics(np, L, N, float(ρ_crit), float(Ωb + Ωcdm))
    out = dict(pos0=pos0, mom0=mom0, mass=mass, boxsize=L, G_Newton=float(G_Newton), gridsize=24, N_rungs=int(N_rungs),
               sr_scale=float(shortrange_params['gravity']['scale']), sr_range=float(shortrange_params['gravity']['range']),
               sr_tablesize=int(shortrange_params['gravity']['tablesize']), softening_kernel=str(softening_kernel), a_begin=0.02)

    def make_component():
        comp = species.Component('matter', 'matter', N=N, mass=mass)
        for d, s in enumerate('xyz'):
            comp.populate(np.ascontiguousarray(pos0[:, d]), 'pos' + s)
            comp.populate(np.ascontiguousarray(mom0[:, d]), 'mom' + s)
        return comp
    if kind == 'shortkick':
        commons.jobid = -1          # import main without launching a run (main.py:2437)
        import main
        import integration
        integration.init_time()
        comp = make_component()
        out['softening_length'] = float(comp.softening_length)
        Δt = 0.004
        out['dt'] = Δt
        out['t0'] = float(universals.t)
        out['a0'] = float(universals.a)
        main.get_time_step_integrals(0, 0, [comp])
        main.initialize_rung_populations([comp], Δt)         # fake kick + assign_rungs
        out['rung_indices_init'] = np.array(comp.rung_indices_mv[:N]).copy()
        out['acc_init'] = np.array(comp.Δmom_mv3[:N]).copy()
        out['mom_after_fake'] = np.array(comp.mom_mv3[:N]).copy()
        main.kick_short([comp], Δt)                           # real half kick on every rung
        out['mom_after_kick'] = np.array(comp.mom_mv3[:N]).copy()
        out['acc_after_kick'] = np.array(comp.Δmom_mv3[:N]).copy()
        out['dt_rungs_1'] = np.array(main.ᔑdt_rungs['1']).copy()
        out['dt_rungs_pair'] = np.array(main.ᔑdt_rungs['a**(-3*w_eff₀-3*w_eff₁-1)', 'matter', 'matter']).copy()
        out['dt_rungs_a2'] = np.array(main.ᔑdt_rungs['a**2']).copy()
        np.savez_compressed(os.path.join(HERE, 'shortkick_p3m_G24.npz'), **out)
        print('shortkick ok; rung populations', np.bincount(out['rung_indices_init'], minlength=8))
        return
    # full run
    snaps, log = {}, dict(step_t=[], step_a=[])

    def get_ic(*a, **k):
        return [make_component()]

    def save(components, filename, *a, **k):
        c = components[0]
        snaps['final'] = (np.array(c.pos_mv3[:N]).copy(), np.array(c.mom_mv3[:N]).copy(), float(universals.t),
                          float(universals.a), np.array(c.rung_indices_mv[:N]).copy())
        return filename
    snapshot.get_initial_conditions = get_ic
    snapshot.save = save
    trace = dict(pos=[], mom=[], t=[], n_short=[], kind=[])
    counter = dict(short=0)
    strace = dict(pos=[], dmom=[], rung=[], jumped=[], lowest_active=[], t=[])
    orig_gravity = interactions.gravity

    def gravity_tap(method, receivers, suppliers, ᔑdt, interaction_type, printout):
        res = orig_gravity(method, receivers, suppliers, ᔑdt, interaction_type, printout)
        c = receivers[0]
        if 'long' in interaction_type:
            trace['pos'].append(np.array(c.pos_mv3[:N]).copy())
            trace['mom'].append(np.array(c.mom_mv3[:N]).copy())
            trace['t'].append(float(universals.t))
            trace['n_short'].append(counter['short'])
        else:
            counter['short'] += 1
            if counter['short'] <= 14:
                strace['pos'].append(np.array(c.pos_mv3[:N]).copy())
                strace['dmom'].append(np.array(c.Δmom_mv3[:N]).copy())     # raw Δmom of this kick (before apply/convert)
                strace['rung'].append(np.array(c.rung_indices_mv[:N]).copy())
                strace['jumped'].append(np.array(c.rung_indices_jumped_mv[:N]).copy())
                strace['lowest_active'].append(int(c.lowest_active_rung))
                strace['t'].append(float(universals.t))
        return res
    interactions.gravity = gravity_tap
    try:
        import main  # noqa: F401
    except SystemExit as e:
        print('reference exited with', e.code)
    p, m, t, a, r = snaps['final']
    out.update(pos_final=p, mom_final=m, t_final=t, a_final=a, rung_final=r)
    out.update(trace_pos=np.asarray(trace['pos']), trace_mom=np.asarray(trace['mom']), trace_t=np.asarray(trace['t']),
               trace_n_short=np.asarray(trace['n_short']))
    out.update({f'strace_{k}': np.asarray(v) for k, v in strace.items()})
    np.savez_compressed(os.path.join(HERE, 'run_p3m_8.npz'), **out)
    print('run ok; final a', a, 'rungs', np.bincount(r, minlength=8))


if __name__ == '__main__':
    worker(sys.argv[1] if len(sys.argv) > 1 else 'shortkick')
