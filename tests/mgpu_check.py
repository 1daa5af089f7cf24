"""Multi-GPU parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank builds the same seeded particle set, keeps its x-slab, and the distributed path
(deposit halo +=, 2-D FFT → all-to-all → 1-D FFT, k-space in the transposed layout, inverse, halo =,
gather/kick, drift, slab migration) is compared with the numpy oracle on rank 0, particle by particle
(matched through ids)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concept_b200 import commons, communication, interactions, mesh  # noqa: E402
from concept_b200.species import Component  # noqa: E402
from oracle import pm_oracle as O  # noqa: E402


def relerr(a, b):
    return float(np.max(np.abs(a - b))/np.max(np.abs(b)))


def main():
    rank, P = communication.init()
    L, N = 60.0, 40000
    failures = []
    # MGPU_ONLY=pm,mixed,powerspec,p3m restricts the run to some sections (development aid; default: all)
    only = {w for w in os.environ.get('MGPU_ONLY', '').split(',') if w}
    want = lambda section: not only or section in only
    # G = 128: hand-written slab transform, x-solve over CUDA-IPC peer pointers; G = 64: cuFFT 2-D + the
    # first-generation x-solve over peer pointers; G = 48: cuFFT + NCCL all-to-all transpose
    for G, order, diff, interlace, dtype in [(128, 2, 2, False, 'f64'), (128, 3, 4, False, 'f64'), (64, 2, 2, False, 'f64'), (64, 3, 4, False, 'f64'), (64, 4, 8, False, 'f64'),
                                             (64, 2, 0, False, 'f64'), (64, 3, 2, True, 'f64'),
                                             (48, 2, 2, False, 'f64'), (48, 3, 4, False, 'f64')]:
        if G % P or G//P < 7 or not want('pm'):
            continue
        interp = {2: 'CIC', 3: 'TSC', 4: 'PCS'}[order]
        commons.load_params(f'''
boxsize = {L}*Mpc
potential_options = {{'gridsize': {{'gravity': {{'pm': {G}}}}}, 'interpolation': {{'gravity': {{'pm': '{interp}'}}}},
                     'interlace': {{'gravity': {{'pm': ({interlace}, {interlace})}}}},
                     'differentiation': {{'default': {{'gravity': {{'pm': {diff if diff else "'fourier'"}, 'p3m': 4}}}}}}}}
select_forces = {{'matter': {{'gravity': 'pm'}}}}
''')
        commons.universals.a = 0.5
        rng = np.random.default_rng(100 + order)
        pos = rng.random((N, 3))*L
        pos[:8000, 0] = (rng.random(8000)*0.02 + np.repeat(np.arange(8)/8, 1000))*L % L   # crowd the slab faces
        mom = rng.standard_normal((N, 3))*3.0
        mass = 2.0
        c = Component('matter', 'matter', N=N, mass=mass)
        c.set_particles(pos, mom)
        n_locals = communication.allgather(c.N_local)
        assert sum(n_locals) == N, n_locals
        ᔑdt = {'1': 0.02, ('a**(-3*w_eff-1)', 'matter'): 0.041, ('a**(-3*w_eff)', 'matter'): 0.0199, 'a**(-2)': 0.08}
        interactions.gravity('pm', [c], [c], ᔑdt, 'long-range', False)
        got_pos, got_mom = c.gather_global()
        fused = mesh.get_context(G).fused_solve_available
        ids_before = set(c.ids[:c.N_local].cpu().numpy().tolist())
        c.drift(ᔑdt)                      # moves particles up to several cells → migration
        ids_now = c.ids[:c.N_local].cpu().numpy()
        stayers = ids_now[np.isin(ids_now, np.fromiter(ids_before, dtype=np.int64))]
        order_kept = communication.allgather(bool(np.all(np.diff(stayers) > 0)))
        n_after = communication.allgather(c.N_local)
        drift_pos, drift_mom = c.gather_global()
        # ownership after migration
        own = communication.slab_owner(c.pos_local[:, 0], L, G)
        ok_owner = bool((own == rank).all().item())
        oks = communication.allgather(ok_owner)
        if rank == 0:
            ref = O.pm_kick(pos, mom, mass=mass, boxsize=L, gridsize=G, order=order, G_Newton=commons.G_Newton,
                            dt_rho_over_dt1=0.041/0.02, dt_kick=0.0199, diff_order=diff, interlace=interlace)
            e1 = relerr(got_mom - mom, ref - mom)
            ref_pos = O.drift(pos, ref, 0.08*1.0/mass, L)
            e2 = float(np.max(np.abs(drift_pos - O.drift(pos, got_mom, 0.08/mass, L))))
            # (arrivals fill the holes the movers leave, like the reference's exchange: the stayers keep their places, and
            # only when more particles leave than arrive are a few moved from the tail — `order kept` is informational)
            status = 'ok' if (e1 < 1e-9 and e2 == 0.0 and all(oks) and sum(n_after) == N
                              and np.array_equal(got_pos, pos)) else 'FAIL'
            print(f'[P={P}] G={G} fused={fused} order={order} diff={diff} interlace={interlace}: kick relerr {e1:.2e}, '
                  f'drift max|Δ| {e2:.1e}, owners ok {all(oks)}, order kept {all(order_kept)}, N {n_locals}->{n_after}  {status}', flush=True)
            if status != 'ok':
                failures.append((G, order, diff, interlace))
        mesh.free_contexts()
    # ---- component-specific upstream / downstream grid sizes on several ranks: every grid is slab-decomposed and
    # pm_fourier_copy_modes moves the rows of the shared mode cube between the ranks (copy_modes' subslab exchange,
    # mesh.py:1105-1230); compared with the oracle restatement of particle_mesh on rank 0
    for Gg, grids, order, diff, interlace in [(128, {'cdm': (64, 128), 'baryons': (128, 64)}, 3, 4, False),
                                              (64, {'cdm': (32, 64), 'baryons': (64, 32)}, 2, 0, True)]:
        all_grids = {Gg} | {g for pair in grids.values() for g in pair}
        if any(g % P or g//P < 8 for g in all_grids) or not want('mixed'):
            continue
        interp = {2: 'CIC', 3: 'TSC', 4: 'PCS'}[order]
        grids_txt = ''.join(f"        '{name}': {{'gravity': {{'pm': {pair}}}}},\n" for name, pair in grids.items())
        commons.load_params(f"""
boxsize = {L}*Mpc
potential_options = {{
    'gridsize': {{
        'global': {{'gravity': {{'pm': {Gg}}}}},
{grids_txt}    }},
    'interpolation': {{'gravity': {{'pm': '{interp}'}}}},
    'interlace': {{'gravity': {{'pm': ({interlace}, {interlace})}}}},
    'differentiation': {{'default': {{'gravity': {{'pm': {diff if diff else "'fourier'"}}}}}}},
}}
select_forces = {{'all': {{'gravity': 'pm'}}}}
Ωcdm = 0.25
Ωb = 0.05
""")
        commons.universals.a = 0.5
        rng = np.random.default_rng(500 + order)
        species = {'cdm': 'cold dark matter', 'baryons': 'baryons'}
        comps, host, ᔑdt = [], [], {'1': 0.0123}
        for name, Nc, mass, dt_rho, dt_kick in [('cdm', 9000, 2.9, 0.024846, 0.012177), ('baryons', 7000, 4.2, 0.025338, 0.011808)]:
            pos = rng.random((Nc, 3))*L
            pos[:1600, 0] = (rng.random(1600)*0.02 + np.repeat(np.arange(8)/8, 200))*L % L   # crowd the slab faces
            mom = rng.standard_normal((Nc, 3))*3.0
            c = Component(name, species[name], N=Nc, mass=mass)
            c.set_particles(pos, mom)
            assert c.potential_gridsizes['gravity']['pm'] == grids[name], c.potential_gridsizes
            ᔑdt['a**(-3*w_eff-1)', name] = dt_rho
            ᔑdt['a**(-3*w_eff)', name] = dt_kick
            comps.append(c)
            host.append(dict(pos=pos, mom=mom, mass=mass, upstream=grids[name][0], downstream=grids[name][1], dt_rho=dt_rho,
                             dt_kick=dt_kick, diff_order=diff))
        interactions.gravity('pm', comps, comps, ᔑdt, 'long-range', False)
        got = [c.gather_global()[1] for c in comps]
        for g in all_grids:
            mesh.get_context(g, 'f64').check_async_error()
        if rank == 0:
            ref = O.pm_kick_multigrid(host, boxsize=L, gridsize_global=Gg, order=order, G_Newton=commons.G_Newton, dt1=0.0123,
                                      deconvolve=True, interlace=interlace)
            errs = [relerr(m - h['mom'], r - h['mom']) for m, r, h in zip(got, ref, host)]
            status = 'ok' if max(errs) < 1e-9 else 'FAIL'
            print(f'[P={P}] mixed grid sizes global {Gg}, {grids}, order={order} diff={diff} interlace={interlace}: '
                  f'kick relerr {max(errs):.2e}  {status}', flush=True)
            if status != 'ok':
                failures.append(('mixed', Gg))
        mesh.free_contexts()
    # ---- P(k) estimator on several ranks (compute_powerspec, analysis.py:500-579): slab-decomposed PCS deposit on two
    # interlaced lattices, distributed FFT, Σ|δ̂|² per k² all-reduced over the ranks; against the oracle estimator
    from concept_b200 import analysis
    Gk, Lk, Nk = 128, 500.0, 60_000
    if Gk % P == 0 and Gk//P >= 8 and want('powerspec'):
        commons.load_params(f'boxsize = {Lk}*Mpc\nH0 = 70*km/s/Mpc\nΩcdm = 0.25\nΩb = 0.05\n')
        commons.universals.a = 0.5
        rng = np.random.default_rng(77)
        pos = rng.random((Nk, 3))*Lk
        centres = rng.random((40, 3))*Lk
        pos[:Nk//2] = (centres[rng.integers(0, 40, Nk//2)] + rng.standard_normal((Nk//2, 3))*0.01*Lk) % Lk
        c = Component('matter', 'matter', N=Nk, mass=1.7)
        c.set_particles(pos, np.zeros((Nk, 3)))
        centers, power, n_modes = analysis.powerspec([c], Gk)
        mesh.get_context(Gk, 'f64').check_async_error()
        mesh.free_contexts()
        if rank == 0:
            k2_max, _ = analysis.get_powerspec_bins(Gk, boxsize=Lk)
            slab = O.density_fourier(pos, 1.7, 0.5, Lk, Gk, 4, True, True)
            power_k2, count_k2 = O.power_by_k2(slab, k2_max)
            _, idx, centers_ref, n_ref = analysis.get_powerspec_bins(Gk, n_modes_fine=count_k2, boxsize=Lk)
            ref = np.zeros(len(centers_ref))
            np.add.at(ref, idx, power_k2)
            ref *= (0.5**(-3)*c.ϱ_bar)**(-2)*Lk**3/n_ref
            e = float(np.max(np.abs(power/ref - 1)))
            status = 'ok' if (np.array_equal(n_modes, n_ref) and e < 1e-9) else 'FAIL'
            print(f'[P={P}] P(k) G={Gk}: mode counts equal {bool(np.array_equal(n_modes, n_ref))}, max |P/P_ref - 1| {e:.2e}  {status}', flush=True)
            if status != 'ok':
                failures.append(('powerspec', Gk))
    # ---- P³M short range on several ranks: ghost particles across the slab faces (and across the periodic boundary),
    # receivers on active rungs only, Δmom / rung indices migrating with their particles
    import ctypes
    from concept_b200 import shortrange
    from concept_b200._lib import check
    Lp, Gp, Np = 60.0, 128, 12000
    if Gp % P == 0 and Lp/P >= 2*4.5*1.25*Lp/Gp and want('p3m'):
        commons.load_params(f"""
boxsize = {Lp}*Mpc
potential_options = {{'gridsize': {{'gravity': {{'p3m': {Gp}}}}}}}
select_forces = {{'matter': {{'gravity': 'p3m'}}}}
""")
        commons.universals.a = 0.5
        rng = np.random.default_rng(77)
        pos = rng.random((Np, 3))*Lp
        pos[:3000, 0] = (rng.random(3000)*0.03 + np.repeat(np.arange(8)/8, 375) - 0.015) % 1.0*Lp     # crowd the slab faces
        pos[3000:4000] = (pos[4000:5000] + 0.05*rng.standard_normal((1000, 3))) % Lp                    # close pairs
        c = Component('matter', 'matter', N=Np, mass=2.0)
        c.set_particles(pos, np.zeros((Np, 3)))
        shortrange.ensure_rung_state(c)
        rung_of = lambda ids: ((ids*7919) % 5).to(torch.int8)
        n = c.N_local
        c.rung_indices[:n] = rung_of(c.ids[:n])
        jumped_of = lambda ids: (rung_of(ids) + 8*((ids % 7) == 0).to(torch.int8) + 16*((ids % 7) == 3).to(torch.int8)).to(torch.int8)
        c.rung_indices_jumped[:n] = jumped_of(c.ids[:n])
        factors = np.linspace(1.0, 3.3, 23)
        table, maxr2, rng_sr, size = shortrange.get_shortrange_table(Gp, c.softening_length, c.device)
        ctx = c._pm_context()
        check(ctx.lib.pm_shortrange(ctx._h, c.pos.data_ptr(), n, c.rung_indices.data_ptr(), c.rung_indices_jumped.data_ptr(), 2,
                                    factors.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 23, rng_sr, table.data_ptr(), size, maxr2,
                                    c.Δmom.data_ptr()))
        ctx.check_async_error()
        parts = communication.allgather((c.Δmom[:n].cpu().numpy(), c.ids[:n].cpu().numpy()))
        # migration with the P³M state: move everything by a few cells, then every particle must still carry its own Δmom / rungs
        c.mom[:n] = torch.as_tensor(rng.standard_normal((Np, 3)), device=c.device)[c.ids[:n]]*40.0
        c.drift({'a**(-2)': 0.08})
        n2 = c.N_local
        ids2 = c.ids[:n2]
        state_ok = bool(torch.equal(c.rung_indices[:n2], rung_of(ids2)) and torch.equal(c.rung_indices_jumped[:n2], jumped_of(ids2)))
        dm_all = np.zeros((Np, 3)); 
        for dm, ii in parts:
            dm_all[ii] = dm
        state_ok = state_ok and bool(np.array_equal(c.Δmom[:n2].cpu().numpy(), dm_all[ids2.cpu().numpy()]))
        oks = communication.allgather((state_ok, n2))
        if rank == 0:
            rung_h = ((np.arange(Np)*7919) % 5).astype(np.int8)
            jumped_h = (rung_h + 8*((np.arange(Np) % 7) == 0) + 16*((np.arange(Np) % 7) == 3)).astype(np.int64)
            tab_o, maxr2_o = O.shortrange_table(*shortrange.shortrange_params(Gp)[:2], size, c.softening_length)
            active = rung_h >= 2
            ref = O.shortrange_sums(pos, Lp, rng_sr, tab_o, maxr2_o, active=active)*factors[jumped_h][:, None]
            e = relerr(dm_all[active], ref[active])
            status = 'ok' if (e < 1e-11 and np.all(dm_all[~active] == 0) and all(o[0] for o in oks) and sum(o[1] for o in oks) == Np) else 'FAIL'
            print(f'[P={P}] P3M short range G={Gp}: pair-kick relerr {e:.2e}, inactive untouched {bool(np.all(dm_all[~active] == 0))}, '
                  f'rung state follows migration {all(o[0] for o in oks)}, N {sum(o[1] for o in oks)}  {status}', flush=True)
            if status != 'ok':
                failures.append(('p3m', Gp))
        mesh.free_contexts()
    failures = communication.bcast(failures)
    communication.barrier()
    if rank == 0:
        print('MGPU_CHECK', 'PASSED' if not failures else f'FAILED {failures}', flush=True)
    if P > 1:
        torch.distributed.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == '__main__':
    main()
