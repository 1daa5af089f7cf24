#!/usr/bin/env python
"""bench.py — PM-gravity cycle benchmark (BASELINE.json: particle-updates/s and ms per PM cycle,
256³ particles / 512³ grid, CIC, fp64, on 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one PM cycle: long-range kick (zero grid → CIC deposit → r2c FFT →
Green's function/deconvolution → c2r FFT → fused gradient + gather + kick + Σmom²) followed by a
full drift and, on several GPUs, the slab migration of particles.  Strong scaling: the same
256³/512³ problem is split into x-slabs over the N ranks.

Prints ONE JSON line (rank 0):
  value        device-resident throughput of configs[1] (256³/512³, CIC, fp64)
  e2e          the same cycle through the C-ABI entry point that takes HOST buffers (pm_kick_long_host)
  roofline     the dominant hand-written kernel, timed live with CUDA events inside the timed steps,
               plus the per-kernel table and the whole-cycle fraction
  parity       BEFORE the timed region, at this N: a distributed kick + drift + migration at G = 128 and at the
               benchmark's G = 512 on 10⁵ seeded particles, compared particle by particle (ids) with the CPU
               restatement of the reference on rank 0 (oracle/, used here as the checker only); the run FAILS if the
               kick differs by more than 1e-9 of its maximum, the drift is not bit-exact or a particle is lost —
               the reference's own multi-process pin (test/nprocs_pm/analyze.py:121) made visible to the driver
  particles_after, sum_mom2   measured after exactly warmup + steps cycles (comparable across N)
  extra_configs   configs[2] (TSC, fp32 grid) at every N; configs[3] (512³/1024³) at N = 8; the P³M cycle of configs[4]
  cpu_baseline    the C/OpenMP restatement of the reference loops (oracle/pm_oracle.c) on this box's host cores

--impl reference: the reference's CPU implementation of the same cycle.  The compiled reference cannot be built
in this image (no MPI/FFTW/GSL), so this is the oracle port with all host threads (set explicitly:
torchrun exports OMP_NUM_THREADS=1) on the full 256³/512³ workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SIDE, GRID, BOXSIZE = 256, 512, 512.0
G_NEWTON = 4.4985024439973154e-05
SIGMA = 0.3
DT_OVER_MASS = 1e-4
METRIC = 'particle-updates/sec (one PM cycle: long-range kick + drift), 256^3 particles / 512^3 grid'
UNIT = 'particle-updates/s'
PARITY_TOL = 1e-9

# BASELINE.json configs
CONFIGS = {
    'config2': dict(n_side=256, grid=512, order=2, diff=2, dtype='f64', r_scale_cells=0.0,
                    label='configs[1]: 256^3 particles, 512^3 PM grid, CIC, fp64, deconvolution order 4, finite-difference order 2'),
    'config3': dict(n_side=256, grid=512, order=3, diff=2, dtype='f32', r_scale_cells=0.0,
                    label='configs[2]: 256^3 particles, 512^3 PM grid, TSC, fp32 grid (particles fp64), deconvolution order 6'),
    'config4': dict(n_side=512, grid=1024, order=2, diff=2, dtype='f64', r_scale_cells=0.0,
                    label='configs[3]: 512^3 particles, 1024^3 PM grid, CIC, fp64, 8 x-slabs'),
    'config5': dict(n_side=256, grid=512, order=2, diff=4, dtype='f64', r_scale_cells=1.25,
                    label='configs[4]: 256^3 particles, 512^3 grid, P3M: long-range PM (Gaussian split r_s = 1.25 cells, '
                          'difference order 4) + short-range pair kick within 4.5 r_s (spline softening, one rung)'),
}


def kick_kwargs(cfg, boxsize=None):
    L = BOXSIZE*cfg['grid']/GRID if boxsize is None else boxsize
    return dict(mass=1.0, boxsize=L, gridsize=cfg['grid'], order=cfg['order'], G_Newton=G_NEWTON, dt_rho_over_dt1=2.0,
                dt_kick=1e-3, diff_order=cfg['diff'], r_scale=cfg['r_scale_cells']*L/cfg['grid'])


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons during the timed region."""
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.proc = None
        self.lines = []
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_cycle_rate(cfg, n_side, grid, cycles=1, warm=1):
    """Times the C/OpenMP restatement on (n_side³, grid³); returns (updates/s, threads, seconds/cycle)."""
    from oracle import c_oracle as C
    from concept_b200.synthetic import zeldovich_particles
    threads = C.set_num_threads()          # every host core, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    L = BOXSIZE*grid/GRID
    pos, mom = zeldovich_particles(n_side, L, SIGMA, seed=0)
    pos, mom = pos.numpy(), mom.numpy()
    work = C.Workspace(grid)
    kw = dict(kick_kwargs(dict(cfg, grid=grid), boxsize=L), work=work)
    ts = []
    for it in range(warm + cycles):
        t0 = time.perf_counter()
        C.kick_long(pos, mom, **kw)
        C.sum_mom2(mom)
        C.drift(pos, mom, DT_OVER_MASS, L)
        if it >= warm:
            ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return n_side**3/sec, threads, sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import c_oracle as C
    from concept_b200.synthetic import zeldovich_particles
    cfg = CONFIGS['config2']
    threads = C.set_num_threads()
    n_side, grid = N_SIDE, GRID
    L = BOXSIZE

    def make(n_side, grid):
        L = BOXSIZE*grid/GRID
        pos, mom = zeldovich_particles(n_side, L, SIGMA, seed=0)
        return pos.numpy(), mom.numpy(), dict(kick_kwargs(dict(cfg, grid=grid), boxsize=L), work=C.Workspace(grid)), L
    pos, mom, kw, L = make(n_side, grid)

    def step():
        C.kick_long(pos, mom, **kw)
        C.sum_mom2(mom)
        C.drift(pos, mom, DT_OVER_MASS, L)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    sample = f'full workload: {n_side}^3 particles / {grid}^3 grid per step'
    if first*(args.steps + args.warmup) > 240:
        # too few host cores to finish K steps of the full workload in minutes: 1/8 of it, same particles per cell
        n_side, grid = N_SIDE//2, GRID//2
        pos, mom, kw, L = make(n_side, grid)
        sample = (f'{n_side}^3 particles / {grid}^3 grid per step (1/8 of the workload, same particles per cell; one full-size '
                  f'cycle took {first:.1f} s on {threads} threads)')
        step()
    for _ in range(max(args.warmup - 1, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sec = (time.perf_counter() - t0)/args.steps
    value = n_side**3/sec
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec*1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(cfg, args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'compiled reference unbuildable here (needs mpicc/FFTW-MPI/GSL); C/OpenMP port of its loops, pinned to '
                'golden vectors from the reference run in pure-Python mode; OpenMP threads set to os.cpu_count() explicitly',
    }))


def workload_config(cfg, n_gpus):
    return {'workload': cfg['label'] + f", Zel'dovich-displaced lattice (sigma = {SIGMA} spacings, seed 0)",
            'particles': cfg['n_side']**3, 'grid': cfg['grid'], 'interpolation': {2: 'CIC', 3: 'TSC', 4: 'PCS'}[cfg['order']],
            'grid_dtype': cfg['dtype'], 'decomposition': f'{n_gpus} x-slab(s)',
            'l2_policy': 'inputs larger than L2 (particles and grid are each several times the 126 MB L2 per GPU)'}


# ----------------------------------------------------------------------------- GPU arm
class Dist:
    """The process layout of one bench run (one rank per GPU under torchrun)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run for --gpus > 1')
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)

    def bcast(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def allgather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None]*self.world
        self.dist.all_gather_object(out, obj)
        return out

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [v.item() for v in t]

    def sum_(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [v.item() for v in t]


def make_context(D, grid, boxsize, dtype):
    from concept_b200.pmsolver import PMContext
    ctx = PMContext(grid, boxsize, dtype=dtype, rank=D.rank, nranks=D.world, device=D.local_rank)
    if D.world > 1:
        ctx.connect(D.bcast, D.allgather, D.rank == 0)
    return ctx


def distribute(D, ctx, pos, mom, ids=None, slack=1.5):
    """Every rank holds the same full particle set; keep this rank's x-slab in capacity-sized buffers."""
    torch = D.torch
    n_total = pos.shape[0]
    if D.world == 1:
        return pos, mom, ids, n_total
    G, L = ctx.gridsize, ctx.boxsize
    owner = torch.clamp((pos[:, 0]*(G/L)).to(torch.int64), 0, G - 1)//ctx.nx_local
    keep = owner == D.rank
    n_local = int(keep.sum().item())
    capacity = int(n_total/D.world*slack) + 4096
    pbuf = torch.zeros((capacity, 3), dtype=torch.float64, device=D.dev)
    mbuf = torch.zeros((capacity, 3), dtype=torch.float64, device=D.dev)
    pbuf[:n_local] = pos[keep]
    mbuf[:n_local] = mom[keep]
    ibuf = None
    if ids is not None:
        ibuf = torch.zeros(capacity, dtype=torch.int64, device=D.dev)
        ibuf[:n_local] = ids[keep]
    return pbuf, mbuf, ibuf, n_local


def parity_check(D, ctx, cfg, n_sub, seed):
    """One distributed kick + drift + migration of n_sub seeded particles on `ctx`, compared on rank 0 with the CPU
    restatement of the reference (C/OpenMP oracle, pinned to the reference's golden vectors).  Returns the record."""
    import numpy as np
    torch = D.torch
    from concept_b200.pmsolver import make_kick_params
    G, L = ctx.gridsize, ctx.boxsize
    rng = np.random.default_rng(seed)
    pos_h = rng.random((n_sub, 3))*L
    # crowd the slab faces so that halos and migration are exercised at every N
    nf = n_sub//5
    pos_h[:nf, 0] = (np.repeat(np.arange(8)/8, (nf + 7)//8)[:nf] + (rng.random(nf) - 0.5)*3.0/G) % 1.0*L
    mom_h = rng.standard_normal((n_sub, 3))
    kw = kick_kwargs(dict(cfg, grid=G), boxsize=L)
    dt = 0.7*L/G                      # |mom| ~ 1: a good fraction of a cell per step, so that particles do change slab
    pos = torch.as_tensor(pos_h, device=D.dev)
    mom = torch.as_tensor(mom_h, device=D.dev)
    ids = torch.arange(n_sub, dtype=torch.int64, device=D.dev)
    pbuf, mbuf, ibuf, n = distribute(D, ctx, pos, mom, ids, slack=3.0)
    sum2 = torch.zeros(1, dtype=torch.float64, device=D.dev)
    # calibrate the particle mass (the kick scales with mass²) so that max|Δmom| ≈ max|mom| ≈ 1: the comparison below then
    # measures the kick itself and not the rounding of mom + Δmom
    probe = torch.zeros_like(mbuf[:n])
    ctx.kick_long(pbuf[:n], probe, make_kick_params(**kw))
    dmax = D.max_([float(probe.abs().max().item()) if n else 0.0])[0]
    kw['mass'] = kw['mass']*(1.0/dmax)**0.5
    params = make_kick_params(**kw)
    ctx.kick_drift(pbuf[:n], mbuf[:n], params, dt, sum_mom2=sum2)
    if D.world > 1:
        ctx.allreduce_sum(sum2)
        n = ctx.exchange(pbuf, mbuf, ibuf, n)
    ctx.check_async_error()
    own_ok = True
    if D.world > 1 and n:
        owner = torch.clamp((pbuf[:n, 0]*(G/L)).to(torch.int64), 0, G - 1)//ctx.nx_local
        own_ok = bool((owner == D.rank).all().item())
    parts = D.allgather((pbuf[:n].cpu().numpy(), mbuf[:n].cpu().numpy(), ibuf[:n].cpu().numpy() if ibuf is not None
                         else np.arange(n), own_ok))
    rec = None
    if D.rank == 0:
        from oracle import c_oracle as C
        C.set_num_threads()
        gp = np.concatenate([p[0] for p in parts]); gm = np.concatenate([p[1] for p in parts])
        gi = np.concatenate([p[2] for p in parts])
        n_conserved = bool(gi.shape[0] == n_sub and np.array_equal(np.sort(gi), np.arange(n_sub)))
        rec = {'grid': G, 'particles': n_sub, 'n_conserved': n_conserved, 'owners_ok': all(p[3] for p in parts)}
        if n_conserved:
            o = np.argsort(gi, kind='stable')
            gp, gm = gp[o], gm[o]
            ref = C.kick_long(pos_h.copy(), mom_h.copy(), **kw)
            dmax = float(np.max(np.abs(ref - mom_h)))
            rec['kick_relerr'] = float(np.max(np.abs(gm - ref)))/dmax
            rec['max_dmom_over_max_mom'] = dmax/float(np.max(np.abs(mom_h)))
            # drift: bit-exact given the (GPU-)kicked momenta — pos = mod(pos + mom·Δ, L) has one correct rounding
            rec['drift_exact'] = bool(np.array_equal(gp, C.drift(pos_h.copy(), np.ascontiguousarray(gm), dt, L)))
            rec['sum_mom2_relerr'] = abs(float(sum2.item()) - float(np.sum(gm*gm)))/float(np.sum(gm*gm))
            owner0 = np.clip((pos_h[:, 0]*(G/L)).astype(np.int64), 0, G - 1)//ctx.nx_local
            owner1 = np.clip((gp[:, 0]*(G/L)).astype(np.int64), 0, G - 1)//ctx.nx_local
            rec['migrated'] = int(np.count_nonzero(owner0 != owner1))
        rec['ok'] = bool(rec['n_conserved'] and rec['owners_ok'] and rec.get('kick_relerr', 1) < PARITY_TOL
                         and rec.get('drift_exact', False) and rec.get('sum_mom2_relerr', 1) < 1e-12)
    rec = D.bcast(rec)
    del pbuf, mbuf, ibuf
    return rec


STAGES = ['grid_zero', 'deposit', 'halo_add', 'fft2d_forward', 'xsolve', 'fft2d_inverse', 'halo_fill', 'gather_kick_drift', 'migrate']


def kernel_names(cfg):
    T = 'double' if cfg['dtype'] == 'f64' else 'float'
    G, o = cfg['grid'], cfg['order']
    reach = 1 if cfg['diff'] <= 2 else cfg['diff']//2
    return {'fft2d_forward': f'fft2d_kernel<{T},{G},-1> (r2c along z + c2c along y per x plane, L2-resident hand-over)',
            'xsolve': f'xsolve2_kernel<{T},{G}> (c2c along x, Green\'s function, inverse c2c along x; peer loads/stores over NVLink)',
            'fft2d_inverse': f'fft2d_kernel<{T},{G},+1> (inverse c2c along y + c2r along z per x plane)',
            'gather_kick_drift': f'gather_kick_kernel<{o},{reach},{T},drift> (fused gradient + gather + kick + sum mom^2 + drift)',
            'deposit': f'deposit_kernel<{o},{T}> (mass scatter, red.global.add)', 'grid_zero': 'cudaMemsetAsync'}


def run_pm_config(D, cfg, steps, warmup, lib, ctx=None, pos=None, mom=None, want_e2e=False):
    """Times the PM cycle of one configuration; returns (record, ctx, state) — device-resident, CUDA events per stage."""
    torch = D.torch
    from concept_b200.pmsolver import make_kick_params
    from concept_b200.synthetic import zeldovich_particles
    L = BOXSIZE*cfg['grid']/GRID
    if ctx is None:
        ctx = make_context(D, cfg['grid'], L, cfg['dtype'])
    if pos is None:
        pos, mom = zeldovich_particles(cfg['n_side'], L, SIGMA, seed=0, device=D.dev)
    n_total = pos.shape[0]
    pbuf, mbuf, _, n_local = distribute(D, ctx, pos, mom)
    del pos, mom
    params = make_kick_params(**kick_kwargs(cfg, boxsize=L))
    sum2 = torch.zeros(1, dtype=torch.float64, device=D.dev)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(STAGES) + 1)] for _ in range(steps)]
    state = {'n': n_local}
    staged = ctx.hand_fft_available

    def cycle(i=None):
        n = state['n']
        p, m = pbuf[:n], mbuf[:n]
        sum2.zero_()
        if i is None:
            ctx.kick_drift(p, m, params, DT_OVER_MASS, sum_mom2=sum2)
        else:
            # same sequence as pm_kick_drift, split so that every kernel can be bracketed by events
            e = evs[i]
            e[0].record()
            ctx.grid_zero(); e[1].record()
            ctx.deposit(p, params.order, params.contribution); e[2].record()
            ctx.halo_add(); e[3].record()
            if staged:
                ctx.solve_fused_stage(params.prefactor, params.deconv_order, params.gauss, 1); e[4].record()
                ctx.solve_fused_stage(params.prefactor, params.deconv_order, params.gauss, 2); e[5].record()
                ctx.solve_fused_stage(params.prefactor, params.deconv_order, params.gauss, 3); e[6].record()
            else:
                if ctx.fused_solve_available:
                    ctx.solve_fused(params.prefactor, params.deconv_order, params.gauss)
                else:
                    ctx.fft_forward()
                    ctx.kspace_potential(params.prefactor, params.deconv_order, params.gauss, 1.0)
                    ctx.fft_backward()
                e[4].record(); e[5].record(); e[6].record()
            if D.world > 1:
                ctx.halo_fill_for(params.order, params.diff_order)
            e[7].record()
            ctx.gather_kick_drift(p, m, params.order, params.diff_order, params.kick_factor, DT_OVER_MASS, None, sum2)
            e[8].record()
        if D.world > 1:
            ctx.allreduce_sum(sum2)
            state['n'] = ctx.exchange(pbuf, mbuf, None, n)
        if i is not None:
            evs[i][9].record()

    warmup = max(warmup, 3)
    for _ in range(warmup):
        cycle()
    D.barrier()
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
    launches0 = lib.pm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    e0.record()
    for i in range(steps):
        cycle(i)
    e1.record()
    D.barrier()
    launches = lib.pm_launch_count() - launches0
    ctx.check_async_error()      # a tile dependency or a barrier that timed out would have invalidated the run: fail loudly
    clocks = sampler.stop() if D.rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    stage_ms = [sum(e[k].elapsed_time(e[k + 1]) for e in evs)/len(evs) for k in range(len(STAGES))]
    t = D.max_([ms_total] + stage_ms)
    ms_step = t[0]/steps
    stage_ms = t[1:]
    # measured state after exactly warmup + steps cycles: comparable across N
    n_after = int(round(D.sum_([float(state['n'])])[0]))
    sum_mom2 = float(sum2.item())

    peak, peak_src = measured_peaks()
    es = 8 if cfg['dtype'] == 'f64' else 4
    n_per, g3_per = n_total/D.world, cfg['grid']**3/D.world
    # algorithmic bytes per launch (SURVEY §8d / DESIGN §4), per rank
    alg = {'grid_zero': es*g3_per, 'deposit': 24*n_per + es*g3_per, 'fft2d_forward': 2*es*g3_per, 'xsolve': 2*es*g3_per,
           'fft2d_inverse': 2*es*g3_per, 'gather_kick_drift': 96*n_per + es*g3_per}
    kernels = {k: {'ms': ms, 'algorithmic_GB': alg[k]/1e9 if k in alg else None,
                   'achieved_GBps': (alg[k]/(ms*1e-3)/1e9 if (k in alg and ms > 0) else None),
                   'frac': (alg[k]/(ms*1e-3)/1e9/peak if (k in alg and ms > 0) else None)}
               for k, ms in zip(STAGES, stage_ms)}
    if D.world == 1 and staged and kernels['grid_zero']['ms'] < 0.02:
        # self-cleaning density grid: the forward 2-D kernel nullifies the rows it has read, so grid_zero finds nothing to do and
        # its 8·G³ bytes are written by fft2d_forward.  `frac` above keeps the kernel's own 16·G³; this is the same time with the
        # memset's algorithmic bytes attributed to the kernel that now writes them.
        fw = kernels['fft2d_forward']
        fw['achieved_GBps_incl_zeroing'] = (alg['fft2d_forward'] + alg['grid_zero'])/(fw['ms']*1e-3)/1e9
        fw['frac_incl_zeroing'] = fw['achieved_GBps_incl_zeroing']/peak
        kernels['grid_zero'].update(achieved_GBps=None, frac=None, note='the grid is already nullified (self-cleaning forward transform)')
    if D.world > 1 and staged and kernels['xsolve']['ms'] > 0:
        # several ranks: the x solve is bound by NVLink, not HBM.  Per rank and direction: its own remote loads plus the peers'
        # stores into it, 2·(P−1)/P of a rank's half-spectrum; reference 770 GB/s per direction (measured peer copy,
        # B200_PROFILING.md; 900 nominal).  The stage time includes the two device barriers and the local B -> A re-layout.
        nv = 2*(D.world - 1)/D.world*es*g3_per
        xs = kernels['xsolve']
        xs['nvlink'] = {'bytes_per_direction': nv, 'achieved_GBps': nv/(xs['ms']*1e-3)/1e9, 'peak_GBps': 770.0,
                        'frac': nv/(xs['ms']*1e-3)/1e9/770.0}
    names = kernel_names(cfg)
    dom = max((k for k in alg if k in names and k != 'grid_zero'), key=lambda k: kernels[k]['ms'])
    b_alg = 120*n_total + 6*es*cfg['grid']**3
    rec = {'ms_per_step': ms_step, 'value': n_total/(ms_step*1e-3), 'n_total': n_total, 'launches': int(launches), 'clocks': clocks,
           'particles_after': n_after, 'sum_mom2': sum_mom2, 'kernels': kernels, 'dominant': dom, 'dominant_name': names[dom],
           'alg_dominant': alg[dom], 'peak': peak, 'peak_source': peak_src,
           'cycle': {'algorithmic_bytes': b_alg, 'achieved_GBps_per_gpu': b_alg/D.world/(ms_step*1e-3)/1e9,
                     'frac': b_alg/D.world/(ms_step*1e-3)/1e9/peak},
           'hand_written_fft': bool(staged)}
    return rec, ctx, (pbuf, mbuf, state, params, cycle)


def run_particle_order(D, ctx, pbuf, mbuf, n, params):
    """One GPU: what the memory order of the particles is worth — the same cycle on a randomly permuted array, the time of
    pm_sort_particles (the tile_sort analogue), and the cycle after sorting."""
    torch = D.torch
    perm = torch.randperm(n, device=D.dev)
    p, m = pbuf[:n][perm].contiguous(), mbuf[:n][perm].contiguous()
    sum2 = torch.zeros(1, dtype=torch.float64, device=D.dev)

    def timed(fn, reps):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)/reps
    cyc = lambda: ctx.kick_drift(p, m, params, DT_OVER_MASS, sum_mom2=sum2)
    shuffled = timed(cyc, 3)
    ctx.sort_particles(p, m)          # warm-up (scratch allocation); the array is sorted now
    q, r = p[perm % n].contiguous(), m[perm % n].contiguous()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.sort_particles(q, r); e1.record(); torch.cuda.synchronize()
    sort_ms = e0.elapsed_time(e1)
    p, m = q, r
    after = timed(lambda: ctx.kick_drift(p, m, params, DT_OVER_MASS, sum_mom2=sum2), 3)
    return {'cycle_ms_random_order': shuffled, 'sort_ms': sort_ms, 'cycle_ms_after_sort': after,
            'note': 'main.timeloop re-orders by grid cell at synchronised steps every cell_sort_period (default 64) base steps'}


def run_clustered(D, ctx, cfg, params):
    """One GPU: SURVEY §8d's second synthetic set — the same lattice displaced by sigma = 2 inter-particle spacings (four
    grid cells: shell crossing everywhere, cells with many particles, reductions that collide) — in lattice order and after
    pm_sort_particles."""
    from concept_b200.synthetic import zeldovich_particles
    torch = D.torch
    L = float(cfg['grid'])
    p, m = zeldovich_particles(cfg['n_side'], L, 2.0, seed=0, device=D.dev)
    sum2 = torch.zeros(1, dtype=torch.float64, device=D.dev)

    def timed(reps):
        ctx.kick_drift(p, m, params, 0.0, sum_mom2=sum2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ctx.kick_drift(p, m, params, 0.0, sum_mom2=sum2)      # no drift: the same particle set every time
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)/reps
    lattice_order = timed(5)
    ctx.sort_particles(p, m)
    after = timed(5)
    ctx.check_async_error()
    return {'workload': 'the lattice of configs[1] displaced by sigma = 2.0 spacings (4 grid cells), kick without drift',
            'cycle_ms_lattice_order': lattice_order, 'cycle_ms_after_sort': after}


def run_e2e(D, ctx, pbuf, mbuf, state, params, cycle, steps):
    """The same cycle with HOST particle buffers (pinned), host<->device copies inside the timed region."""
    torch = D.torch
    n = state['n']
    n_total = int(round(D.sum_([float(n)])[0]))
    e2e_steps = max(3, min(steps, 10))
    if D.world == 1:
        hp = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
        hm = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
        hp.copy_(pbuf[:n]); hm.copy_(mbuf[:n])
        hp_np, hm_np = hp.numpy(), hm.numpy()
        for _ in range(2):
            ctx.kick_long_host(hp_np, hm_np, params, dt_over_mass=DT_OVER_MASS, want_sum=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.kick_long_host(hp_np, hm_np, params, dt_over_mass=DT_OVER_MASS, want_sum=True)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0)/e2e_steps
        api = 'pm_kick_long_host (pinned host pos/mom in, kick + fused drift, pos/mom out)'
    else:
        hp = torch.empty((pbuf.shape[0], 3), dtype=torch.float64, pin_memory=True)
        hm = torch.empty((pbuf.shape[0], 3), dtype=torch.float64, pin_memory=True)
        hp[:n].copy_(pbuf[:n]); hm[:n].copy_(mbuf[:n])

        def e2e_step():
            nn = state['n']
            pbuf[:nn].copy_(hp[:nn], non_blocking=True)
            mbuf[:nn].copy_(hm[:nn], non_blocking=True)
            cycle()
            nn = state['n']
            hp[:nn].copy_(pbuf[:nn], non_blocking=True)
            hm[:nn].copy_(mbuf[:nn], non_blocking=True)
            torch.cuda.synchronize()
        for _ in range(2):
            e2e_step()
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        D.barrier()
        sec = D.max_([(time.perf_counter() - t0)/e2e_steps])[0]
        api = 'pinned host slabs -> device, distributed cycle, device -> host (per rank)'
    gbs = 2*24*n_total/sec/1e9
    return {'value': n_total/sec, 'unit': UNIT, 'h2d_bytes_per_step': 2*24*n_total, 'd2h_bytes_per_step': 2*24*n_total + 8,
            'ms_per_step': sec*1e3, 'h2d_GBps': gbs, 'd2h_GBps': gbs, 'api': api}


def extra_record(rec, cfg, world):
    out = {'workload': cfg['label'], 'ms_per_cycle': rec['ms_per_step'], 'particle_updates_per_s': rec['value'],
           'cycle_frac_of_hbm_peak': rec['cycle']['frac'], 'algorithmic_bytes': rec['cycle']['algorithmic_bytes'],
           'hand_written_fft': rec['hand_written_fft'], 'particles_after': rec['particles_after'],
           'kernels': {k: {'ms': v['ms'], 'frac': v['frac']} for k, v in rec['kernels'].items()}}
    return out


def run_p3m_config(D, cfg, steps, lib):
    """configs[4]: the P³M cycle — long-range kick (Gaussian-split potential, order-4 differences) + short-range pair
    kick of all particles (one rung) + drift + migration.  Device-resident; CUDA events."""
    torch = D.torch
    import numpy as np
    from concept_b200.pmsolver import make_kick_params
    from concept_b200.synthetic import zeldovich_particles
    from concept_b200.shortrange import PairKick
    L = BOXSIZE*cfg['grid']/GRID
    ctx = make_context(D, cfg['grid'], L, cfg['dtype'])
    pos, mom = zeldovich_particles(cfg['n_side'], L, SIGMA, seed=0, device=D.dev)
    n_total = pos.shape[0]
    pbuf, mbuf, _, n = distribute(D, ctx, pos, mom)
    del pos, mom
    params = make_kick_params(**kick_kwargs(cfg, boxsize=L))
    job = PairKick(ctx, L, cfg['grid'], n_total, pbuf.shape[0], G_NEWTON)
    sum2 = torch.zeros(1, dtype=torch.float64, device=D.dev)
    state = {'n': n}
    names = ['long_range_kick', 'short_range_kick', 'drift_migrate']
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]

    def cycle(i=None):
        n = state['n']
        p, m = pbuf[:n], mbuf[:n]
        e = evs[i] if i is not None else None
        if e: e[0].record()
        sum2.zero_()
        ctx.kick_long(p, m, params, sum_mom2=sum2)
        if e: e[1].record()
        job.kick(pbuf, mbuf, n)
        if e: e[2].record()
        ctx.drift(p, m, DT_OVER_MASS)
        if D.world > 1:
            state['n'] = job.exchange(pbuf, mbuf, n)
        if e: e[3].record()
    for _ in range(2):
        cycle()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        cycle(i)
    e1.record()
    D.barrier()
    ctx.check_async_error()
    t = D.max_([e0.elapsed_time(e1)] + [sum(e[k].elapsed_time(e[k + 1]) for e in evs)/steps for k in range(3)])
    pairs, cands = job.pair_stats(pbuf, state['n'])
    pairs_total, cands_total = D.sum_([float(pairs), float(cands)])
    ms = t[0]/steps
    rec = {'workload': cfg['label'], 'ms_per_cycle': ms, 'particle_updates_per_s': n_total/(ms*1e-3),
           'stages_ms': dict(zip(names, t[1:])), 'pairs_within_range': pairs_total, 'candidates_examined': cands_total,
           'pair_interactions_per_s': 2*pairs_total/(t[2]*1e-3) if t[2] > 0 else None,
           'note': 'every pair is evaluated from both sides (gather form, no atomics): pair_interactions_per_s counts both',
           'n_gpus': D.world,
           'particles_after': int(round(D.sum_([float(state['n'])])[0]))}
    ctx.close()
    return rec


def run_example_basic():
    """configs[0]: the reference's param/example_basic, unmodified (vendored byte-identical as tests/golden/example_basic),
    through `python -m concept_b200 -p`: 64³ particles realised on the fly, P³M on a 128³ grid with 8 rungs, a = 0.02 → 1,
    power spectrum at a = 1.  Wall time of the whole process (imports, initial conditions, time loop, output)."""
    import tempfile
    import numpy as np
    with tempfile.TemporaryDirectory(prefix='example_basic_') as tmp:
        env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
        t0 = time.perf_counter()
        r = subprocess.run([sys.executable, '-m', 'concept_b200', '-p', os.path.join(ROOT, 'tests', 'golden', 'example_basic')],
                           cwd=tmp, env=env, capture_output=True, text=True, timeout=1200)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {'error': (r.stdout[-500:] + r.stderr[-500:])}
        k, modes, power, linear = np.loadtxt(os.path.join(tmp, 'output', 'example_basic', 'powerspec_a=1.00')).T
    large = (modes >= 100) & (k < 0.1)
    return {'workload': 'configs[0]: param/example_basic unmodified (64^3 particles, P3M grid 128, 8 rungs, a = 0.02 -> 1)',
            'wall_s': wall, 'finished': 'a = 1' in r.stdout.splitlines()[-1],
            'P_over_linear_large_scales': [float(v) for v in power[large]/linear[large]],
            'P_over_linear_at_k_max': float(power[-1]/linear[-1])}


def run_gpu(args):
    D = Dist(args)
    torch = D.torch
    from concept_b200 import _lib
    lib = _lib.load()
    cfg = CONFIGS['config2']

    # ---- parity at this N, before anything is timed ----
    parity = {'tolerance': PARITY_TOL, 'checker': 'oracle/pm_oracle.c (C/OpenMP restatement of the reference, pinned to its golden vectors)'}
    ctx128 = make_context(D, 128, 128.0, 'f64')
    parity['G128'] = parity_check(D, ctx128, dict(cfg, grid=128), 100000, seed=11)
    ctx128.close()
    ctx = make_context(D, cfg['grid'], BOXSIZE, cfg['dtype'])
    parity['G512'] = parity_check(D, ctx, cfg, 100000, seed=12)
    parity['ok'] = bool(parity['G128']['ok'] and parity['G512']['ok'])
    if not parity['ok']:
        if D.rank == 0:
            print(json.dumps({'metric': METRIC, 'value': None, 'unit': UNIT, 'n_gpus': D.world, 'parity': parity,
                              'error': 'parity check failed before the timed region'}))
        ctx.close()
        if D.world > 1:
            D.dist.destroy_process_group()
        raise SystemExit(3)

    # ---- configs[1]: the headline ----
    rec, ctx, (pbuf, mbuf, state, params, cycle) = run_pm_config(D, cfg, args.steps, args.warmup, lib, ctx=ctx)
    e2e = run_e2e(D, ctx, pbuf, mbuf, state, params, cycle, args.steps)
    particle_order = None
    if D.world == 1 and not args.no_extra:
        try:
            particle_order = run_particle_order(D, ctx, pbuf, mbuf, state['n'], params)
        except Exception as exc:
            particle_order = {'error': repr(exc)}
        try:
            particle_order['clustered'] = run_clustered(D, ctx, cfg, params)
        except Exception as exc:
            particle_order['clustered'] = {'error': repr(exc)}
    del pbuf, mbuf, cycle
    ctx.close()
    torch.cuda.empty_cache()

    # ---- the other configurations of BASELINE.json ----
    extra = {}
    if not args.no_extra:
        try:
            r3, c3, st3 = run_pm_config(D, CONFIGS['config3'], min(args.steps, 10), 3, lib)
            extra['config3'] = extra_record(r3, CONFIGS['config3'], D.world)
            extra['config3']['parity_note'] = ('fp32 grid has no reference counterpart (the reference is fp64-only); '
                                               'tests/test_gpu_parity.py states rms force error <= 1e-5 against the fp64 oracle')
            del st3
            c3.close()
            torch.cuda.empty_cache()
        except Exception as exc:
            extra['config3'] = {'error': repr(exc)}
        if D.world == 8:
            try:
                c4 = make_context(D, CONFIGS['config4']['grid'], BOXSIZE*2, 'f64')
                # parity of the 1024³ distributed path first (same checker, 10⁵ particles; the oracle needs ~30 GB of host memory)
                parity['G1024'] = parity_check(D, c4, CONFIGS['config4'], 100000, seed=13)
                parity['ok'] = bool(parity['ok'] and parity['G1024']['ok'])
                r4, c4, st4 = run_pm_config(D, CONFIGS['config4'], min(args.steps, 10), 3, lib, ctx=c4)
                extra['config4'] = extra_record(r4, CONFIGS['config4'], D.world)
                extra['config4']['parity'] = parity['G1024']
                del st4
                c4.close()
                torch.cuda.empty_cache()
            except Exception as exc:
                extra['config4'] = {'error': repr(exc)}
        else:
            extra['config4'] = {'skipped': 'configs[3] is defined on 8 GPUs (run with --gpus 8)'}
        try:
            extra['config5'] = run_p3m_config(D, CONFIGS['config5'], min(args.steps, 5), lib)
        except Exception as exc:
            extra['config5'] = {'error': repr(exc)}
        if D.world == 1:
            try:
                extra['config1'] = run_example_basic()
            except Exception as exc:
                extra['config1'] = {'error': repr(exc)}
        else:
            extra['config1'] = {'skipped': 'configs[0] (param/example_basic, 64^3 particles) is run by the one-GPU bench'}

    if D.rank == 0:
        dom = rec['dominant']
        roofline = {'bound': 'hbm', 'kernel': rec['dominant_name'], 'achieved': rec['kernels'][dom]['achieved_GBps'], 'peak': rec['peak'],
                    'unit': 'GB/s', 'frac': rec['kernels'][dom]['frac'], 'traffic': committed_traffic(dom, D.world),
                    'traffic_source': 'profiles/ (ncu --set full capture of this kernel, per launch; not re-measured in this run)',
                    'peak_source': rec['peak_source'], 'kernel_ms': rec['kernels'][dom]['ms'],
                    'algorithmic_bytes_per_launch': rec['alg_dominant'], 'kernels': rec['kernels'], 'cycle': rec['cycle']}
        if dom == 'xsolve' and rec['kernels'][dom].get('nvlink'):
            roofline['note'] = ('on several ranks the dominant kernel moves its lines over NVLink: see "nvlink" '
                                '(bytes per direction and rank against the measured 770 GB/s peer-copy rate); "frac" is its HBM fraction')
            roofline['nvlink'] = rec['kernels'][dom]['nvlink']
        cpu = None
        if D.world == 1:
            try:
                rate, cores, sec = cpu_cycle_rate(cfg, N_SIDE//2, GRID//2, cycles=2, warm=1)
                sample = f'{N_SIDE//2}^3 particles / {GRID//2}^3 grid (1/8 of the workload, same particles per cell), 2 cycles'
                if sec < 1.5:   # plenty of cores: time the full workload as well
                    rate, cores, sec = cpu_cycle_rate(cfg, N_SIDE, GRID, cycles=2, warm=1)
                    sample = f'full workload {N_SIDE}^3 / {GRID}^3, 2 cycles after 1 warm-up'
                cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample, 'sec_per_cycle': sec}
            except Exception as exc:   # the baseline is a report, never a gate
                cpu = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'failed: {exc}'}
        out = {
            'metric': METRIC, 'value': rec['value'], 'unit': UNIT, 'n_gpus': D.world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': rec['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(cfg, D.world), 'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': rec['launches'], 'clocks': rec['clocks'], 'parity': parity,
            'sum_mom2': rec['sum_mom2'], 'particles_after': rec['particles_after'],
            'cycles_before_state': max(args.warmup, 3) + args.steps, 'particle_order': particle_order, 'extra_configs': extra,
        }
        print(json.dumps(out))
    if D.world > 1:
        D.dist.destroy_process_group()


def committed_traffic(kernel, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed ncu --set full capture"""
    if world != 1:
        return None
    for name in ('r02_traffic.json', 'r01_traffic.json'):
        try:
            with open(os.path.join(ROOT, 'profiles', name)) as f:
                return json.load(f)['dram_bytes_per_launch'].get(kernel)
        except Exception:
            continue
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=int(os.environ.get('WORLD_SIZE', '1')))
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-extra', action='store_true', help='only configs[1] (skip the extra_configs legs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
