#!/usr/bin/env python
"""bench.py — PM-gravity cycle benchmark (BASELINE.json: particle-updates/s and ms per PM cycle,
256³ particles / 512³ grid, CIC, fp64, on 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one PM cycle: long-range kick (zero grid → CIC deposit → r2c FFT →
Green's function/deconvolution → c2r FFT → fused gradient + gather + kick + Σmom²) followed by a
full drift and, on several GPUs, the slab migration of particles.  Strong scaling: the same
256³/512³ problem is split into x-slabs over the N ranks.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` is the same cycle
through the C-ABI entry point that takes HOST buffers (pm_kick_long_host: H2D of pos+mom, cycle,
D2H of pos+mom every step).  `roofline` is for the dominant hand-written kernel (fused
gradient/gather/kick), timed live with CUDA events inside the timed steps.  `cpu_baseline` is the
C/OpenMP restatement of the reference loops (oracle/pm_oracle.c) on this box's host cores.

--impl reference: the reference's CPU implementation of the same cycle.  The compiled reference
cannot be built in this image (no MPI/FFTW/GSL), so this is the oracle port with
all host threads, on a bounded sample of the workload per step.
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
ORDER, DIFF_ORDER = 2, 2
G_NEWTON = 4.4985024439973154e-05
SIGMA = 0.3
KICK = dict(mass=1.0, boxsize=BOXSIZE, gridsize=GRID, order=ORDER, G_Newton=G_NEWTON,
            dt_rho_over_dt1=2.0, dt_kick=1e-3, diff_order=DIFF_ORDER)
DT_OVER_MASS = 1e-4
METRIC = 'particle-updates/sec (one PM cycle: long-range kick + drift), 256^3 particles / 512^3 grid'
UNIT = 'particle-updates/s'


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
def cpu_cycle_rate(n_side, grid, cycles=1, warm=1):
    """Times the C/OpenMP restatement on (n_side³, grid³); returns (updates/s, threads, seconds/cycle)."""
    from oracle import c_oracle as C
    from concept_b200.synthetic import zeldovich_particles
    L = BOXSIZE*grid/GRID
    pos, mom = zeldovich_particles(n_side, L, SIGMA, seed=0)
    pos, mom = pos.numpy(), mom.numpy()
    work = C.Workspace(grid)
    kw = dict(KICK, boxsize=L, gridsize=grid, work=work)
    ts = []
    for it in range(warm + cycles):
        t0 = time.perf_counter()
        C.kick_long(pos, mom, **kw)
        C.sum_mom2(mom)
        C.drift(pos, mom, DT_OVER_MASS, L)
        if it >= warm:
            ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return n_side**3/sec, C.num_threads(), sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_side, grid = N_SIDE//2, GRID//2      # bounded sample: 1/8 of the workload, same particles per cell
    from oracle import c_oracle as C
    from concept_b200.synthetic import zeldovich_particles
    L = BOXSIZE*grid/GRID
    pos, mom = zeldovich_particles(n_side, L, SIGMA, seed=0)
    pos, mom = pos.numpy(), mom.numpy()
    work = C.Workspace(grid)
    kw = dict(KICK, boxsize=L, gridsize=grid, work=work)

    def step():
        C.kick_long(pos, mom, **kw)
        C.sum_mom2(mom)
        C.drift(pos, mom, DT_OVER_MASS, L)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sec = (time.perf_counter() - t0)/args.steps
    value = n_side**3/sec
    sample = f'{n_side}^3 particles / {grid}^3 grid per step (1/8 of the workload, same particles per cell)'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec*1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': C.num_threads(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'compiled reference unbuildable here (needs mpicc/FFTW-MPI/GSL); C/OpenMP port of its loops, pinned to '
                'golden vectors from the reference run in pure-Python mode',
    }))


def workload_config(n_gpus):
    return {'workload': 'configs[1]: 256^3 particles, 512^3 PM grid, CIC, fp64, deconvolution order 4, '
                        'finite-difference order 2, Zel\'dovich-displaced lattice (sigma = 0.3 spacings, seed 0)',
            'particles': N_SIDE**3, 'grid': GRID, 'interpolation': 'CIC', 'grid_dtype': 'f64',
            'decomposition': f'{n_gpus} x-slab(s)', 'l2_policy': 'inputs larger than L2 (0.4 GB particles, 1.08 GB grid)'}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from concept_b200.pmsolver import PMContext, make_kick_params
    from concept_b200.synthetic import zeldovich_particles
    from concept_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run for --gpus > 1')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    # synthetic inputs: every rank builds the same particle load and keeps its slab
    pos, mom = zeldovich_particles(N_SIDE, BOXSIZE, SIGMA, seed=0, device=dev)
    n_total = pos.shape[0]
    ctx = PMContext(GRID, BOXSIZE, dtype='f64', rank=rank, nranks=world, device=local_rank)
    if world > 1:
        def _bcast(obj):
            box = [obj]
            dist.broadcast_object_list(box, src=0)
            return box[0]

        def _allgather(obj):
            out = [None]*world
            dist.all_gather_object(out, obj)
            return out
        ctx.connect(_bcast, _allgather, rank == 0)
        owner = torch.clamp((pos[:, 0]*(GRID/BOXSIZE)).to(torch.int64), 0, GRID - 1)//ctx.nx_local
        keep = owner == rank
        n_local = int(keep.sum().item())
        capacity = int(n_total/world*1.5) + 4096
        pbuf = torch.zeros((capacity, 3), dtype=torch.float64, device=dev)
        mbuf = torch.zeros((capacity, 3), dtype=torch.float64, device=dev)
        pbuf[:n_local] = pos[keep]
        mbuf[:n_local] = mom[keep]
        del pos, mom, owner, keep
    else:
        n_local = n_total
        pbuf, mbuf = pos, mom
    params = make_kick_params(**KICK)
    sum2 = torch.zeros(1, dtype=torch.float64, device=dev)
    # stages of one cycle, each bracketed by CUDA events inside the timed region
    STAGES = ['grid_zero', 'deposit', 'halo_add', 'fft2d_forward', 'xsolve', 'fft2d_inverse', 'halo_fill', 'gather_kick_drift', 'migrate']
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(STAGES) + 1)] for _ in range(args.steps)]
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
            if world > 1:
                ctx.halo_fill()
            e[7].record()
            ctx.gather_kick_drift(p, m, params.order, params.diff_order, params.kick_factor, DT_OVER_MASS, None, sum2)
            e[8].record()
        if world > 1:
            ctx.allreduce_sum(sum2)
            state['n'] = ctx.exchange(pbuf, mbuf, None, n)
        if i is not None:
            evs[i][9].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        cycle()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.pm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        cycle(i)
    e1.record()
    barrier()
    launches = lib.pm_launch_count() - launches0
    ctx.check_async_error()      # a tile dependency that timed out would have invalidated the run: fail loudly
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    stage_ms = [sum(e[k].elapsed_time(e[k + 1]) for e in evs)/len(evs) for k in range(len(STAGES))]
    t = torch.tensor([ms_total] + stage_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t[0].item()/args.steps
    stage_ms = [v.item() for v in t[1:]]
    value = n_total/(ms_step*1e-3)

    # ---- end-to-end through the host-buffer C-ABI entry point (rank-local slab) ----
    n = state['n']
    hp = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
    hm = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
    hp.copy_(pbuf[:n]); hm.copy_(mbuf[:n])
    hp_np, hm_np = hp.numpy(), hm.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    e2e = None
    if world == 1:
        for _ in range(2):
            ctx.kick_long_host(hp_np, hm_np, params, dt_over_mass=DT_OVER_MASS, want_sum=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.kick_long_host(hp_np, hm_np, params, dt_over_mass=DT_OVER_MASS, want_sum=True)
        torch.cuda.synchronize()
        e2e_sec = (time.perf_counter() - t0)/e2e_steps
        e2e = {'value': n_total/e2e_sec, 'unit': UNIT, 'h2d_bytes_per_step': 2*24*n, 'd2h_bytes_per_step': 2*24*n + 8,
               'ms_per_step': e2e_sec*1e3, 'api': 'pm_kick_long_host (pinned host pos/mom in, kick + fused drift, pos/mom out)'}
    else:
        # every rank moves its own slab's particles host<->device around the same distributed cycle
        def e2e_step():
            nn = state['n']
            pbuf[:nn].copy_(hp[:nn], non_blocking=True)
            mbuf[:nn].copy_(hm[:nn], non_blocking=True)
            cycle()
            nn = state['n']
            hp[:nn].copy_(pbuf[:nn], non_blocking=True) if nn <= hp.shape[0] else None
            hm[:nn].copy_(mbuf[:nn], non_blocking=True) if nn <= hm.shape[0] else None
            torch.cuda.synchronize()
        hp = torch.empty((pbuf.shape[0], 3), dtype=torch.float64, pin_memory=True)
        hm = torch.empty((pbuf.shape[0], 3), dtype=torch.float64, pin_memory=True)
        hp[:n].copy_(pbuf[:n]); hm[:n].copy_(mbuf[:n])
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        sec = torch.tensor([(time.perf_counter() - t0)/e2e_steps], dtype=torch.float64, device=dev)
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        e2e = {'value': n_total/sec.item(), 'unit': UNIT, 'h2d_bytes_per_step': 2*24*n_total, 'd2h_bytes_per_step': 2*24*n_total + 8,
               'ms_per_step': sec.item()*1e3, 'api': 'pinned host slabs -> device, distributed cycle, device -> host (per rank)'}

    if rank == 0:
        peak, peak_src = measured_peaks()
        n_per = n_total/world
        g3_per = GRID**3/world
        es = 8
        # algorithmic bytes per launch (SURVEY §8d / DESIGN §4), per rank
        alg = {'grid_zero': es*g3_per, 'deposit': 24*n_per + es*g3_per, 'fft2d_forward': 2*es*g3_per, 'xsolve': 2*es*g3_per,
               'fft2d_inverse': 2*es*g3_per, 'gather_kick_drift': 96*n_per + es*g3_per}
        kernels = {k: {'ms': ms, 'algorithmic_GB': alg[k]/1e9 if k in alg else None,
                       'achieved_GBps': (alg[k]/(ms*1e-3)/1e9 if (k in alg and ms > 0) else None)}
                   for k, ms in zip(STAGES, stage_ms)}
        names = {'fft2d_forward': 'fft2d_kernel<double,512,-1> (r2c along z + c2c along y per x plane, L2-resident hand-over)',
                 'xsolve': 'xsolve2_kernel<double,512> (c2c along x, Green\'s function, inverse c2c along x)',
                 'fft2d_inverse': 'fft2d_kernel<double,512,+1> (inverse c2c along y + c2r along z per x plane)',
                 'gather_kick_drift': 'gather_kick_kernel<2,1,double,drift> (fused gradient + CIC gather + kick + sum mom^2 + drift)',
                 'deposit': 'deposit_kernel<2,double> (CIC scatter, red.global.add.f64)', 'grid_zero': 'cudaMemsetAsync'}
        dom = max((k for k in alg if k in names), key=lambda k: kernels[k]['ms'])
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed ncu --set full capture
        try:
            with open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')) as f:
                traffic = json.load(f)['dram_bytes_per_launch'].get(dom) if world == 1 else None
        except Exception:
            traffic = None
        b_alg = 120*n_total + 48*GRID**3
        roofline = {'bound': 'hbm', 'kernel': names[dom], 'achieved': kernels[dom]['achieved_GBps'], 'peak': peak, 'unit': 'GB/s',
                    'frac': kernels[dom]['achieved_GBps']/peak, 'traffic': traffic,
                    'peak_source': peak_src, 'kernel_ms': kernels[dom]['ms'], 'algorithmic_bytes_per_launch': alg[dom],
                    'kernels': kernels,
                    'cycle': {'algorithmic_bytes': b_alg, 'achieved_GBps_per_gpu': b_alg/world/(ms_step*1e-3)/1e9,
                              'frac': b_alg/world/(ms_step*1e-3)/1e9/peak}}
        cpu = None
        if world == 1:
            try:
                rate, cores, sec = cpu_cycle_rate(N_SIDE//2, GRID//2, cycles=2, warm=1)
                sample = f'{N_SIDE//2}^3 particles / {GRID//2}^3 grid (1/8 of the workload, same particles per cell), 2 cycles'
                if sec < 1.5:   # plenty of cores: time the full workload once as well
                    rate, cores, sec = cpu_cycle_rate(N_SIDE, GRID, cycles=1, warm=1)
                    sample = f'full workload {N_SIDE}^3 / {GRID}^3, 1 cycle after 1 warm-up'
                cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample, 'sec_per_cycle': sec}
            except Exception as exc:   # the baseline is a report, never a gate
                cpu = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'failed: {exc}'}
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(world), 'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': int(launches), 'clocks': clocks,
            'sum_mom2': float(sum2.item()), 'particles_after': int(n_total),
        }
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=int(os.environ.get('WORLD_SIZE', '1')))
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
