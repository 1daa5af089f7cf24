"""Per-stage timings of the distributed PM cycle (development aid).  Launch with torchrun."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200.pmsolver import PMContext, make_kick_params  # noqa: E402
from concept_b200.synthetic import zeldovich_particles  # noqa: E402


def main():
    world, rank, lr = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    n_side = int(os.environ.get('PM_N', 256)); G = int(os.environ.get('PM_G', 512)); L = 512.0
    torch.cuda.set_device(lr)
    dev = torch.device('cuda', lr)
    dist.init_process_group('nccl', device_id=dev)
    pos, mom = zeldovich_particles(n_side, L, 0.3, seed=0, device=dev)
    N = pos.shape[0]
    ctx = PMContext(G, L, rank=rank, nranks=world, device=lr)
    def _bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def _allgather(obj):
        out = [None]*world
        dist.all_gather_object(out, obj)
        return out
    ctx.connect(_bcast, _allgather, rank == 0)
    fused = ctx.fused_solve_available and not int(os.environ.get('PM_NOFUSE', '0'))
    keep = torch.clamp((pos[:, 0]*(G/L)).to(torch.int64), 0, G - 1)//ctx.nx_local == rank
    n = int(keep.sum())
    cap = int(N/world*1.5) + 4096
    pb = torch.zeros((cap, 3), dtype=torch.float64, device=dev); mb = torch.zeros_like(pb)
    pb[:n] = pos[keep]; mb[:n] = mom[keep]
    del pos, mom, keep
    p = make_kick_params(mass=1.0, boxsize=L, gridsize=G, order=2, G_Newton=4.4985e-5, dt_rho_over_dt1=2.0, dt_kick=1e-3)
    s = torch.zeros(1, dtype=torch.float64, device=dev)
    st = {'n': n}
    stages = [
        ('grid_zero', lambda: ctx.grid_zero()),
        ('deposit', lambda: ctx.deposit(pb[:st['n']], 2, p.contribution)),
        ('halo_add', lambda: ctx.halo_add()),
        *([('solve_fused', lambda: ctx.solve_fused(p.prefactor, p.deconv_order, 0.0))] if fused else [
            ('fft_forward', lambda: ctx.fft_forward()),
            ('kspace', lambda: ctx.kspace_potential(p.prefactor, p.deconv_order, 0.0, 1.0)),
            ('fft_backward', lambda: ctx.fft_backward())]),
        ('halo_fill', lambda: ctx.halo_fill()),
        ('gather_kick', lambda: ctx.gather_kick(pb[:st['n']], mb[:st['n']], 2, 2, p.kick_factor, None, s)),
        ('drift', lambda: ctx.drift(pb[:st['n']], mb[:st['n']], 1e-4)),
        ('allreduce', lambda: ctx.allreduce_sum(s)),
        ('exchange', lambda: st.__setitem__('n', ctx.exchange(pb, mb, None, st['n']))),
    ]
    acc = {k: [] for k, _ in stages}
    tot = []
    for rep in range(8):
        dist.barrier(); torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
        evs[0].record()
        for i, (_, fn) in enumerate(stages):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        if rep >= 3:
            for i, (k, _) in enumerate(stages):
                acc[k].append(evs[i].elapsed_time(evs[i + 1]))
            tot.append(evs[0].elapsed_time(evs[-1]))
    out = {k: round(sorted(v)[len(v)//2], 4) for k, v in acc.items()}
    out['cycle_ms'] = round(sorted(tot)[len(tot)//2], 4)
    outs = [None]*world
    dist.all_gather_object(outs, out)
    if rank == 0:
        print(json.dumps({'world': world, 'N': N, 'G': G, 'per_rank': outs}))
    ctx.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
