"""Timing of the three solve kernels on several ranks (development aid).  Launch with torchrun; PM_G = grid size."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concept_b200.pmsolver import PMContext, make_kick_params  # noqa: E402


def main():
    world, rank, lr = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    G = int(os.environ.get('PM_G', 512)); L = float(G)
    torch.cuda.set_device(lr)
    dev = torch.device('cuda', lr)
    dist.init_process_group('nccl', device_id=dev)
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
    n = 200000
    pos = torch.rand((n, 3), dtype=torch.float64, device=dev)*L
    pos[:, 0] = (pos[:, 0]/world + rank*L/world).clamp(max=(rank + 1)*L/world*(1 - 1e-12))
    p = make_kick_params(mass=1.0, boxsize=L, gridsize=G, order=2, G_Newton=4.4985e-5, dt_rho_over_dt1=2.0, dt_kick=1e-3)
    acc = [[], [], []]
    for rep in range(7):
        ctx.grid_zero(); ctx.deposit(pos, 2, p.contribution); ctx.halo_add()
        dist.barrier(); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for st in (1, 2, 3):
            ctx.solve_fused_stage(p.prefactor, p.deconv_order, p.gauss, st)
            ev[st].record()
        torch.cuda.synchronize()
        if rep >= 2:
            for k in range(3):
                acc[k].append(ev[k].elapsed_time(ev[k + 1]))
    ctx.check_async_error()
    out = [round(sorted(v)[len(v)//2], 4) for v in acc]
    outs = [None]*world
    dist.all_gather_object(outs, out)
    if rank == 0:
        print(json.dumps({'world': world, 'G': G, 'x_local': os.environ.get('PM_X_LOCAL', '0'), 'fwd_x_inv_ms_per_rank': outs}))
    ctx.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
