"""Process layout for the x-slab decomposition: one process per GPU.

Reference counterpart: communication.py — `domain_subdivisions` / `get_domain_info` (:692-741,
:1765-1807), `exchange` (:135-517).  The reference cuts the box into a 3-D cuboid of MPI domains; here
rank r owns the x-slab [r·L/P, (r+1)·L/P) so that particle domains coincide with the FFT slabs.
`torch.distributed` (NCCL, or gloo on CPU for host-logic tests) is the plumbing for small host-side
objects; bulk particle/grid traffic goes through the library's own NCCL communicator (pm_comm_init).
"""
import os

import torch

rank = 0
nprocs = 1
local_rank = 0
master = True
_initialized = False


def init(backend=None):
    """Read RANK / WORLD_SIZE / LOCAL_RANK (torchrun) and join the process group if world > 1."""
    global rank, nprocs, local_rank, master, _initialized
    nprocs = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', str(rank)))
    master = rank == 0
    if nprocs > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = 'nccl' if torch.cuda.is_available() else 'gloo'
            kw = {}
            if backend == 'nccl':
                torch.cuda.set_device(local_rank)
                kw['device_id'] = torch.device('cuda', local_rank)
            dist.init_process_group(backend, **kw)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    _initialized = True
    return rank, nprocs


def finalize():
    """Leave the process group (end of a run under torchrun)."""
    global _initialized
    if nprocs > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    _initialized = False


def bcast(obj, root=0):
    if nprocs == 1:
        return obj
    import torch.distributed as dist
    box = [obj if rank == root else None]
    dist.broadcast_object_list(box, src=root)
    return box[0]


def allreduce_sum(value):
    """Sum of a python float over ranks (host-side scalars such as Σmom², analysis.py:3971)."""
    if nprocs == 1:
        return value
    import torch.distributed as dist
    dev = torch.device('cuda', local_rank) if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    return float(t.item())


def allreduce_ints(values):
    """Element-wise sum of a list of python ints over the ranks (rung populations, flags)."""
    if nprocs == 1:
        return [int(v) for v in values]
    import torch.distributed as dist
    dev = torch.device('cuda', local_rank) if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    return [int(v) for v in t.tolist()]


def allgather(obj):
    if nprocs == 1:
        return [obj]
    import torch.distributed as dist
    out = [None]*nprocs
    dist.all_gather_object(out, obj)
    return out


def barrier():
    if nprocs > 1:
        import torch.distributed as dist
        dist.barrier()


def slab_owner(x, boxsize, gridsize, n_ranks=None):
    """Owner rank of positions x (numpy or torch): the slab holding cell int(x·G/L).
    Same rule as owner_of() in csrc/pm_particles.cu (which_domain analogue, communication.py:756-772).
    A grid size the ranks do not divide has no slabs (pm_create refuses it, like fft.c:105-212); a component whose own
    grid is of that kind — it may still be analysed on other grids — is cut at x = r·L/P exactly, the slab boundaries of
    every grid the ranks do divide."""
    n_ranks = nprocs if n_ranks is None else n_ranks
    if gridsize % n_ranks:
        if isinstance(x, torch.Tensor):
            return torch.clamp((x*(n_ranks/boxsize)).to(torch.int64), 0, n_ranks - 1)
        import numpy as np
        return np.clip((np.asarray(x)*(n_ranks/boxsize)).astype(np.int64), 0, n_ranks - 1)
    nxl = gridsize//n_ranks
    if isinstance(x, torch.Tensor):
        cell = torch.clamp((x*(gridsize/boxsize)).to(torch.int64), 0, gridsize - 1)
        return cell//nxl
    import numpy as np
    cell = np.clip((np.asarray(x)*(gridsize/boxsize)).astype(np.int64), 0, gridsize - 1)
    return cell//nxl
