"""CPU-only, world_size 2 over gloo: the host-side logic of the N>1 path — process-group helpers and
the slab ownership rule that pm_exchange / Component.set_particles rely on."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from concept_b200 import communication
    communication.init(backend='gloo')
    assert (communication.rank, communication.nprocs, communication.master) == (rank, world, rank == 0)
    # broadcast of the 128-byte communicator id, as mesh.get_context does
    uid = communication.bcast(bytes(range(128)) if rank == 0 else None)
    assert uid == bytes(range(128))
    # Σ mom² style reduction and object gather
    assert communication.allreduce_sum(float(rank + 1)) == 3.0
    assert communication.allgather(rank*10) == [0, 10]
    # slab ownership: every particle has exactly one owner; owners tile the box in x
    G, L = 16, 10.0
    rng = np.random.default_rng(0)
    x = rng.random(1000)*L
    x[:4] = [0.0, np.nextafter(L, 0), L/2, np.nextafter(L/2, 0)]
    owner = communication.slab_owner(x, L, G)
    mine = int((owner == rank).sum())
    counts = communication.allgather(mine)
    assert sum(counts) == 1000
    assert owner[0] == 0 and owner[1] == world - 1 and owner[2] == 1 and owner[3] == 0
    t = communication.slab_owner(torch.as_tensor(x), L, G)
    assert np.array_equal(t.numpy(), owner)
    communication.barrier()
    torch.distributed.destroy_process_group()
    out.put((rank, 'ok'))


def test_two_rank_host_logic_over_gloo():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [(0, 'ok'), (1, 'ok')]
