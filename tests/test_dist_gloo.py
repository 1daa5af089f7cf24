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


def _ic_worker(rank, world, port, out):
    """Replicated initial-condition realisation on two ranks (concept_b200.ic._get_context): the kernels are the
    numpy model of tests/ic_mock_context.py, everything else — parameters, noise, orchestration, slab ownership,
    all-gather of the result — is the product's host code over a real gloo process group."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from concept_b200 import commons, communication, ic, integration
    from concept_b200.species import Component
    from ic_mock_context import MockContext
    communication.init(backend='gloo')
    d = np.load(os.path.join(here, 'golden', 'ic_1lpt_sc_G8.npz'))
    commons.load_params(f"boxsize = {float(d['boxsize'])}*Mpc\nH0 = 70*km/s/Mpc\nΩcdm = 0.25\nΩb = 0.05\na_begin = {float(d['a'])}\n"
                        f"primordial_spectrum = {{'A_s': {float(d['A_s'])}, 'n_s': {float(d['n_s'])}, 'α_s': {float(d['alpha_s'])}, "
                        f"'pivot': {float(d['pivot'])}/Mpc}}\n"
                        "potential_options = {'gridsize': {'gravity': {'pm': 8}}}\nselect_forces = {'matter': {'gravity': 'pm'}}\n")
    integration.init_time()
    A_d, A_t, k0 = (float(x) for x in d['transfer'])
    a = float(d['a'])

    class Spline:
        def __init__(self, f):
            self.eval = f
    ic.compute_transfer = lambda component, variable, *args, **kw: (
        Spline((lambda k: -A_d*a*k**2/(1 + (k/k0)**2)**1.1) if variable == 0 else (lambda k: A_t*a**0.5*k**2/(1 + (k/k0)**2)**1.1)), None)
    ic._get_context = lambda gridsize: MockContext(gridsize, commons.params.boxsize)
    Component.device = property(lambda self: torch.device('cpu'))
    c = Component('matter', 'matter', N=8**3)
    ic.realize_particles(c, a)
    counts = communication.allgather(c.N_local)
    assert sum(counts) == 512 and all(n > 0 for n in counts)
    pos, mom = c.gather_global()          # all ranks' particles ordered by id
    L = float(d['boxsize'])
    dp = np.abs(pos - d['pos'])
    assert np.minimum(dp, L - dp).max() < 1e-11 and np.abs(mom - d['mom']).max() < 1e-11*np.abs(d['mom']).max()
    x = c.pos[:c.N_local, 0].numpy()
    assert np.all(communication.slab_owner(x, L, 8) == rank)
    communication.barrier()
    torch.distributed.destroy_process_group()
    out.put((rank, 'ok'))


def test_two_rank_replicated_initial_conditions_over_gloo():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ic_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [(0, 'ok'), (1, 'ok')]


def test_slab_owner_of_a_grid_the_ranks_do_not_divide():
    """A component whose own grid size (default 2·∛N) is not divisible by the number of ranks is still cut at x = r·L/P —
    the slab boundaries of the grids the ranks do divide.  (Cutting it by cells of its own grid put the boundaries of a
    78³ grid on 8 ranks up to 8.6 cells of a 128³ grid away from that grid's slabs: the P(k) deposit lost mass.)"""
    import numpy as np
    import torch
    from concept_b200 import communication
    L = 500.0
    x = np.random.default_rng(0).random(20000)*L
    for P in (4, 8):
        own = communication.slab_owner(x, L, 78, P)
        assert np.array_equal(own, np.minimum((x/L*P).astype(np.int64), P - 1))
        assert np.array_equal(communication.slab_owner(torch.as_tensor(x), L, 78, P).numpy(), own)
        # and it agrees with the cell rule of a divisible grid away from the cell that holds the boundary
        cells = communication.slab_owner(x, L, 128, P)
        assert np.array_equal(own, cells)
