"""The host mirror's time loop on the CPU: concept_b200.main.timeloop — adaptive time-step controller, kick/drift
sequencing, dump times, static time-stepping — with the libpmgrav entry points replaced by their numpy model
(tests/ic_mock_context.py::PMKickMockContext, built from the oracle).  The reference's own 142-step run
(test/pure_python_pm configuration; tests/golden/run_pm_8.npz, made by tests/golden/gen_golden_run.py) is replayed;
the GPU counterpart is tests/test_gpu_host_api.py::test_timeloop_reproduces_reference_run."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
PM8 = '''
boxsize = 8*Mpc
potential_options = {'gridsize': {'gravity': {'pm': 8}}}
H0      = 70*km/s/Mpc
Ωcdm    = 0.25
Ωb      = 0.05
a_begin = 0.02
output_times = {'snapshot': (0.1, 0.5, 1)}
select_forces = {'matter': {'gravity': 'pm'}}
'''


def _run(monkeypatch, **overrides):
    import torch
    from concept_b200 import commons, main, mesh
    from concept_b200.species import Component
    from ic_mock_context import PMKickMockContext
    d = np.load(os.path.join(GOLDEN, 'run_pm_8.npz'))
    commons.load_params(PM8, **overrides)
    contexts = {}
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: contexts.setdefault(
        int(gridsize), PMKickMockContext(gridsize, commons.params.boxsize)))
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    c = Component('matter', 'matter', N=d['pos0'].shape[0], mass=float(d['mass']))
    c.populate(d['pos0'], 'pos')
    c.populate(d['mom0'], 'mom')
    snaps, steps = {}, []

    def on_dump(components, dump_time):
        snaps[f'{dump_time.a:.6f}'] = (components[0].pos_mv3.copy(), components[0].mom_mv3.copy(), commons.universals.t)
    nsteps = main.timeloop([c], on_dump=on_dump, on_step=lambda *a: steps.append(a))
    return d, nsteps, steps, snaps


def test_timeloop_reproduces_reference_run_on_the_cpu(monkeypatch):
    d, nsteps, steps, snaps = _run(monkeypatch)
    assert sorted(snaps) == ['0.100000', '0.500000', '1.000000']
    assert nsteps == len(d['drift_dt']) == 142 and len(steps) == 142
    L = float(d['boxsize'])
    for key, (pos, mom, t) in snaps.items():
        assert t == pytest.approx(float(d[f'snap_t_{key}']), rel=1e-11)
        dx = pos - d[f'snap_pos_{key}']
        dx -= L*np.round(dx/L)
        assert np.mean(np.sqrt((dx**2).sum(1)))/L < 1e-9      # the reference's own tolerance (test/nprocs_pm/analyze.py:121)
        assert np.abs(mom - d[f'snap_mom_{key}']).max() < 1e-7*np.abs(d[f'snap_mom_{key}']).max()


def test_static_timestepping_record_then_replay(monkeypatch, tmp_path):
    """A run records its time-stepping; a second run replaying the file takes the same steps (main.py:499-656, :897-912)"""
    path = str(tmp_path/'timestepping')
    _, nsteps_a, steps_a, snaps_a = _run(monkeypatch, static_timestepping=path)
    table = np.loadtxt(path)
    assert table.ndim == 2 and table.shape[1] == 2 and len(table) >= 10 and np.all(np.diff(table[:, 0]) >= 0)
    _, nsteps_b, steps_b, snaps_b = _run(monkeypatch, static_timestepping=path)
    assert nsteps_b == nsteps_a == 142
    # same base steps to the file's 9 significant digits, same final state to that accuracy
    Δt_a, Δt_b = np.array([s[3] for s in steps_a]), np.array([s[3] for s in steps_b])
    assert np.abs(Δt_b/Δt_a - 1).max() < 1e-6
    pos_a, pos_b = snaps_a['1.000000'][0], snaps_b['1.000000'][0]
    dx = pos_b - pos_a
    dx -= 8*np.round(dx/8)
    assert np.abs(dx).max() < 1e-4
