"""GADGET-2 snapshot I/O against files written by the unmodified reference
(tests/golden/gen_golden_snapshot.py): our writer must reproduce them byte for byte, our reader must
return the particle data the reference stored."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, 'golden', 'snapshot_gadget_*.npz')))


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_writer_is_byte_identical_to_reference(path, tmp_path):
    from concept_b200 import snapshot
    d = np.load(path)
    fn = str(tmp_path/'snap')
    snapshot.write_gadget(fn, d['pos'], d['mom'], mass=float(d['mass']), a=float(d['a']), boxsize=float(d['boxsize']),
                          H0=float(d['H0']), Ωm=float(d['Omega_m']))
    ours = np.frombuffer(open(fn, 'rb').read(), dtype=np.uint8)
    ref = d['file_bytes']
    assert len(ours) == len(ref)
    diff = np.nonzero(ours != ref)[0]
    assert diff.size == 0, f'{diff.size} differing bytes, first at offset {diff[0]}'


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_reader_recovers_reference_particles(path, tmp_path):
    from concept_b200 import snapshot
    d = np.load(path)
    fn = str(tmp_path/'snap')
    with open(fn, 'wb') as f:
        f.write(d['file_bytes'].tobytes())
    s = snapshot.read_gadget(fn)
    L = float(d['boxsize'])
    assert len(s['pos']) == len(d['pos'])
    assert abs(s['mass']/float(d['mass']) - 1) < 1e-14 and abs(s['boxsize']/L - 1) < 1e-14
    assert s['a'] == float(d['a']) and abs(s['H0']/float(d['H0']) - 1) < 1e-14
    assert np.max(np.abs(s['pos'] - d['pos'])) < 1e-7*L                    # float32 storage
    assert np.max(np.abs(s['mom'] - d['mom'])) < 1e-6*np.max(np.abs(d['mom']))
    assert np.array_equal(s['ids'], np.arange(len(d['pos'])))


def test_position_wrap_and_double_precision(tmp_path):
    """Positions that reach BoxSize in the stored precision wrap to 0 (snapshot.py:1381-1382); 64-bit blocks."""
    from concept_b200 import commons, snapshot
    L = 64.0
    pos = np.array([[np.nextafter(L, 0), 1.0, 2.0], [3.0, 4.0, 5.0]])
    mom = np.zeros((2, 3))
    kw = dict(mass=1.0, a=1.0, boxsize=L, H0=70*commons.units.km/(commons.units.s*commons.units.Mpc), Ωm=0.3)
    fn = str(tmp_path/'s32')
    snapshot.write_gadget(fn, pos, mom, **kw)
    assert snapshot.read_gadget(fn)['pos'][0, 0] == 0.0
    fn = str(tmp_path/'s64')
    snapshot.write_gadget(fn, pos, mom, bits_pos=64, bits_vel=64, **kw)
    assert abs(snapshot.read_gadget(fn)['pos'][0, 0] - pos[0, 0]) < 1e-12


def _rewrite_header(path, **fields):
    """patch header fields of a SnapFormat-2 file in place (the HEAD block starts 16 + 4 bytes into the file)"""
    import struct
    from concept_b200 import snapshot
    with open(path, 'r+b') as f:
        off = 20
        for name, fmt in snapshot.HEADER_FIELDS:
            if name in fields:
                f.seek(off)
                vals = fields[name] if isinstance(fields[name], (list, tuple)) else [fields[name]]
                f.write(struct.pack('<' + fmt, *vals))
            off += struct.calcsize('<' + fmt)


def test_reader_of_a_snapshot_spread_over_several_files(tmp_path):
    """GADGET snapshots of large runs come as name.0, name.1, … (header NumFiles > 1, Nall = total, Npart = this file) —
    as files next to each other or inside a directory (how the reference writes them, snapshot.py:1316-1323)."""
    from concept_b200 import commons, snapshot
    commons.load_params('boxsize = 32*Mpc\nH0 = 70*km/(s*Mpc)\nΩb = 0.05\nΩcdm = 0.25\n')
    p = commons.params
    rng = np.random.default_rng(4)
    N, cuts = 1000, [0, 300, 650, 1000]
    pos, mom = rng.random((N, 3))*p.boxsize, rng.standard_normal((N, 3))
    kw = dict(mass=2.5, a=0.5, boxsize=p.boxsize, H0=p.H0, Ωm=p.Ωm, bits_pos=64, bits_vel=64)
    whole = tmp_path/'whole'
    snapshot.write_gadget(str(whole), pos, mom, **kw)
    ref = snapshot.read_gadget(str(whole))
    for layout in ('flat', 'directory'):
        base = tmp_path/layout
        if layout == 'directory':
            base.mkdir()
        for i, (lo, hi) in enumerate(zip(cuts, cuts[1:])):
            name = str(base/f'snapshot.{i}') if layout == 'directory' else f'{base}.{i}'
            snapshot.write_gadget(name, pos[lo:hi], mom[lo:hi], ids=np.arange(lo, hi), **kw)
            nall = [0]*6
            nall[snapshot.HALO] = N
            _rewrite_header(name, NumFiles=3, Nall=nall)
        got = snapshot.read_gadget(str(base))
        assert all(np.array_equal(got[key], ref[key]) for key in ('pos', 'mom', 'ids'))
        assert got['mass'] == ref['mass'] and got['a'] == 0.5
        with pytest.raises(commons.ConceptAbort):       # one file of several is not a snapshot
            snapshot.read_gadget(str(base/'snapshot.1') if layout == 'directory' else f'{base}.1')


def test_reader_of_snapformat_1(tmp_path):
    """SnapFormat 1 (what GADGET's own initial-condition generators write by default): the same records without the
    16-byte name records in front of them, in the order HEAD, POS, VEL, ID."""
    from concept_b200 import commons, snapshot
    commons.load_params('boxsize = 32*Mpc\nH0 = 70*km/(s*Mpc)\nΩb = 0.05\nΩcdm = 0.25\n')
    p = commons.params
    rng = np.random.default_rng(5)
    pos, mom = rng.random((400, 3))*p.boxsize, rng.standard_normal((400, 3))
    two = tmp_path/'format2'
    snapshot.write_gadget(str(two), pos, mom, mass=1.5, a=0.25, boxsize=p.boxsize, H0=p.H0, Ωm=p.Ωm)
    blob, out, o = two.read_bytes(), b'', 0
    while o < len(blob):                      # drop the name records
        size = int.from_bytes(blob[o + 16:o + 20], 'little')
        out += blob[o + 16:o + 16 + size + 8]
        o += 16 + size + 8
    one = tmp_path/'format1'
    one.write_bytes(out)
    a, b = snapshot.read_gadget(str(one)), snapshot.read_gadget(str(two))
    assert all(np.array_equal(a[key], b[key]) for key in ('pos', 'mom', 'ids')) and a['mass'] == b['mass'] and a['a'] == 0.25
    with pytest.raises(commons.ConceptAbort):
        (tmp_path/'junk').write_bytes(b'\x01'*400)
        snapshot.read_gadget(str(tmp_path/'junk'))


def test_cli_utilities_info_and_powerspec(monkeypatch, tmp_path, capsys, host_kernels):
    """`python -m concept_b200 -u info SNAPSHOT` and `-u powerspec SNAPSHOT` (the reference's util/info, util/powerspec;
    utilities.py:465-497): box and cosmology come from the snapshot when no parameter file is given, the spectrum is written
    next to the snapshot as powerspec_<name> and equals analysis.powerspec of the loaded component (kernels replaced by
    their numpy model)."""
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from concept_b200 import __main__ as cli, analysis, commons, mesh, snapshot
    from concept_b200.species import Component
    import ic_mock_context
    monkeypatch.setattr(ic_mock_context.MeshMockContext, 'lib', host_kernels)
    contexts = {}
    monkeypatch.setattr(mesh, 'get_context', lambda gridsize, dtype=None: contexts.setdefault(
        int(gridsize), ic_mock_context.MeshMockContext(gridsize, commons.params.boxsize)))
    monkeypatch.setattr(mesh, 'free_contexts', lambda: contexts.clear())
    monkeypatch.setattr(Component, 'device', property(lambda self: torch.device('cpu')))
    commons.load_params('boxsize = 40*Mpc\nH0 = 70*km/(s*Mpc)\nΩb = 0.049\nΩcdm = 0.251\n')
    p = commons.params
    rng = np.random.default_rng(6)
    pos, mom = rng.random((8**3, 3))*p.boxsize, rng.standard_normal((8**3, 3))
    path = str(tmp_path/'snapshot_a=0.50')
    snapshot.write_gadget(path, pos, mom, mass=3.0, a=0.5, boxsize=p.boxsize, H0=p.H0, Ωm=p.Ωm, bits_pos=64, bits_vel=64)
    assert cli.cli(['-u', 'info', path]) == 0
    out = capsys.readouterr().out
    assert 'particles        512' in out and 'a                0.5 ' in out and 'boxsize          40 Mpc' in out
    assert cli.cli(['-u', 'powerspec', path]) == 0
    table = np.loadtxt(str(tmp_path/'powerspec_snapshot_a=0.50'))
    assert commons.universals.a == 0.5 and commons.params.boxsize == pytest.approx(40.0, rel=1e-12)
    c = snapshot.load(path)
    k, power, n_modes = analysis.powerspec([c], 16, gridsizes_upstream=[16])
    assert table.shape == (len(k), 4) and np.allclose(table[:, 0], k, rtol=1e-7) and np.allclose(table[:, 2], power, rtol=1e-7)
    assert np.array_equal(table[:, 1], n_modes)


@pytest.mark.gpu
def test_component_save_load_round_trip(tmp_path):
    """snapshot.save / snapshot.load through a GPU-resident Component."""
    pytest.importorskip('torch')
    from concept_b200 import commons, mesh, snapshot
    from concept_b200.species import Component
    L, N = 100.0, 5000
    commons.load_params(f'boxsize = {L}*Mpc\nH0 = 70*km/s/Mpc\nΩcdm = 0.25\nΩb = 0.05\n')
    commons.universals.a = 0.25
    rng = np.random.default_rng(3)
    pos, mom = rng.random((N, 3))*L, rng.standard_normal((N, 3))*5
    c = Component('matter', 'matter', N=N, mass=0.8)
    c.set_particles(pos, mom)
    fn = snapshot.save(c, str(tmp_path/'snap'))
    commons.universals.a = 1.0
    c2 = snapshot.load(fn)
    assert commons.universals.a == 0.25 and c2.N == N and abs(c2.mass/0.8 - 1) < 1e-14
    p2, m2 = c2.gather_global()
    assert np.max(np.abs(p2 - pos)) < 1e-6*L and np.max(np.abs(m2 - mom)) < 1e-6*np.max(np.abs(mom))
    mesh.free_contexts()
