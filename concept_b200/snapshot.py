"""GADGET-2 snapshots: exchange particle data with a real CO*N*CEPT / GADGET run.

Mirrors the reference's GadgetSnapshot (snapshot.py:640-2640) for what the PM path needs: one particle
component stored as GADGET type 1 ("halo"), SnapFormat 2 (named blocks), written as a single file and read from one
or several files, POS and VEL in
32 or 64 bits, 32-bit IDs (64-bit above 2³² particles), GADGET units kpc/h, km/s, 10¹⁰ m☉/h
(commons.py:2787-2804).  Conversions (snapshot.py:1520-1552, :1376-1383, :2560-2573):

    POS   = pos/(kpc/h),  values that reach BoxSize in the stored precision wrap to 0
    VEL   = mom/(km/s · mass · a^1.5)            (peculiar velocity u = a·dx/dt divided by √a)
    Massarr[1] = mass/(10¹⁰ m☉/h),  BoxSize = boxsize/(kpc/h),  Time = a,  Redshift = 1/a − 1,
    Omega0 = Ωm,  OmegaLambda = 1 − Ωm,  HubbleParam = H0/(100 km/s/Mpc)

The array-level functions (`write_gadget`, `read_gadget`) are pure numpy so that they are pinned on the
CPU against files written by the unmodified reference (tests/test_snapshot.py: byte-identical);
`save`/`load` wrap them for `Component`s whose particles live on the GPU.
"""
import struct

import numpy as np

from . import commons

NUM_TYPES = 6
HEADER_SIZE = 256
# (name, struct format) in file order — GadgetSnapshot.header_fields (snapshot.py:673-691)
HEADER_FIELDS = [('Npart', '6I'), ('Massarr', '6d'), ('Time', 'd'), ('Redshift', 'd'), ('FlagSfr', 'i'),
                 ('FlagFeedback', 'i'), ('Nall', '6I'), ('FlagCooling', 'i'), ('NumFiles', 'i'), ('BoxSize', 'd'),
                 ('Omega0', 'd'), ('OmegaLambda', 'd'), ('HubbleParam', 'd'), ('FlagAge', 'i'), ('FlagMetals', 'i'),
                 ('NallHW', '6I'), ('flag_entr_ics', 'i')]
HALO = 1   # GADGET particle type of a matter component


def _units(h):
    """'kpc/h', 'km/s', '10¹⁰ m☉/h' (commons.py:2799-2803) evaluated like eval_unit does: a division is a
    multiplication by the −1st power (commons.py:1647-1700), which fixes the last bit."""
    u = commons.units
    return u.kpc*h**(-1), u.km*u.s**(-1), 1e10*u.m_sun*h**(-1)


def correct_float(val_raw):
    """commons.py:5356-5388, applied by GadgetSnapshot.write to every double of the header
    (snapshot.py:1731-1733): snap a value to the shortest decimal representation within ±10 ϵ,
    e.g. 44799.99999999999 → 44800.0."""
    val_raw = float(val_raw)
    val_g = float(f'{val_raw:g}')
    if val_g == val_raw:
        return val_g
    val_str = str(abs(val_raw))
    if 'e' in val_str:
        val_str = val_str[:val_str.index('e')]
    val_str = val_str.replace('.', '')
    if len(val_str) < 15:
        return val_raw
    eps = 2.220446049250313e-16
    lower, upper = val_raw*(1 - 10*eps), val_raw*(1 + 10*eps)
    val_correct = val_new = lower
    while val_new <= upper:
        if len(str(val_new)) < len(str(val_correct)):
            val_correct = val_new
        val_new = float(np.nextafter(val_new, np.inf))
    return val_correct if len(str(val_correct)) < len(str(val_raw)) - 2 else val_raw


def _block(f, name, payload):
    """SnapFormat 2 block: 16-byte name record (name + size of the data record incl. its two markers)
    followed by the Fortran-style data record (snapshot.py:1671-1711).  payload: bytes or a contiguous numpy array
    (written straight from its buffer)."""
    size = payload.nbytes if isinstance(payload, np.ndarray) else len(payload)
    f.write(struct.pack('<I4sII', 8, name.ljust(4).encode('ascii'), size + 8, 8))
    f.write(struct.pack('<I', size))
    f.write(memoryview(payload).cast('B') if isinstance(payload, np.ndarray) else payload)
    f.write(struct.pack('<I', size))


def write_gadget(filename, pos, mom, *, mass, a, boxsize, H0, Ωm, ids=None, bits_pos=32, bits_vel=32):
    """Write one particle component (arrays in internal units, shape (N, 3)) as a single-file GADGET-2
    snapshot.  Returns the header dict."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    mom = np.ascontiguousarray(mom, dtype=np.float64)
    N = pos.shape[0]
    h = H0/(100*commons.units.km/(commons.units.s*commons.units.Mpc))
    unit_length, unit_velocity, unit_mass = _units(h)
    header = {name: ([0]*NUM_TYPES if fmt[0] == '6' and fmt[1] == 'I' else [0.0]*NUM_TYPES if fmt[0] == '6'
                     else (0.0 if fmt == 'd' else 0)) for name, fmt in HEADER_FIELDS}
    header['Npart'][HALO] = N % 2**32
    header['Nall'][HALO] = N % 2**32
    header['NallHW'][HALO] = N//2**32
    header['Massarr'][HALO] = mass/unit_mass
    header['Time'] = a
    header['Redshift'] = 1/a - 1
    header['NumFiles'] = 1
    header['BoxSize'] = boxsize/unit_length
    header['Omega0'] = Ωm
    header['OmegaLambda'] = 1 - Ωm
    header['HubbleParam'] = h
    def packed(fmt, v):
        vals = v if isinstance(v, list) else [v]
        if fmt[-1] == 'd':
            vals = [correct_float(x) for x in vals]
        return struct.pack('<' + fmt, *vals)
    raw = b''.join(packed(fmt, header[name]) for name, fmt in HEADER_FIELDS)
    raw += b'\0'*(HEADER_SIZE - len(raw))

    def convert(data, unit, bits, wrap):
        dtype = np.float32 if bits == 32 else np.float64
        # the product in fp64, rounded once to the file's precision (no fp64 temporary of the whole array)
        out = np.empty(data.size, dtype=dtype)
        np.multiply(data.reshape(-1), 1/unit, out=out, casting='same_kind')
        if wrap:     # round-off guard of the writer (snapshot.py:1381-1382)
            box = dtype(boxsize/unit)
            over = np.flatnonzero(out >= box)
            out[over] -= box
        return out
    with open(filename, 'wb') as f:
        _block(f, 'HEAD', raw)
        _block(f, 'POS', convert(pos, unit_length, bits_pos, True))
        _block(f, 'VEL', convert(mom, unit_velocity*mass*a**1.5, bits_vel, False))
        id_dtype = np.uint32 if N <= 2**32 else np.uint64
        ids = np.arange(N, dtype=id_dtype) if ids is None else np.ascontiguousarray(np.asarray(ids).astype(id_dtype))
        _block(f, 'ID', ids)
    return header


def _gadget_files(filename):
    """The files of a snapshot: `filename` itself, or — a snapshot spread over several files (header NumFiles > 1) —
    `filename.0`, `filename.1`, … next to each other, or the `*.0`, `*.1`, … inside the directory `filename` (how the
    reference writes them, snapshot.py:1316-1323)."""
    import glob
    import os
    import re
    if os.path.isfile(filename):
        return [filename]
    if os.path.isdir(filename):
        files = [f for f in glob.glob(os.path.join(filename, '*.*')) if re.search(r'\.\d+$', f)]
    else:
        files = [f for f in glob.glob(filename + '.*') if re.search(r'\.\d+$', f)]
    if not files:
        commons.abort(f'No GADGET snapshot at "{filename}"')
    return sorted(files, key=lambda f: int(f.rsplit('.', 1)[1]))


def read_gadget(filename):
    """Read a GADGET snapshot (SnapFormat 2 with named blocks, or the older SnapFormat 1) with one populated particle type,
    in one file or spread over several.
    Returns dict(header, pos, mom, ids, mass, a, boxsize, H0, Ωm) in internal units."""
    files = _gadget_files(filename)
    if len(files) == 1:
        return _read_gadget_file(files[0], single=True)
    parts = [_read_gadget_file(f, single=False) for f in files]
    first = parts[0]
    if first['header']['NumFiles'] != len(files):
        commons.abort(f'"{filename}": {len(files)} files found, the header says NumFiles = {first["header"]["NumFiles"]}')
    out = dict(first)
    out['pos'] = np.concatenate([q['pos'] for q in parts])
    out['mom'] = np.concatenate([q['mom'] for q in parts])
    out['ids'] = None if any(q['ids'] is None for q in parts) else np.concatenate([q['ids'] for q in parts])
    t = first['type']
    n_all = first['header']['Nall'][t] + 2**32*first['header']['NallHW'][t]
    if len(out['pos']) != n_all:
        commons.abort(f'"{filename}": the files hold {len(out["pos"])} particles, the header says {n_all}')
    return out


def _read_gadget_file(filename, single):
    with open(filename, 'rb') as f:
        blob = f.read()
    blocks, o = {}, 0
    if len(blob) < 264:
        commons.abort(f'"{filename}" is not a GADGET snapshot')
    snapformat = {8: 2, HEADER_SIZE: 1}.get(struct.unpack_from('<I', blob, 0)[0])
    if snapformat is None:
        commons.abort(f'"{filename}" is not a GADGET snapshot of SnapFormat 1 or 2')
    unnamed = iter(('HEAD', 'POS', 'VEL', 'ID', 'MASS'))       # SnapFormat 1: the blocks come in this order, without names
    while o < len(blob):
        if snapformat == 2:
            n, name, _, n2 = struct.unpack_from('<I4sII', blob, o)
            if n != 8 or n2 != 8:
                commons.abort(f'"{filename}" is not a SnapFormat 2 GADGET snapshot')
            name = name.decode('ascii').strip()
            o += 16
        else:
            name = next(unnamed, None)
            if name is None:
                break                                           # further blocks (gas, …) are of no concern here
        size = struct.unpack_from('<I', blob, o)[0]
        payload = memoryview(blob)[o + 4:o + 4 + size]          # a view: the blocks are hundreds of MB
        if o + 8 + size > len(blob) or struct.unpack_from('<I', blob, o + 4 + size)[0] != size:
            commons.abort(f'Corrupt block "{name}" in "{filename}"')
        blocks[name] = payload
        o += size + 8
    head, header, off = blocks['HEAD'], {}, 0
    for name, fmt in HEADER_FIELDS:
        vals = struct.unpack_from('<' + fmt, head, off)
        header[name] = list(vals) if len(vals) > 1 else vals[0]
        off += struct.calcsize('<' + fmt)
    types = [t for t in range(NUM_TYPES) if header['Npart'][t]]
    if len(types) != 1:
        commons.abort('concept_b200 reads GADGET snapshots with one particle type')
    if single and header['NumFiles'] > 1:
        commons.abort(f'"{filename}" is one of {header["NumFiles"]} files of a snapshot: give the common name (without .0)')
    t = types[0]
    N = header['Npart'][t]
    h, a = header['HubbleParam'], header['Time']
    unit_length, unit_velocity, unit_mass = _units(h)
    mass = header['Massarr'][t]*unit_mass
    if mass == 0:
        commons.abort('GADGET snapshots with individual particle masses are not supported')

    def floats(payload):
        return np.frombuffer(payload, dtype='<f4' if len(payload) == 12*N else '<f8').astype(np.float64).reshape(N, 3)
    pos = floats(blocks['POS'])*unit_length
    mom = floats(blocks['VEL'])*(unit_velocity*mass*a**1.5)
    ids = None
    if 'ID' in blocks:
        ids = np.frombuffer(blocks['ID'], dtype='<u4' if len(blocks['ID']) == 4*N else '<u8').astype(np.int64)
    u = commons.units
    return dict(header=header, pos=pos, mom=mom, ids=ids, mass=mass, a=a, boxsize=header['BoxSize']*unit_length,
                H0=h*100*u.km/(u.s*u.Mpc), Ωm=header['Omega0'], type=t)


def save(component, filename):
    """snapshot.save(component, filename) for snapshot_type = 'gadget' (snapshot.py:3045-3119): gathers the
    component from all ranks; the master writes."""
    from . import communication
    p = commons.params
    pos, mom = component.gather_global()
    # gadget_snapshot_params['dataformat'] (commons.py:2581-2600): 32 or 64 bits for positions and velocities
    fmt = commons.user_params.get('gadget_snapshot_params', {})
    fmt = {str(k).upper(): v for k, v in dict(fmt.get('dataformat', {})).items()} if isinstance(fmt, dict) else {}
    bits = {key: (int(fmt[key]) if str(fmt.get(key, 32)).isdigit() else 32) for key in ('POS', 'VEL')}
    if communication.master:
        write_gadget(filename, pos, mom, mass=component.mass, a=commons.universals.a, boxsize=p.boxsize, H0=p.H0,
                     Ωm=p.Ωb + p.Ωcdm, bits_pos=bits['POS'], bits_vel=bits['VEL'])
    communication.barrier()
    return filename


def load(filename, name='matter', species='matter'):
    """snapshot.load(filename) (snapshot.py:3120-3205) → a Component distributed over the ranks by x-slab."""
    from .species import Component
    d = read_gadget(filename)
    if abs(d['boxsize']/commons.params.boxsize - 1) > 1e-6:
        commons.abort(f'Snapshot boxsize {d["boxsize"]} differs from the boxsize parameter {commons.params.boxsize}')
    c = Component(name, species, N=len(d['pos']), mass=d['mass'])
    pos = np.mod(d['pos'], commons.params.boxsize)
    c.set_particles(pos, d['mom'], ids=d['ids'])
    commons.universals.a = d['a']
    return c
