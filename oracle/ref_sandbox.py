"""TEST INFRASTRUCTURE ONLY — builds a throw-away sandbox in which the
*unmodified* CO*N*CEPT reference (read-only at /root/reference) runs in its own
pure-Python mode, single rank, without MPI/FFTW/GSL/HDF5/CLASS.

Nothing here is imported by the product (concept_b200/). It is used by
tests/golden/gen_golden.py (run in the build container only — /root/reference
does not exist on the GPU box) to produce the committed golden vectors that pin
the oracle in oracle/pm_oracle.py.

Recipe:
  1. copy src/ + param/ to a scratch dir and write a .path file pointing into it
     (commons.py:1692-1731 finds .path via cwd / sys.path);
  2. put the parameter text at <scratch>/job/<jobid>/param (commons.py:1758-1782);
  3. stub modules on sys.path: mpi4py (size-1 communicator), blessings, matplotlib;
  4. two patches to the *copied* commons.py: drop the np.compat monkey patch
     (commons.py:486-491; NumPy >= 2 has no np.compat) and short-circuit
     get_matplotlib() (commons.py:509-515);
  5. dummy CLASS headers (linear.py:3719-3742 opens them at import time);
  6. sys.argv = ['x', "param='…'", 'jobid=N'] before `from commons import *`.

No reference source is copied into this repository; the copy lives in /tmp.
"""
import os
import shutil
import sys
import textwrap

REFERENCE = os.environ.get('CONCEPT_REFERENCE', '/root/reference')

_MPI4PY_INIT = '''
from . import rc
'''
_MPI4PY_RC = '''
threads = False
'''
_MPI4PY_MPI = '''
import numpy as np

SUM, MAX, MIN, LOR, LAND, PROD = 'SUM', 'MAX', 'MIN', 'LOR', 'LAND', 'PROD'
IN_PLACE = 'IN_PLACE'
ANY_SOURCE = -1
ANY_TAG = -1

def Get_processor_name():
    return 'localhost'
def Finalize():
    pass
def Is_finalized():
    return False
def Wtime():
    import time
    return time.time()

def _arr(buf):
    """Unwrap (buf, dtypechar) / (buf, counts) message tuples to the ndarray."""
    if isinstance(buf, (tuple, list)):
        buf = buf[0]
    return np.asarray(buf)

class _Request:
    def wait(self, *a, **k): return None
    Wait = wait
    def test(self): return (True, None)

class _Status:
    def Get_count(self, *a): return 0
    def Get_source(self): return 0
    def Get_tag(self): return 0

def Status():
    return _Status()

class _Comm:
    size = 1
    rank = 0
    def __init__(self):
        self._queue = []
        self._oqueue = []
    def Get_size(self): return 1
    def Get_rank(self): return 0
    def Barrier(self): pass
    def barrier(self): pass
    def Abort(self, code=1):
        raise SystemExit(code)
    # Upper-case (buffer) collectives: identity on one rank
    def _copy(self, sendbuf, recvbuf):
        if sendbuf is IN_PLACE or (isinstance(sendbuf, str) and sendbuf == IN_PLACE):
            return
        if recvbuf is None:
            return
        src = _arr(sendbuf).reshape(-1)
        dst = _arr(recvbuf).reshape(-1)
        dst[:src.size] = src
    def Allgather(self, sendbuf, recvbuf): self._copy(sendbuf, recvbuf)
    def Allgatherv(self, sendbuf, recvbuf): self._copy(sendbuf, recvbuf)
    def Gather(self, sendbuf, recvbuf, root=0): self._copy(sendbuf, recvbuf)
    def Gatherv(self, sendbuf, recvbuf, root=0): self._copy(sendbuf, recvbuf)
    def Allreduce(self, sendbuf, recvbuf, op=SUM): self._copy(sendbuf, recvbuf)
    def Reduce(self, sendbuf, recvbuf, op=SUM, root=0): self._copy(sendbuf, recvbuf)
    def Bcast(self, buf, root=0): pass
    def Sendrecv(self, sendbuf, dest=0, sendtag=0, recvbuf=None, source=0, recvtag=0, status=None):
        self._copy(sendbuf, recvbuf)
    def Isend(self, buf, dest=0, tag=0):
        self._queue.append(_arr(buf).copy())
        return _Request()
    def Send(self, buf, dest=0, tag=0):
        self._queue.append(_arr(buf).copy())
    def Recv(self, buf, source=0, tag=0, status=None):
        src = self._queue.pop(0).reshape(-1)
        dst = _arr(buf).reshape(-1)
        dst[:src.size] = src
    # Lower-case (object) collectives
    def allgather(self, obj): return [obj]
    def allreduce(self, obj, op=SUM): return obj
    def bcast(self, obj=None, root=0): return obj
    def gather(self, obj, root=0): return [obj]
    def reduce(self, obj, op=SUM, root=0): return obj
    def scatter(self, objs, root=0): return objs[0]
    def iprobe(self, source=0, tag=0, status=None): return bool(self._oqueue)
    def isend(self, obj, dest=0, tag=0):
        self._oqueue.append(obj)
        return _Request()
    def send(self, obj, dest=0, tag=0): self._oqueue.append(obj)
    def recv(self, buf=None, source=0, tag=0, status=None): return self._oqueue.pop(0)
    def sendrecv(self, sendobj, dest=0, sendtag=0, recvbuf=None, source=0, recvtag=0, status=None):
        return sendobj

COMM_WORLD = _Comm()
COMM_SELF = COMM_WORLD
'''
_BLESSINGS = '''
class _Fmt(str):
    def __call__(self, *args):
        return ''.join(str(a) for a in args)
class Terminal:
    width = 120
    height = 40
    def __init__(self, *a, **k): pass
    def __getattr__(self, name):
        return _Fmt('')
'''
_MATPLOTLIB = '''
import types, sys
class _Anything:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Anything()
    def __getattr__(self, name): return _Anything()
    def __iter__(self): return iter(())
    def __getitem__(self, k): return _Anything()
    def __setitem__(self, k, v): pass
def _modattr(name):
    if name.startswith('__'):
        raise AttributeError(name)
    return _Anything()
rcParams = {}
def use(*a, **k): pass
class _ColorConverter:
    _named = {'k': (0., 0., 0.), 'w': (1., 1., 1.), 'r': (1., 0., 0.), 'g': (0., .5, 0.),
              'b': (0., 0., 1.), 'c': (0., .75, .75), 'm': (.75, 0., .75), 'y': (.75, .75, 0.)}
    def to_rgb(self, c):
        if isinstance(c, str):
            if c in self._named: return self._named[c]
            if c.startswith('#') and len(c) == 7:
                return tuple(int(c[i:i+2], 16)/255 for i in (1, 3, 5))
            if c.startswith('C') and c[1:].isdigit():
                return (.12, .47, .71)
            try:
                g = float(c); return (g, g, g)
            except ValueError:
                return (.5, .5, .5)
        c = tuple(float(x) for x in c)
        return c[:3]
colors = types.ModuleType('matplotlib.colors')
colors.ColorConverter = _ColorConverter
colors.to_rgb = _ColorConverter().to_rgb
colors.CSS4_COLORS = {}
colors.__getattr__ = _modattr
sys.modules['matplotlib.colors'] = colors
pyplot = types.ModuleType('matplotlib.pyplot')
pyplot.__getattr__ = _modattr
sys.modules['matplotlib.pyplot'] = pyplot
cm = types.ModuleType('matplotlib.cm')
cm.__getattr__ = _modattr
sys.modules['matplotlib.cm'] = cm
def __getattr__(name):
    return _modattr(name)
'''


def _write(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'w', encoding='utf-8') as f:
        f.write(textwrap.dedent(text).lstrip('\n'))


def build_sandbox(dest):
    """Create the sandbox at `dest` (removed first if it exists)."""
    if not os.path.isdir(f'{REFERENCE}/src'):
        raise RuntimeError(f'reference not found at {REFERENCE} (only present in the build container)')
    if os.path.exists(dest):
        shutil.rmtree(dest)
    os.makedirs(dest)
    shutil.copytree(f'{REFERENCE}/src', f'{dest}/src')
    shutil.copytree(f'{REFERENCE}/param', f'{dest}/param')
    for fn in os.listdir(f'{dest}/src'):
        os.chmod(f'{dest}/src/{fn}', 0o644)
    os.chmod(f'{dest}/src', 0o755)
    # .path
    dirs = ['build', 'dep', 'doc', 'ic', 'job', 'output', 'param', '.reusable', 'src', 'test', '.tmp', 'util']
    names = ['build_dir', 'dep_dir', 'doc_dir', 'ic_dir', 'job_dir', 'output_dir', 'param_dir',
             'reusable_dir', 'src_dir', 'test_dir', 'tmp_dir', 'util_dir']
    lines = [f"concept_dir='{dest}'"]
    for n, d in zip(names, dirs):
        os.makedirs(f'{dest}/{d}', exist_ok=True)
        lines.append(f"{n}='{dest}/{d}'")
    for n, d in [('blas_dir', 'dep/openblas'), ('class_dir', 'dep/class'), ('fftw_dir', 'dep/fftw'),
                 ('fftw_for_gadget_dir', 'dep/gadget/fftw'), ('gadget_dir', 'dep/gadget'),
                 ('Gadget2_dir', 'dep/gadget/Gadget2'), ('gsl_dir', 'dep/gsl'), ('hdf5_dir', 'dep/hdf5'),
                 ('mpi_dir', 'dep/mpich'), ('mpi_compilerdir', 'dep/mpich/bin'), ('mpi_bindir', 'dep/mpich/bin'),
                 ('mpi_libdir', 'dep/mpich/lib'), ('mpi_includedir', 'dep/mpich/include'),
                 ('mpi_symlinkdir', 'dep/.mpi_symlinks'), ('python_dir', 'dep/python'), ('zlib_dir', 'dep/zlib')]:
        lines.append(f"{n}='{dest}/{d}'")
    for n, f in [('concept', 'concept'), ('env', '.env'), ('install', 'install'),
                 ('mpicc', 'dep/mpich/bin/mpicc'), ('mpiexec', 'dep/mpich/bin/mpiexec')]:
        lines.append(f"{n}='{dest}/{f}'")
    _write(f'{dest}/.path', '\n'.join(lines) + '\n')
    _write(f'{dest}/.env', '')
    # Stubs
    _write(f'{dest}/stubs/mpi4py/__init__.py', _MPI4PY_INIT)
    _write(f'{dest}/stubs/mpi4py/rc.py', _MPI4PY_RC)
    _write(f'{dest}/stubs/mpi4py/MPI.py', _MPI4PY_MPI)
    _write(f'{dest}/stubs/blessings.py', _BLESSINGS)
    _write(f'{dest}/stubs/matplotlib/__init__.py', _MATPLOTLIB)
    # Dummy CLASS files grepped at import by linear.py
    _write(f'{dest}/dep/class/include/common.h', '#define _VERSION_ "v2.7.2"\n')
    _write(f'{dest}/dep/class/include/parser.h', '#define _ARGUMENT_LENGTH_MAX_ 1024\n')
    _write(f'{dest}/dep/class/source/perturbations.c', '\n')
    # Patches to the COPY of commons.py
    fn = f'{dest}/src/commons.py'
    with open(fn, encoding='utf-8') as f:
        src = f.read()
    out = []
    for line in src.split('\n'):
        if line.startswith('np.compat.py3k'):
            line = '# [sandbox] ' + line
        out.append(line)
        if line.startswith('def get_matplotlib():'):
            out.append('    import matplotlib; return matplotlib  # [sandbox] stub')
    with open(fn, 'w', encoding='utf-8') as f:
        f.write('\n'.join(out))
    return dest


def enter_reference(dest, param_text, jobid=1, extra_argv=()):
    """Prepare this *process* to `from commons import *` with the given parameter text.
    One parameter set per process: the reference turns parameters into module globals."""
    os.makedirs(f'{dest}/job/{jobid}', exist_ok=True)
    with open(f'{dest}/job/{jobid}/param', 'w', encoding='utf-8') as f:
        f.write(param_text)
    os.chdir(dest)
    sys.path.insert(0, f'{dest}/src')
    sys.path.insert(0, f'{dest}/stubs')
    sys.argv = ['x', f"param='{dest}/job/{jobid}/param'", f'jobid={jobid}', *extra_argv]


if __name__ == '__main__':
    print(build_sandbox(sys.argv[1] if len(sys.argv) > 1 else '/tmp/concept_ref_sandbox'))
