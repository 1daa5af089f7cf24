"""Build libpmgrav.so (CUDA, sm_100a) in-tree with nvcc.

    python -m concept_b200.build [--force] [--verbose]

The shared library lands in concept_b200/lib/ (git-ignored, but it travels to the GPU box with
the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libpmgrav.so')
STAMP = os.path.join(LIBDIR, 'libpmgrav.stamp')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
# --fmad=false: keep the reference's rounding (separate multiply and add; the reference is built
# for baseline x86-64 without FMA contraction).  Every kernel here is memory-bound.
NVCC_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '--fmad=false', '-Xcompiler', '-fPIC',
              '--expt-relaxed-constexpr']


def _nccl_paths():
    """Prefer the NCCL that PyTorch bundles (2.28.x) so that one libnccl.so.2 serves the process."""
    inc, lib = '/usr/include', '/usr/lib/x86_64-linux-gnu'
    try:
        import nvidia.nccl as n  # type: ignore
        base = os.path.dirname(n.__file__) if getattr(n, '__file__', None) else list(n.__path__)[0]
        if os.path.isfile(os.path.join(base, 'include', 'nccl.h')):
            inc = os.path.join(base, 'include')
            lib = os.path.join(base, 'lib')
    except Exception:
        pass
    return inc, lib


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _fingerprint():
    h = hashlib.sha256()
    for fn in _sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + [os.path.join(ROOT, 'include', 'pmgrav.h'), __file__]:
        with open(fn, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    fp = _fingerprint()
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP) and open(STAMP).read().strip() == fp:
        return LIB
    if not os.path.isfile(nvcc):
        if os.path.isfile(LIB):
            return LIB   # GPU box without a toolchain change: use the shipped binary
        raise RuntimeError('nvcc not found and no prebuilt libpmgrav.so')
    os.makedirs(LIBDIR, exist_ok=True)
    nccl_inc, nccl_lib = _nccl_paths()
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), 'lib64')
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, '-I', os.path.join(ROOT, 'include'), '-I', nccl_inc, '-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
            print(' '.join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f'--- {src}\n{out}\n')
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError('nvcc failed')
    link = [nvcc, *ARCH, '-shared', '-o', LIB, *objs, '-L', cuda_lib, '-L', nccl_lib, '-lcufft', '-l:libnccl.so.2',
            '-Xlinker', f'-rpath={nccl_lib}', '-Xlinker', f'-rpath={cuda_lib}',
            '-Xlinker', '-rpath=/usr/lib/x86_64-linux-gnu']
    if verbose:
        print(' '.join(link))
    subprocess.run(link, check=True)
    with open(STAMP, 'w') as f:
        f.write(fp)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
