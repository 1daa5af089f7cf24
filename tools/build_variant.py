"""Build an experiment variant of libpmgrav.so with extra nvcc flags (development aid):

    python tools/build_variant.py <name> [-DPM_FFT2D_OCC=2 ...]   ->  concept_b200/lib/libpmgrav_<name>.so

Select it at run time with PM_LIB_PATH=concept_b200/lib/libpmgrav_<name>.so (tools/stage_times.py, bench.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concept_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    nvcc = '/usr/local/cuda/bin/nvcc'
    out_dir = os.path.join(B.LIBDIR, f'var_{name}')
    os.makedirs(out_dir, exist_ok=True)
    nccl_inc, nccl_lib = B._nccl_paths()
    cuda_lib = '/usr/local/cuda/lib64'
    procs, objs = [], []
    for src in B._sources():
        obj = os.path.join(out_dir, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc, *B.ARCH, *B.NVCC_FLAGS, *flags, '-I', os.path.join(ROOT, 'include'), '-I', nccl_inc, '-c', src, '-o', obj]
        procs.append(subprocess.Popen(cmd))
        objs.append(obj)
    if any(p.wait() for p in procs):
        raise SystemExit('nvcc failed')
    lib = os.path.join(B.LIBDIR, f'libpmgrav_{name}.so')
    subprocess.run([nvcc, *B.ARCH, '-shared', '-o', lib, *objs, '-L', cuda_lib, '-L', nccl_lib, '-lcufft', '-l:libnccl.so.2',
                    '-Xlinker', f'-rpath={nccl_lib}', '-Xlinker', f'-rpath={cuda_lib}', '-Xlinker', '-rpath=/usr/lib/x86_64-linux-gnu'],
                   check=True)
    for o in objs:
        os.remove(o)
    print(lib)


if __name__ == '__main__':
    main()
